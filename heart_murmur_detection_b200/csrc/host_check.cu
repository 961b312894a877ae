// CPU emulation of the warp algorithm in logmel.cu / kaldi_fbank.cu.
//
// The build container has no GPU.  The lane-level phases (fft_core.cuh, logmel_core.cuh)
// are __host__ __device__ templates; this program runs exactly the same arithmetic with a
// loop over the 32 lanes standing in for the warp, so index math (bit reversal, 32x32
// split, twiddle planes, exchange tile, partner lanes, banded mel slots) and float32
// accuracy are validated against the oracle before any GPU time is spent
// (tests/test_host_emulation.py).  Test infrastructure, not part of libhmfe.so.
//
// usage: host_check logmel <scalar|packed> <hop> <n_mels> <fmin> <fmax> <in.f32> <out.f32>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "logmel_core.cuh"
#include "tables.h"

using namespace hmfe;

template <typename V>
static void run_logmel(const std::vector<float>& x, int hop, int n_mels, double fmin, double fmax, std::vector<float>& out) {
    constexpr int NV = lanes_of<V>::value, FR = 2 * NV;
    const int nsamp = (int)x.size();
    const int T = 1 + nsamp / hop;
    const std::vector<float> dense = mel_filterbank_slaney(16000, kNfft, n_mels, fmin, fmax);
    const BandedMel bm = build_banded(dense, n_mels, kNfft / 2 + 1);
    const std::vector<float> win = half_hann_periodic(kNfft);
    const std::vector<float> twv = twiddle_plane(kNfft, 32);
    const float2* tw = reinterpret_cast<const float2*>(twv.data());
    out.assign((size_t)T * n_mels, 0.0f);
    std::vector<xelem<V>> tile(32 * kXStride);
    struct Lane {
        V re[32], im[32];
    };
    std::vector<Lane> L(32);
    for (int f0 = 0; f0 < T; f0 += FR) {
        for (int lane = 0; lane < 32; ++lane) {
            auto fetch = [&](int t, bool second, int n) -> float {
                const int f = f0 + 2 * t + (second ? 1 : 0);
                if (f >= T) return 0.0f;
                const int i = f * hop - kNfft / 2 + n;
                return (i >= 0 && i < nsamp) ? x[i] : 0.0f;
            };
            load_window<V>(lane, win.data(), fetch, L[lane].re, L[lane].im);
            fft_dit<32, V>(L[lane].re, L[lane].im);
            apply_twiddle<V>(lane, tw, L[lane].re, L[lane].im);
            exchange_store<V>(lane, tile.data(), L[lane].re, L[lane].im);
        }
        for (int lane = 0; lane < 32; ++lane) {
            exchange_load<V>(lane, tile.data(), L[lane].re, L[lane].im);
            fft_dit<32, V>(L[lane].re, L[lane].im);
        }
        for (int lane = 0; lane < 32; ++lane) {
            const int src = (32 - lane) & 31;
            for (int k1 = 0; k1 < 16; ++k1) {
                const int preg = src == 0 ? ((32 - k1) & 31) : 31 - k1;  // what lane `src` provides
                tile[lane + 32 * k1] = frame_powers<V>(L[lane].re[k1], L[lane].im[k1], L[src].re[preg], L[src].im[preg]);
            }
            if (lane == 0) tile[512] = frame_powers<V>(L[0].re[16], L[0].im[16], L[0].re[16], L[0].im[16]);
        }
        for (int lane = 0; lane < 32; ++lane)
            for (int s = 0; s < bm.n_slots; ++s) {
                V aa, ab;
                mel_slot<V>(lane, tile.data(), bm.w.data() + (size_t)bm.wbase[s] * 32, bm.start[s * 32 + lane], bm.trip[s], aa,
                            ab);
                const int row = bm.row[s * 32 + lane];
                if (row < 0) continue;
                for (int t = 0; t < NV; ++t) {
                    const int fa = f0 + 2 * t;
                    if (fa < T) out[(size_t)fa * n_mels + row] = vget(aa, t);
                    if (fa + 1 < T) out[(size_t)(fa + 1) * n_mels + row] = vget(ab, t);
                }
            }
    }
}

static std::vector<float> read_f32(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f) / 4;
    fseek(f, 0, SEEK_SET);
    std::vector<float> v(n);
    if (n && fread(v.data(), 4, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}
static void write_f32(const char* path, const std::vector<float>& v) {
    FILE* f = fopen(path, "wb");
    if (!f) {
        perror(path);
        exit(2);
    }
    fwrite(v.data(), 4, v.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc >= 9 && !strcmp(argv[1], "logmel")) {
        const bool packed = !strcmp(argv[2], "packed");
        const int hop = atoi(argv[3]), n_mels = atoi(argv[4]);
        const double fmin = atof(argv[5]), fmax = atof(argv[6]);
        const std::vector<float> x = read_f32(argv[7]);
        std::vector<float> out;
        if (packed)
            run_logmel<f32x2>(x, hop, n_mels, fmin, fmax, out);
        else
            run_logmel<float>(x, hop, n_mels, fmin, fmax, out);
        write_f32(argv[8], out);
        return 0;
    }
    fprintf(stderr, "usage: host_check logmel <scalar|packed> <hop> <n_mels> <fmin> <fmax> <in.f32> <out.f32>\n");
    return 1;
}
