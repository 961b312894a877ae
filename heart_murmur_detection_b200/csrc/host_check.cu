// CPU emulation of the warp algorithm in logmel.cu / kaldi_fbank.cu.
//
// The build container has no GPU.  The lane-level phases (fft_core.cuh, logmel_core.cuh)
// are __host__ __device__ templates; this program runs exactly the same arithmetic with a
// loop over the 32 lanes standing in for the warp, so index math (bit reversal, 32x32
// split, twiddle planes, exchange tile, partner lanes, banded mel slots) and float32
// accuracy are validated against the oracle before any GPU time is spent
// (tests/test_host_emulation.py).  Test infrastructure, not part of libhmfe.so.
//
// usage: host_check logmel <scalar|packed|pair> <hop> <n_mels> <fmin> <fmax> <in.f32> <out.f32>
//        host_check fbank <n_mels> <in.f32> <out.f32>
//        host_check banded
//        host_check hear <n_samples> <n_padded> <out_rows> <audio.f32> <window.f32> <mel.f32> <out_mel.f32> <out_pcen.f32>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <math.h>

#include "hear_core.cuh"
#include "logmel_core.cuh"
#include "tables.h"

using namespace hmfe;

template <typename V>
static void run_logmel(const std::vector<float>& x, int hop, int n_mels, double fmin, double fmax, std::vector<float>& out) {
    constexpr int NV = lanes_of<V>::value, FR = 2 * NV;
    const int nsamp = (int)x.size();
    const int T = 1 + nsamp / hop;
    const std::vector<float> dense = mel_filterbank_slaney(16000, kNfft, n_mels, fmin, fmax);
    const BandedMel bm = build_banded(dense, n_mels, kNfft / 2 + 1, lanes_of<V>::value == 2 ? 8 : 16, kBinsPad);
    if (!verify_banded(bm, dense, kBinsPad)) {
        fprintf(stderr, "banded mel verification failed\n");
        exit(3);
    }
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

// Mirrors logmel_power_pair_kernel (logmel.cu) lane for lane: one complex transform per item, elements paired.
static void run_logmel_pair(const std::vector<float>& x, int hop, int n_mels, double fmin, double fmax, std::vector<float>& out) {
    const int nsamp = (int)x.size();
    const int T = 1 + nsamp / hop;
    const std::vector<float> dense = mel_filterbank_slaney(16000, kNfft, n_mels, fmin, fmax);
    const BandedMel bm = build_banded(dense, n_mels, kNfft / 2 + 1, 16, kBinsPad);
    if (!verify_banded(bm, dense, kBinsPad)) {
        fprintf(stderr, "banded mel verification failed\n");
        exit(3);
    }
    const std::vector<float> win = half_hann_periodic(kNfft);
    const std::vector<float> twv = twiddle_plane_paired(kNfft);
    const float4* tw = reinterpret_cast<const float4*>(twv.data());
    out.assign((size_t)T * n_mels, 0.0f);
    std::vector<float> planes(2 * kPlane);
    float* pre = planes.data();
    float* pim = pre + kPlane;
    std::vector<f32x2> ptile(kBinsPad);
    struct Lane {
        f32x2 re[16], im[16];
    };
    std::vector<Lane> L(32);
    for (int f0 = 0; f0 < T; f0 += 2) {
        for (int lane = 0; lane < 32; ++lane) {
            auto fetch = [&](bool second, int n) -> float {
                const int f = f0 + (second ? 1 : 0);
                if (f >= T) return 0.0f;
                const int i = f * hop - kNfft / 2 + n;
                return (i >= 0 && i < nsamp) ? x[i] : 0.0f;
            };
            pair_load_window(lane, win.data(), fetch, L[lane].re, L[lane].im);
            fft32_paired(L[lane].re, L[lane].im);
            pair_twiddle(lane, tw, L[lane].re, L[lane].im);
            pair_exchange_store(lane, pre, pim, L[lane].re, L[lane].im);
        }
        for (int lane = 0; lane < 32; ++lane) {
            pair_exchange_load(lane, pre, pim, L[lane].re, L[lane].im);
            fft32_paired(L[lane].re, L[lane].im);
        }
        for (auto& e : ptile) e = f32x2{0.0f, 0.0f};
        for (int lane = 0; lane < 32; ++lane) {
            const int src = (32 - lane) & 31;
            for (int k1 = 0; k1 < 16; ++k1) {
                const int preg = src == 0 ? ((32 - k1) & 31) : 31 - k1;  // what lane `src` provides
                const xelem<float> pw = frame_powers<float>(L[lane].re[k1].x, L[lane].im[k1].x, pair_get(L[src].re, preg),
                                                            pair_get(L[src].im, preg));
                ptile[lane + 32 * k1] = f32x2{pw.a, pw.b};
            }
            if (lane == 0) {
                const xelem<float> pw = frame_powers<float>(L[0].re[0].y, L[0].im[0].y, L[0].re[0].y, L[0].im[0].y);
                ptile[512] = f32x2{pw.a, pw.b};
            }
        }
        for (int lane = 0; lane < 32; ++lane)
            for (int s = 0; s < bm.n_slots; ++s) {
                f32x2 acc;
                mel_slot_ab(lane, ptile.data(), bm.w.data() + (size_t)bm.wbase[s] * 32, bm.start[s * 32 + lane], bm.trip[s], acc);
                const int row = bm.row[s * 32 + lane];
                if (row < 0) continue;
                if (f0 < T) out[(size_t)f0 * n_mels + row] = acc.x;
                if (f0 + 1 < T) out[(size_t)(f0 + 1) * n_mels + row] = acc.y;
            }
    }
}

// Mirrors fbank_kernel (kaldi_fbank.cu) lane for lane.
static void run_fbank(const std::vector<float>& x, int n_mels, std::vector<float>& out) {
    const int win = 400, shift = 160, nsamp = (int)x.size();
    const float preemph = 0.97f;
    const int m = nsamp >= win ? 1 + (nsamp - win) / shift : 0;
    const std::vector<float> dense = mel_banks_kaldi(n_mels, 512, 16000.0, 20.0, 0.0);
    const BandedMel bm = build_banded(dense, n_mels, 257, 8, 320, 16);  // kFbMelGroup, prefer 16 as in kaldi_fbank.cu
    if (!verify_banded(bm, dense, 320)) {
        fprintf(stderr, "banded mel verification failed\n");
        exit(3);
    }
    const std::vector<float> w = half_hann_symmetric(win);
    const std::vector<float> twv = twiddle_plane(512, 16);
    const float2* tw = reinterpret_cast<const float2*>(twv.data());
    out.assign((size_t)m * n_mels, 0.0f);
    const int PS = 320;
    std::vector<xelem<float>> tile(32 * kXStride);
    struct Lane {
        f32x2 re[16], im[16];
        float zr[32], zi[32];
    };
    std::vector<Lane> L(32);
    for (int f0 = 0; f0 < m; f0 += 4) {
        float mean[4];
        for (int t = 0; t < 4; ++t) {
            float acc = 0.0f;
            if (f0 + t < m)
                for (int n = 0; n < win; ++n) acc += x[(size_t)(f0 + t) * shift + n];
            mean[t] = acc / (float)win;
        }
        for (int lane = 0; lane < 32; ++lane) {
            Lane& l = L[lane];
            for (int n2 = 0; n2 < 16; ++n2) {
                const int n = lane + 32 * n2;
                float y[4];
                for (int t = 0; t < 4; ++t) {
                    float v = 0.0f;
                    if (f0 + t < m && n < win) {
                        const float* xf = x.data() + (size_t)(f0 + t) * shift;
                        const float a = xf[n] - mean[t], pv = xf[n > 0 ? n - 1 : 0] - mean[t];
                        v = (a - preemph * pv) * w[n];
                    }
                    y[t] = v;
                }
                l.re[brev(n2, 4)] = f32x2{y[0], y[2]};
                l.im[brev(n2, 4)] = f32x2{y[1], y[3]};
            }
            fft_dit<16, f32x2>(l.re, l.im);
            for (int k2 = 1; k2 < 16; ++k2) {
                const float2 t2 = tw[k2 * 32 + lane];
                const f32x2 nr = vfnmas(l.im[k2], t2.y, vmuls(l.re[k2], t2.x));
                const f32x2 ni = vfmas(l.im[k2], t2.x, vmuls(l.re[k2], t2.y));
                l.re[k2] = nr;
                l.im[k2] = ni;
            }
            for (int k2 = 0; k2 < 16; ++k2) {
                tile[k2 * kXStride + lane] = xelem<float>{l.re[k2].x, l.im[k2].x};
                tile[(16 + k2) * kXStride + lane] = xelem<float>{l.re[k2].y, l.im[k2].y};
            }
        }
        for (int lane = 0; lane < 32; ++lane) {
            Lane& l = L[lane];
            for (int n1 = 0; n1 < 32; ++n1) {
                const xelem<float> e = tile[lane * kXStride + n1];
                l.zr[brev(n1, 5)] = e.a;
                l.zi[brev(n1, 5)] = e.b;
            }
            fft_dit<32, float>(l.zr, l.zi);
        }
        for (int lane = 0; lane < 32; ++lane) {
            const int k2 = lane & 15, src = (lane & 16) | ((16 - k2) & 15);
            xelem<float>* pt = tile.data() + (lane >> 4) * PS;
            const int src_k2 = src & 15;
            for (int k1 = 0; k1 < 16; ++k1) {
                const int preg = src_k2 == 0 ? ((32 - k1) & 31) : 31 - k1;
                pt[16 * k1 + k2] = frame_powers<float>(L[lane].zr[k1], L[lane].zi[k1], L[src].zr[preg], L[src].zi[preg]);
            }
            if (k2 == 0) pt[256] = frame_powers<float>(L[lane].zr[16], L[lane].zi[16], L[lane].zr[16], L[lane].zi[16]);
        }
        for (int lane = 0; lane < 32; ++lane)
            for (int s = 0; s < bm.n_slots; ++s) {
                const int start = bm.start[s * 32 + lane], row = bm.row[s * 32 + lane];
                float a[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
                for (int i = 0; i < bm.trip[s]; i += 2) {
                    const float w0 = bm.w[((size_t)bm.wbase[s] + i) * 32 + lane], w1 = bm.w[((size_t)bm.wbase[s] + i + 1) * 32 + lane];
                    const xelem<float> ea = tile[start + i], eb = tile[PS + start + i];
                    const xelem<float> fa = tile[start + i + 1], fb = tile[PS + start + i + 1];
                    a[0] = fmaf(ea.a, w0, a[0]);
                    a[1] = fmaf(ea.b, w0, a[1]);
                    a[2] = fmaf(eb.a, w0, a[2]);
                    a[3] = fmaf(eb.b, w0, a[3]);
                    c[0] = fmaf(fa.a, w1, c[0]);
                    c[1] = fmaf(fa.b, w1, c[1]);
                    c[2] = fmaf(fb.a, w1, c[2]);
                    c[3] = fmaf(fb.b, w1, c[3]);
                }
                for (int t = 0; t < 4; ++t) a[t] += c[t];
                if (row < 0) continue;
                for (int t = 0; t < 4; ++t)
                    if (f0 + t < m) out[(size_t)(f0 + t) * n_mels + row] = logf(fmaxf(a[t], 1.1920929e-7f));
            }
    }
}

// Mirrors hear_mel_kernel + hear_pcen_resize_kernel (hear_pcen.cu) lane for lane.  audio = [n_clips][n_samples],
// window = [400], mel = [201][n_mels] (the reference's layout).
static void run_hear(const std::vector<float>& audio, int n_samples, int n_padded, int out_rows, const std::vector<float>& window,
                     const std::vector<float>& mel, std::vector<float>& out_mel, std::vector<float>& out_pcen) {
    const int n_mels = (int)(mel.size() / kHearBins);
    const int n_clips = n_samples > 0 ? (int)(audio.size() / n_samples) : 1;
    const int T = (n_padded + kHearShift - 1) / kHearShift;
    std::vector<float> dense((size_t)n_mels * kHearBins);
    for (int k = 0; k < kHearBins; ++k)
        for (int m = 0; m < n_mels; ++m) dense[(size_t)m * kHearBins + k] = mel[(size_t)k * n_mels + m];
    const BandedMel bm = build_banded(dense, n_mels, kHearBins, 8, kHearPRows);
    if (!verify_banded(bm, dense, kHearPRows)) {
        fprintf(stderr, "banded mel verification failed\n");
        exit(3);
    }
    std::vector<float> win(kHearN);
    for (int n = 0; n < kHearN; ++n) win[n] = 0.5f * window[n];
    std::vector<float2> plane(25 * 16);
    for (int k1 = 0; k1 < 25; ++k1)
        for (int n2 = 0; n2 < 16; ++n2) {
            const double a = 2.0 * kPi * (double)((n2 * k1) % kHearN) / (double)kHearN;
            plane[k1 * 16 + n2] = make_float2((float)cos(a), (float)-sin(a));
        }
    float mn = n_samples < n_padded ? 0.0f : INFINITY, mx = n_samples < n_padded ? 0.0f : -INFINITY;
    for (float v : audio) {
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    const HearScale sc = hear_make_scale(mn, mx);
    out_mel.assign((size_t)n_clips * T * n_mels, 0.0f);
    std::vector<float2> tile(2 * kHearXTile);
    std::vector<xelem<f32x2>> ptile(kHearPRows);
    float* pf = reinterpret_cast<float*>(ptile.data());
    struct Z {
        float zr[16], zi[16];
    };
    std::vector<Z> L(32);
    for (int clip = 0; clip < n_clips; ++clip) {
        const float* x = audio.data() + (size_t)clip * n_samples;
        for (int f0 = 0; f0 < T; f0 += 4) {
            auto fetch = [&](int tr, bool second, int n) -> float {
                const int t = 2 * tr + (second ? 1 : 0);
                if (f0 + t >= T) return 0.0f;
                const int idx = (f0 + t) * kHearShift + n;
                return idx < n_samples ? hear_scale(sc, x[idx]) : (idx < n_padded ? hear_scale(sc, 0.0f) : 0.0f);
            };
            for (int lane = 0; lane < 32; ++lane) hear_pass1(lane, win.data(), plane.data(), fetch, tile.data());
            for (auto& e : ptile) e = xelem<f32x2>{};
            for (int r = 0; r < 2; ++r) {
                for (int lane = 0; lane < 25; ++lane) hear_pass2(lane, tile.data(), r, L[lane].zr, L[lane].zi);
                for (int lane = 0; lane < 25; ++lane) {
                    const int src = (25 - lane) % 25;
                    for (int k2 = 0; k2 < 8; ++k2) {
                        const int preg = hear_give_reg(src == 0, k2);
                        const xelem<float> pw = frame_powers<float>(L[lane].zr[k2], L[lane].zi[k2], L[src].zr[preg], L[src].zi[preg]);
                        const int k = lane + 25 * k2;
                        pf[4 * k + 2 * r] = pw.a;
                        pf[4 * k + 2 * r + 1] = pw.b;
                    }
                }
                const xelem<float> pw = frame_powers<float>(L[0].zr[8], L[0].zi[8], L[0].zr[8], L[0].zi[8]);
                pf[4 * 200 + 2 * r] = pw.a;
                pf[4 * 200 + 2 * r + 1] = pw.b;
            }
            float* o = out_mel.data() + ((size_t)clip * T + f0) * n_mels;
            for (int lane = 0; lane < 32; ++lane)
                for (int s = 0; s < bm.n_slots; ++s) {
                    f32x2 aa, ab;
                    mel_slot<f32x2>(lane, ptile.data(), bm.w.data() + (size_t)bm.wbase[s] * 32, bm.start[s * 32 + lane], bm.trip[s],
                                    aa, ab);
                    const int row = bm.row[s * 32 + lane];
                    if (row < 0) continue;
                    if (f0 < T) o[row] = aa.x;
                    if (f0 + 1 < T) o[n_mels + row] = aa.y;
                    if (f0 + 2 < T) o[2 * n_mels + row] = ab.x;
                    if (f0 + 3 < T) o[3 * n_mels + row] = ab.y;
                }
        }
    }
    PcenParams pp;
    pp.alpha = 0.8f;
    pp.c_in = 0.04f;
    pp.c_state = (float)(1.0 - 0.04);
    pp.delta = 2.0f;
    pp.inv_root = 0.5f;
    pp.floor = 1e-8f;
    pp.delta_root = powf(2.0f, 0.5f);
    out_pcen.assign((size_t)n_clips * out_rows * n_mels, 0.0f);
    for (int clip = 0; clip < n_clips; ++clip)
        for (int c = 0; c < n_mels; ++c) {
            const float* xm = out_mel.data() + (size_t)clip * T * n_mels + c;
            float* o = out_pcen.data() + (size_t)clip * out_rows * n_mels + c;
            hear_pcen_column(pp, T, out_rows, [&](int t) { return xm[(size_t)t * n_mels]; },
                             [&](int i, float v) { o[(size_t)i * n_mels] = v; });
        }
}

// Shared-memory wavefronts of the banded mel tile reads for 8-byte tile elements (a wavefront serves 16 lanes whose
// element indices are distinct modulo 16, or equal): the quantity the lane assignment of build_banded minimises.
static long banded_wavefronts8(const BandedMel& bm) {
    long wf = 0;
    for (int s = 0; s < bm.n_slots; ++s)
        for (int i = 0; i < bm.trip[s]; ++i)
            for (int h = 0; h < 2; ++h) {
                std::vector<int> seen[16];
                for (int l = 16 * h; l < 16 * h + 16; ++l) {
                    const int a = bm.start[s * 32 + l] + i;
                    std::vector<int>& v = seen[a & 15];
                    bool dup = false;
                    for (int e : v) dup = dup || e == a;
                    if (!dup) v.push_back(a);
                }
                size_t mx = 1;
                for (int r = 0; r < 16; ++r) mx = seen[r].size() > mx ? seen[r].size() : mx;
                wf += (long)mx;
            }
    return wf;
}

// host_check banded: the Kaldi mel bank (128 bands, 257 bins) under the three placements of kaldi_fbank.cu
static int run_banded() {
    const std::vector<float> dense = mel_banks_kaldi(128, 512, 16000.0, 20.0, 0.0);
    const int cfg[3][2] = {{16, 0}, {8, 0}, {8, 16}};
    for (const auto& c : cfg) {
        const BandedMel bm = build_banded(dense, 128, 257, c[0], 320, c[1]);
        printf("group %d prefer %d total_trip %d wavefronts %ld ideal %d verify %d\n", c[0], c[1], bm.total_trip,
               banded_wavefronts8(bm), 2 * bm.total_trip, (int)verify_banded(bm, dense, 320));
    }
    // the assignment solver on a matrix with a known optimum (a permutation of zeros in a field of ones)
    std::vector<std::vector<int>> cost(32, std::vector<int>(32, 1));
    for (int i = 0; i < 32; ++i) cost[i][(7 * i + 3) % 32] = 0;
    const std::vector<int> col = assign_min_cost(cost);
    int total = 0;
    std::vector<char> used(32, 0);
    for (int i = 0; i < 32; ++i) {
        total += cost[i][col[i]];
        used[col[i]] = 1;
    }
    int distinct = 0;
    for (char u : used) distinct += u;
    printf("assign total %d distinct %d\n", total, distinct);
    return 0;
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
        if (!strcmp(argv[2], "pair"))
            run_logmel_pair(x, hop, n_mels, fmin, fmax, out);
        else if (packed)
            run_logmel<f32x2>(x, hop, n_mels, fmin, fmax, out);
        else
            run_logmel<float>(x, hop, n_mels, fmin, fmax, out);
        write_f32(argv[8], out);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "fbank")) {
        const std::vector<float> x = read_f32(argv[3]);
        std::vector<float> out;
        run_fbank(x, atoi(argv[2]), out);
        write_f32(argv[4], out);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "banded")) return run_banded();
    if (argc >= 10 && !strcmp(argv[1], "hear")) {
        std::vector<float> om, op;
        run_hear(read_f32(argv[5]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), read_f32(argv[6]), read_f32(argv[7]), om, op);
        write_f32(argv[8], om);
        write_f32(argv[9], op);
        return 0;
    }
    fprintf(stderr, "usage: host_check logmel <scalar|packed|pair> <hop> <n_mels> <fmin> <fmax> <in.f32> <out.f32>\n");
    return 1;
}
