// Log-mel power for frame lengths other than 1024: pre_process_audio_mel_t (/root/reference/src/util.py:481-492)
// takes `nfft` as an argument, and although every caller in the reference leaves it at 1024 the mirror accepts what the
// signature accepts (powers of two, 64 .. 4096).  This is the plain form of the computation - one warp per frame, the
// radix-2 FFT in shared memory, the mel bands as dot products over their non-zero range - not the register-resident
// 32 x 32 kernel of logmel.cu, which stays specialised for the frame length the reference uses.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "api_common.h"
#include "logmel_batch.cuh"
#include "tables.h"

namespace hmfe {

constexpr int kGenWarps = 4;

// One warp per frame.  Shared memory per warp: n_fft float2 (the transform) followed by n_fft / 2 + 1 floats (power).
__global__ void __launch_bounds__(kGenWarps * 32)
logmel_generic_kernel(const LogmelBatch b, const GenericTables tb, int n_fft, int log2n, int n_mels, int64_t n_frames) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_bins = n_fft / 2 + 1;
    const size_t per_warp = (size_t)n_fft * sizeof(float2) + (((size_t)n_bins * sizeof(float) + 15) & ~(size_t)15);
    float2* X = reinterpret_cast<float2*>(gen_smem + warp * per_warp);
    float* P = reinterpret_cast<float*>(X + n_fft);
    int64_t clip = 0;
    for (int64_t f = (int64_t)blockIdx.x * kGenWarps + warp; f < n_frames; f += (int64_t)gridDim.x * kGenWarps) {
        // frame -> (clip, frame in clip)
        const float* x;
        int nsamp, t;
        if (b.uniform_items > 0) {
            clip = f / b.uniform_T;
            t = (int)(f - clip * b.uniform_T);
            nsamp = b.uniform_n;
            x = b.wav + clip * (int64_t)nsamp;
        } else {
            if (!(f >= b.frame_off[clip] && f < b.frame_off[clip + 1])) {
                int64_t lo = 0, hi = b.n_clips;  // largest clip with frame_off[clip] <= f
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (b.frame_off[mid] <= f)
                        lo = mid;
                    else
                        hi = mid;
                }
                clip = lo;
            }
            t = (int)(f - b.frame_off[clip]);
            nsamp = (int)b.clip_len[clip];
            const int64_t s0 = b.clip_start[clip];
            x = s0 >= 0 ? b.wav + s0 : b.wav_alt + (-s0 - 1);
        }
        // centred frame, zero padding, Hann window, bit-reversed order
        const int base = t * b.hop - n_fft / 2;
        for (int n = lane; n < n_fft; n += 32) {
            const int i = base + n;
            const float v = (i >= 0 && i < nsamp) ? __ldg(x + i) * tb.win[n] : 0.0f;
            X[__brev((unsigned)n) >> (32 - log2n)] = make_float2(v, 0.0f);
        }
        __syncwarp();
        for (int s = 1; s <= log2n; ++s) {  // radix-2 decimation in time
            const int half = 1 << (s - 1), tw_step = n_fft >> s;
            for (int j = lane; j < n_fft / 2; j += 32) {
                const int pos = j & (half - 1), i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
                const float2 w = tb.tw[pos * tw_step];  // exp(-2 pi i pos / 2^s)
                const float2 a = X[i0], c = X[i1];
                const float tr = fmaf(c.x, w.x, -c.y * w.y), ti = fmaf(c.x, w.y, c.y * w.x);
                X[i0] = make_float2(a.x + tr, a.y + ti);
                X[i1] = make_float2(a.x - tr, a.y - ti);
            }
            __syncwarp();
        }
        for (int k = lane; k < n_bins; k += 32) P[k] = fmaf(X[k].x, X[k].x, X[k].y * X[k].y);
        __syncwarp();
        float vmax = 0.0f, vmin = INFINITY;
        float* o = b.out + f * n_mels;
        for (int m = lane; m < n_mels; m += 32) {
            const float* wrow = tb.mel + (size_t)m * n_bins;
            float acc = 0.0f;
            for (int k = tb.lo[m]; k < tb.hi[m]; ++k) acc = fmaf(__ldg(wrow + k), P[k], acc);
            o[m] = acc;
            vmax = fmaxf(vmax, acc);
            vmin = fminf(vmin, acc);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
        }
        if (lane == 0) {
            atomicMax(b.stats + 2 * clip, __float_as_uint(vmax));
            atomicMin(b.stats + 2 * clip + 1, __float_as_uint(vmin));
        }
        __syncwarp();
    }
}

bool generic_shape_ok(int n_fft) { return n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0; }

// host tables: full periodic Hann, exp(-2 pi i k / n) for k < n / 2, non-zero range of every mel row
void generic_tables(int n_fft, int n_mels, const std::vector<float>& mel_dense, std::vector<float>& win, std::vector<float>& tw,
                    std::vector<int>& lo, std::vector<int>& hi) {
    const double pi = 3.14159265358979323846;
    win.resize(n_fft);
    for (int n = 0; n < n_fft; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / n_fft));
    tw.resize(n_fft);
    for (int k = 0; k < n_fft / 2; ++k) {
        tw[2 * k] = (float)cos(2.0 * pi * k / n_fft);
        tw[2 * k + 1] = (float)(-sin(2.0 * pi * k / n_fft));
    }
    const int n_bins = n_fft / 2 + 1;
    lo.assign(n_mels, 0);
    hi.assign(n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
        int a = n_bins, z = 0;
        for (int k = 0; k < n_bins; ++k)
            if (mel_dense[(size_t)m * n_bins + k] != 0.0f) {
                a = std::min(a, k);
                z = k + 1;
            }
        lo[m] = a < z ? a : 0;
        hi[m] = a < z ? z : 0;
    }
}

int launch_logmel_generic(hmfe_logmel_plan* p, const LogmelBatch& b, int64_t n_frames, cudaStream_t st) {
    const int n_bins = p->n_fft / 2 + 1;
    const size_t per_warp = (size_t)p->n_fft * sizeof(float2) + (((size_t)n_bins * sizeof(float) + 15) & ~(size_t)15);
    const size_t smem = per_warp * kGenWarps;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(logmel_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int log2n = 0;
    while ((1 << log2n) < p->n_fft) ++log2n;
    const int64_t want = (n_frames + kGenWarps - 1) / kGenWarps;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count * 8));
    GenericTables tb{p->d_win, p->d_tw, p->d_gen_mel, p->d_gen_lo, p->d_gen_hi};
    logmel_generic_kernel<<<grid, kGenWarps * 32, smem, st>>>(b, tb, p->n_fft, log2n, p->n_mels, n_frames);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

}  // namespace hmfe
