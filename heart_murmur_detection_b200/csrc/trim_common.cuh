// Silence-trim index kernel shared by the stand-alone trim (ragged_ops.cu) and the band-pass kernels
// that produce the hop energies on the fly (iir.cu).
#pragma once
#include <limits.h>
#include <stdint.h>

#include "hmfe_common.cuh"

namespace hmfe {

// Silence-trim indices from per-hop energy sums (frame_length == 2 * hop, centred frames):
// frame t covers hop blocks t-1 and t.  Same float32 arithmetic as trim_frame_power_kernel +
// trim_index_kernel (ragged_ops.cu); one CTA per clip.
struct TrimHopBatch {
    const float* hop_energy;
    const int64_t* clip_off;
    const int64_t* hop_off;
    int64_t* start_end;
    int64_t n_clips;
    int hop;
    float top_db;
};

static __global__ void __launch_bounds__(256) trim_index_hop_kernel(const TrimHopBatch b) {
    __shared__ float s_f[8];
    __shared__ int s_i[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float frame_length = (float)(2 * b.hop);
    for (int64_t clip = blockIdx.x; clip < b.n_clips; clip += gridDim.x) {
        const int n = (int)(b.clip_off[clip + 1] - b.clip_off[clip]);
        const int T = 1 + n / b.hop;
        const int H = (int)(b.hop_off[clip + 1] - b.hop_off[clip]);
        const float* e = b.hop_energy + b.hop_off[clip];
        auto power = [&](int t) {
            const float acc = (t >= 1 ? e[t - 1] : 0.0f) + (t < H ? e[t] : 0.0f);
            const float rms = sqrtf(acc / frame_length);
            return rms * rms;
        };
        float m = 0.0f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) m = fmaxf(m, power(t));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) s_f[warp] = m;
        __syncthreads();
        m = s_f[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_f[w]);
        const float amin2 = 1e-10f;
        const float ref_db = (float)(10.0 * log10(fmax(1e-10, (double)m)));
        int first = INT_MAX, last = -1;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const float db = 10.0f * log10f(fmaxf(amin2, power(t))) - ref_db;
            if (db > -b.top_db) {
                first = min(first, t);
                last = max(last, t);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
        }
        if (lane == 0) {
            s_i[warp] = first;
            s_i[8 + warp] = last;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                first = min(first, s_i[w]);
                last = max(last, s_i[8 + w]);
            }
            int64_t st = 0, en = 0;
            if (last >= 0) {
                st = (int64_t)first * b.hop;
                en = min((int64_t)n, (int64_t)(last + 1) * b.hop);
            }
            b.start_end[2 * clip] = st;
            b.start_end[2 * clip + 1] = en;
        }
        __syncthreads();
    }
}

}  // namespace hmfe
