// Fused framing + Hann window + 1024-point real FFT (two real frames per complex FFT)
// + power + banded mel projection, then a per-clip dB / min-max epilogue.
// Replaces pre_process_audio_mel_t (/root/reference/src/util.py:481-501).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "api_common.h"
#include "logmel_batch.cuh"
#include "tables.h"

namespace hmfe {

// raw[t][h][n2] = sample (lane + 32*n2) of frame f0 + 2t + h (zero outside the clip / beyond T)
template <int NV, bool REFLECT = false>
HMFE_D void load_raw(const ItemCtx& c, const LogmelBatch& b, int lane, float (&raw)[NV][2][32]) {
    const int hop = b.hop;
    int base[NV][2];
    bool interior = c.valid;
#pragma unroll
    for (int t = 0; t < NV; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int f = c.f0 + 2 * t + h;
            base[t][h] = f * hop - kNfft / 2;
            interior = interior && f < c.T && base[t][h] >= 0 && base[t][h] + kNfft <= c.nsamp;
            if (f >= c.T) base[t][h] = c.nsamp;  // every sample out of range -> zeros
        }
    if (interior) {
#pragma unroll
        for (int t = 0; t < NV; ++t)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* p = c.x + base[t][h] + lane;
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) raw[t][h][n2] = __ldg(p + 32 * n2);
            }
    } else {
#pragma unroll
        for (int t = 0; t < NV; ++t)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) {
                    int i = base[t][h] + lane + 32 * n2;
                    if constexpr (REFLECT) {  // clips are longer than n_fft / 2 (checked on the host): one fold per side
                        const bool frame_ok = c.f0 + 2 * t + h < c.T;  // base was moved out of range for missing frames
                        i = i < 0 ? -i : i;
                        i = i >= c.nsamp ? 2 * (c.nsamp - 1) - i : i;
                        raw[t][h][n2] = (c.valid && frame_ok && i >= 0 && i < c.nsamp) ? __ldg(c.x + i) : 0.0f;
                    } else {
                        raw[t][h][n2] = (c.valid && i >= 0 && i < c.nsamp) ? __ldg(c.x + i) : 0.0f;
                    }
                }
    }
}

// NSLOTS > 0: number of mel slots known at compile time (2 for 64 mels, 4 for 128); 0: runtime.
// REFLECT (centre padding mirrors the clip, HMFE_PAD_REFLECT) is a compile-time switch: folding it into the edge path at run
// time grew the kernel by 40 % (the edge path is unrolled 128 times, twice) and cost 5 % on c1 through the instruction cache.
template <typename V, int WARPS, int MINB, int NSLOTS, bool REFLECT = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
logmel_power_kernel(const LogmelBatch b, const LogmelTables tb, const MelMeta mm) {
    constexpr int NV = lanes_of<V>::value;
    constexpr int FR = 2 * NV;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_tw = reinterpret_cast<float2*>(smem);
    float* s_win = reinterpret_cast<float*>(s_tw + 1024);
    float* s_melw = s_win + 1024;
    int* s_start = reinterpret_cast<int*>(s_melw + mm.total_trip * 32);
    int* s_row = s_start + mm.n_slots * 32;
    int* s_meta = s_row + mm.n_slots * 32;  // trip[kMaxSlots], wbase[kMaxSlots]
    size_t tbytes = (size_t)(1024 * 8 + 1024 * 4 + mm.total_trip * 128 + mm.n_slots * 256 + 2 * kMaxSlots * 4);
    tbytes = (tbytes + 15) & ~(size_t)15;
    xelem<V>* tile = reinterpret_cast<xelem<V>*>(smem + tbytes) + (threadIdx.x >> 5) * kTileElems;

    for (int i = threadIdx.x; i < 1024; i += WARPS * 32) {
        s_tw[i] = tb.tw[i];
        s_win[i] = tb.win[i];
    }
    for (int i = threadIdx.x; i < mm.total_trip * 32; i += WARPS * 32) s_melw[i] = tb.melw[i];
    for (int i = threadIdx.x; i < mm.n_slots * 32; i += WARPS * 32) {
        s_start[i] = tb.start[i];
        s_row[i] = tb.row[i];
    }
    if (threadIdx.x < kMaxSlots) {
        s_meta[threadIdx.x] = mm.trip[threadIdx.x];
        s_meta[kMaxSlots + threadIdx.x] = mm.wbase[threadIdx.x];
    }
    // The exchange never writes the padding column (index 33*a + 32) of the tile, but the mel
    // windows may read it under a zero weight: clear it once so stale shared memory (NaN bit
    // patterns left by other kernels) cannot turn 0 * x into NaN.
    tile[(threadIdx.x & 31) * kXStride + 32] = xelem<V>{};
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mels = mm.n_mels;
    const int n_slots = NSLOTS > 0 ? NSLOTS : mm.n_slots;
    // Work items are claimed in blocks of kItemBlock consecutive items per warp from a global counter:
    // warps that start late (SMs shared with a concurrent collective kernel) or hit long clips simply
    // claim fewer blocks, instead of holding a fixed 1/grid share of the batch.
    constexpr int kItemBlock = 8;
    const int64_t it_end = item_count(b);
    auto claim = [&]() -> int64_t {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(b.queue, (unsigned long long)kItemBlock);
        return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    int64_t blk_end = 0;
    auto next_item = [&](int64_t item) -> int64_t {  // item following `item` for this warp
        if (item + 1 < blk_end) return item + 1;
        if (item >= it_end) return item;  // drained: no further claims
        const int64_t nb = claim();
        blk_end = nb + kItemBlock;
        return nb;
    };
    int64_t clip_cursor = -1;

    // software pipeline: the samples of item i+WARPS are fetched into registers while the mel
    // projection of item i runs (the FFT registers are dead by then)
    float raw[NV][2][32];
    // Stagger the warps of the CTA: the phases of an item alternate between the FP32 pipe (FFT passes)
    // and the load/store pipe (exchange, mel); warps that start together stay in the same phase and
    // the two pipes take turns idling.
    if (b.stagger_ns > 0) __nanosleep((unsigned)(warp * b.stagger_ns));
    int64_t item = claim();
    blk_end = item + kItemBlock;
    ItemCtx cur = locate_item<FR>(b, n_mels, item, it_end, clip_cursor);
    load_raw<NV, REFLECT>(cur, b, lane, raw);

    while (item < it_end) {
        V re[32], im[32];
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) {
            const float w = s_win[lane + 32 * n2];
            V xa, xb;
#pragma unroll
            for (int t = 0; t < NV; ++t) {
                vput(xa, t, raw[t][0][n2]);
                vput(xb, t, raw[t][1][n2]);
            }
            re[brev(n2, 5)] = vmuls(xa, w);
            im[brev(n2, 5)] = vmuls(xb, w);
        }
        // both 32-point passes run the same unrolled butterfly code (one copy in the instruction
        // cache); the first pass is followed by the twiddle and the lane exchange
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            fft_dit<32, V>(re, im);
            if (pass == 0) {
                apply_twiddle<V>(lane, s_tw, re, im);
                exchange_store<V>(lane, tile, re, im);
                __syncwarp();
                exchange_load<V>(lane, tile, re, im);
                __syncwarp();
            }
        }

        {
            const int src = (32 - lane) & 31;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const V give_r = lane == 0 ? re[(32 - k1) & 31] : re[31 - k1];
                const V give_i = lane == 0 ? im[(32 - k1) & 31] : im[31 - k1];
                const V pr = shfl(give_r, src), pi = shfl(give_i, src);
                const xelem<V> pw = frame_powers<V>(re[k1], im[k1], pr, pi);
                store_halves(&tile[lane + 32 * k1], pw.a, pw.b);
            }
            if (lane == 0) tile[512] = frame_powers<V>(re[16], im[16], re[16], im[16]);
        }
        __syncwarp();

        item = next_item(item);
        const ItemCtx nxt = locate_item<FR>(b, n_mels, item, it_end, clip_cursor);
        load_raw<NV, REFLECT>(nxt, b, lane, raw);

        float vmax = 0.0f, vmin = INFINITY;
#pragma unroll
        for (int s = 0; s < (NSLOTS > 0 ? NSLOTS : kMaxSlots); ++s) {
            if (NSLOTS == 0 && s >= n_slots) break;
            V aa, ab;
            mel_slot<V>(lane, tile, s_melw + s_meta[kMaxSlots + s] * 32, s_start[s * 32 + lane], s_meta[s], aa, ab);
            const int row = s_row[s * 32 + lane];
            if (row >= 0) {
#pragma unroll
                for (int t = 0; t < NV; ++t) {
                    const int fa = cur.f0 + 2 * t;
                    if (fa < cur.T) {
                        const float v = vget(aa, t);
                        cur.o[(int64_t)fa * n_mels + row] = v;
                        vmax = fmaxf(vmax, v);
                        vmin = fminf(vmin, v);
                    }
                    if (fa + 1 < cur.T) {
                        const float v = vget(ab, t);
                        cur.o[(int64_t)(fa + 1) * n_mels + row] = v;
                        vmax = fmaxf(vmax, v);
                        vmin = fminf(vmin, v);
                    }
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
        }
        if (lane == 0) {
            atomicMax(b.stats + 2 * cur.clip, __float_as_uint(vmax));
            atomicMin(b.stats + 2 * cur.clip + 1, __float_as_uint(vmin));
        }
        __syncwarp();
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// "pair" variant (logmel_core.cuh): one complex transform (2 frames) per warp iteration with the
// packed FP32 instructions applied across elements.  64 registers of FFT data per thread instead
// of 128 and an 8.7 KB exchange area per warp instead of 16.9 KB: 20 warps per SM instead of 12.
// ---------------------------------------------------------------------------------------------
template <int WARPS, int MINB, int NSLOTS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
logmel_power_pair_kernel(const LogmelBatch b, const LogmelTables tb, const MelMeta mm) {
    constexpr int FR = 2;
    extern __shared__ __align__(16) unsigned char smem[];
    float4* s_tw = reinterpret_cast<float4*>(smem);  // [16][32]
    float* s_win = reinterpret_cast<float*>(s_tw + 512);
    float* s_melw = s_win + 1024;
    int* s_start = reinterpret_cast<int*>(s_melw + mm.total_trip * 32);
    int* s_row = s_start + mm.n_slots * 32;
    int* s_meta = s_row + mm.n_slots * 32;  // trip[kMaxSlots], wbase[kMaxSlots]
    size_t tbytes = (size_t)(512 * 16 + 1024 * 4 + mm.total_trip * 128 + mm.n_slots * 256 + 2 * kMaxSlots * 4);
    tbytes = (tbytes + 15) & ~(size_t)15;
    float* pre = reinterpret_cast<float*>(smem + tbytes) + (threadIdx.x >> 5) * (2 * kPlane);
    float* pim = pre + kPlane;
    f32x2* ptile = reinterpret_cast<f32x2*>(pre);  // power tile (p_a, p_b) per bin, overlays the planes

    for (int i = threadIdx.x; i < 512; i += WARPS * 32) s_tw[i] = reinterpret_cast<const float4*>(tb.tw)[i];
    for (int i = threadIdx.x; i < 1024; i += WARPS * 32) s_win[i] = tb.win[i];
    for (int i = threadIdx.x; i < mm.total_trip * 32; i += WARPS * 32) s_melw[i] = tb.melw[i];
    for (int i = threadIdx.x; i < mm.n_slots * 32; i += WARPS * 32) {
        s_start[i] = tb.start[i];
        s_row[i] = tb.row[i];
    }
    if (threadIdx.x < kMaxSlots) {
        s_meta[threadIdx.x] = mm.trip[threadIdx.x];
        s_meta[kMaxSlots + threadIdx.x] = mm.wbase[threadIdx.x];
    }
    // the mel windows read up to kBinsPad power entries under zero weights: no stale NaN bit patterns
    for (int i = threadIdx.x & 31; i < 2 * kPlane; i += 32) pre[i] = 0.0f;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mels = mm.n_mels;
    const int n_slots = NSLOTS > 0 ? NSLOTS : mm.n_slots;
    constexpr int kItemBlock = 16;  // 16 items = 32 frames per claim, as in the packed kernel
    const int64_t it_end = item_count(b);
    auto claim = [&]() -> int64_t {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(b.queue, (unsigned long long)kItemBlock);
        return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    int64_t blk_end = 0;
    auto next_item = [&](int64_t item) -> int64_t {
        if (item + 1 < blk_end) return item + 1;
        if (item >= it_end) return item;
        const int64_t nb = claim();
        blk_end = nb + kItemBlock;
        return nb;
    };
    int64_t clip_cursor = -1;
    if (b.stagger_ns > 0) __nanosleep((unsigned)(warp * b.stagger_ns));

    float raw[1][2][32];
    int64_t item = claim();
    blk_end = item + kItemBlock;
    ItemCtx cur = locate_item<FR>(b, n_mels, item, it_end, clip_cursor);
    load_raw<1>(cur, b, lane, raw);

    while (item < it_end) {
        f32x2 re[16], im[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int n2 = 2 * brev(q, 4);
            const f32x2 w{s_win[lane + 32 * n2], s_win[lane + 32 * (n2 + 1)]};
            re[q] = vmul(f32x2{raw[0][0][n2], raw[0][0][n2 + 1]}, w);
            im[q] = vmul(f32x2{raw[0][1][n2], raw[0][1][n2 + 1]}, w);
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            fft32_paired(re, im);
            if (pass == 0) {
                pair_twiddle(lane, s_tw, re, im);
                pair_exchange_store(lane, pre, pim, re, im);
                __syncwarp();
                pair_exchange_load(lane, pre, pim, re, im);
                __syncwarp();
            }
        }
        {
            const int src = (32 - lane) & 31;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const float give_r = lane == 0 ? pair_get(re, (32 - k1) & 31) : pair_get(re, 31 - k1);
                const float give_i = lane == 0 ? pair_get(im, (32 - k1) & 31) : pair_get(im, 31 - k1);
                const float pr = __shfl_sync(0xffffffffu, give_r, src), pi = __shfl_sync(0xffffffffu, give_i, src);
                const xelem<float> pw = frame_powers<float>(re[k1].x, im[k1].x, pr, pi);
                ptile[lane + 32 * k1] = f32x2{pw.a, pw.b};
            }
            if (lane == 0) {
                const xelem<float> pw = frame_powers<float>(re[0].y, im[0].y, re[0].y, im[0].y);
                ptile[512] = f32x2{pw.a, pw.b};
            }
        }
        __syncwarp();

        item = next_item(item);
        const ItemCtx nxt = locate_item<FR>(b, n_mels, item, it_end, clip_cursor);
        load_raw<1>(nxt, b, lane, raw);

        float vmax = 0.0f, vmin = INFINITY;
#pragma unroll
        for (int s = 0; s < (NSLOTS > 0 ? NSLOTS : kMaxSlots); ++s) {
            if (NSLOTS == 0 && s >= n_slots) break;
            f32x2 acc;
            mel_slot_ab(lane, ptile, s_melw + s_meta[kMaxSlots + s] * 32, s_start[s * 32 + lane], s_meta[s], acc);
            const int row = s_row[s * 32 + lane];
            if (row >= 0) {
                if (cur.f0 < cur.T) {
                    cur.o[(int64_t)cur.f0 * n_mels + row] = acc.x;
                    vmax = fmaxf(vmax, acc.x);
                    vmin = fminf(vmin, acc.x);
                }
                if (cur.f0 + 1 < cur.T) {
                    cur.o[(int64_t)(cur.f0 + 1) * n_mels + row] = acc.y;
                    vmax = fmaxf(vmax, acc.y);
                    vmin = fminf(vmin, acc.y);
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
        }
        if (lane == 0 && cur.valid) {
            atomicMax(b.stats + 2 * cur.clip, __float_as_uint(vmax));
            atomicMin(b.stats + 2 * cur.clip + 1, __float_as_uint(vmin));
        }
        __syncwarp();
        cur = nxt;
    }
}

__global__ void logmel_init_stats_kernel(unsigned* stats, int64_t n_clips, unsigned long long* queue) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *queue = 0ull;
    if (i < n_clips) {
        stats[2 * i] = 0u;
        stats[2 * i + 1] = 0x7f800000u;  // +inf
    }
}

// dB re clip max, floor at max - top_db, optional clip-wide min-max normalisation, in place.
__global__ void __launch_bounds__(256)
logmel_finalize_kernel(const LogmelBatch b, int n_mels, int out_mode, float amin, float top_db) {
    for (int64_t clip = blockIdx.x; clip < b.n_clips; clip += gridDim.x) {
        float* o;
        int64_t count;
        if (b.uniform_items > 0) {
            o = b.out + clip * (int64_t)b.uniform_T * n_mels;
            count = (int64_t)b.uniform_T * n_mels;
        } else {
            const int64_t f0 = b.frame_off[clip];
            o = b.out + f0 * n_mels;
            count = (b.frame_off[clip + 1] - f0) * n_mels;
        }
        const float pmax = __uint_as_float(b.stats[2 * clip]);
        const float pmin = __uint_as_float(b.stats[2 * clip + 1]);
        // 10 log10(x) = (10 / log2 10) * log2(x) on the special-function unit: ~1e-5 dB from the
        // correctly rounded value (budget 1e-2 dB).  The reference term goes through the same
        // formula (product rounded, then subtracted, no contraction) so that the clip maximum maps to
        // exactly 0 dB and the normalised output spans exactly [0, 1]; an all-equal clip gives 0.
        const float k10 = 3.01029995663981195f;
        // lg2.approx.ftz: the argument is >= amin = 1e-10, a normal number, so no denormal pre-scaling is needed
        auto lg2 = [](float x) {
            float r;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
            return r;
        };
        // HMFE_LOGMEL_OUT_DB_ABS: power_to_db(ref=1.0, top_db=None) = 10 log10(max(amin, S)), no clip-wide terms
        const bool abs_db = out_mode == HMFE_LOGMEL_OUT_DB_ABS;
        const float ref_db = abs_db ? 0.0f : __fmul_rn(k10, lg2(fmaxf(amin, pmax)));
        auto to_db = [&](float p) { return __fsub_rn(__fmul_rn(k10, lg2(fmaxf(amin, p))), ref_db); };
        const float smax = to_db(pmax);
        const float floor_db = abs_db ? -INFINITY : smax - top_db;
        const float smin = fmaxf(to_db(pmin), floor_db);
        const bool normalise = out_mode == HMFE_LOGMEL_OUT_NORMALISED && smax != smin;
        const float denom = normalise ? smax - smin : 1.0f;
        const float sub = normalise ? smin : 0.0f;
        // (d - smin) / denom as a reciprocal multiply plus one residual correction: correctly rounded for
        // these operands (0 <= d - smin <= denom), in particular exactly 1 at the clip maximum
        const float inv = 1.0f / denom;
        auto norm = [&](float d) {
            const float n = d - sub;
            const float q = n * inv;
            return fmaf(fmaf(-q, denom, n), inv, q);
        };
        float4* o4 = reinterpret_cast<float4*>(o);
        const int64_t n4 = count >> 2;
        constexpr int U = 4;  // float4 loads in flight per thread
        for (int64_t i0 = threadIdx.x; i0 < n4; i0 += (int64_t)U * blockDim.x) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i < n4) v[u] = o4[i];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i >= n4) break;
                float* p = reinterpret_cast<float*>(&v[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = fmaxf(to_db(p[j]), floor_db);
                    p[j] = normalise ? norm(d) : d;
                }
                o4[i] = v[u];
            }
        }
    }
}

}  // namespace hmfe

using namespace hmfe;

namespace hmfe {  // logmel_generic.cu
bool generic_shape_ok(int n_fft);
void generic_tables(int n_fft, int n_mels, const std::vector<float>& mel_dense, std::vector<float>& win, std::vector<float>& tw,
                    std::vector<int>& lo, std::vector<int>& hi);
int launch_logmel_generic(hmfe_logmel_plan* p, const LogmelBatch& b, int64_t n_frames, cudaStream_t st);
}  // namespace hmfe

constexpr int kVariantGeneric = 5;  // internal: n_fft != 1024

namespace hmfe {  // logmel_tc.cu
std::vector<uint32_t> tc_build_a_words(const std::vector<float>& mel_dense, int n_mels, int n_bins);
bool tc_shape_ok(const hmfe_logmel_plan* p);
int launch_logmel_tc(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st);
}  // namespace hmfe

template <typename T>
static int upload_vec(const std::vector<T>& v, T** dptr) {
    HMFE_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), std::max<size_t>(1, v.size()) * sizeof(T)));
    HMFE_CHECK_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return HMFE_OK;
}

template <typename V, int WARPS, int MINB, int NSLOTS, bool REFLECT = false>
static int launch_power_n(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    const size_t smem = p->table_smem + (size_t)WARPS * kTileElems * sizeof(xelem<V>);
    auto kern = logmel_power_kernel<V, WARPS, MINB, NSLOTS, REFLECT>;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (b.n_items + WARPS - 1) / WARPS;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count * MINB));
    LogmelTables tb{p->d_win, p->d_tw, p->d_melw, p->d_start, p->d_row};
    kern<<<grid, WARPS * 32, smem, st>>>(b, tb, p->meta);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

template <typename V, int WARPS, int MINB>
static int launch_power(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    if (p->pad_mode == HMFE_PAD_REFLECT) {
        if constexpr (lanes_of<V>::value == 2) {  // reflect padding exists for the default (packed) variant only
            switch (p->meta.n_slots) {
                case 2: return launch_power_n<V, WARPS, MINB, 2, true>(p, b, st);
                default: return launch_power_n<V, WARPS, MINB, 0, true>(p, b, st);
            }
        }
    }
    switch (p->meta.n_slots) {
        case 2: return launch_power_n<V, WARPS, MINB, 2>(p, b, st);
        case 4: return launch_power_n<V, WARPS, MINB, 4>(p, b, st);
        default: return launch_power_n<V, WARPS, MINB, 0>(p, b, st);
    }
}

constexpr int kPairWarps = 20;

template <int NSLOTS>
static int launch_pair_n(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    // tables: float4 twiddles (8 KB) instead of float2 (8 KB): same size as the other variants' layout
    const size_t tbytes = ((size_t)(512 * 16 + 1024 * 4 + p->meta.total_trip * 128 + p->meta.n_slots * 256 + 2 * kMaxSlots * 4) + 15) &
                          ~(size_t)15;
    const size_t smem = tbytes + (size_t)kPairWarps * 2 * kPlane * sizeof(float);
    auto kern = logmel_power_pair_kernel<kPairWarps, 1, NSLOTS>;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (b.n_items + kPairWarps - 1) / kPairWarps;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count));
    LogmelTables tb{p->d_win, p->d_tw, p->d_melw, p->d_start, p->d_row};
    kern<<<grid, kPairWarps * 32, smem, st>>>(b, tb, p->meta);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

static int launch_pair(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    switch (p->meta.n_slots) {
        case 2: return launch_pair_n<2>(p, b, st);
        case 4: return launch_pair_n<4>(p, b, st);
        default: return launch_pair_n<0>(p, b, st);
    }
}

// stats / queue initialisation, the power kernel of the plan's variant, the dB / min-max epilogue
static int run_logmel(hmfe_logmel_plan* p, const LogmelBatch& b, int out_mode, cudaStream_t st) {
    const int64_t n_clips = b.n_clips;
    logmel_init_stats_kernel<<<(unsigned)((n_clips + 255) / 256), 256, 0, st>>>(b.stats, n_clips, b.queue);
    HMFE_CHECK_CUDA(cudaGetLastError());
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (p->profile) {
        for (int i = 0; i < 3; ++i) {
            HMFE_CHECK_CUDA(cudaEventCreate(&ev[i]));
            p->prof_events.push_back(ev[i]);
        }
        HMFE_CHECK_CUDA(cudaEventRecord(ev[0], st));
    }
    int rc = p->variant == kVariantGeneric       ? launch_logmel_generic(p, b, b.n_frames_total, st)
             : p->variant == HMFE_VARIANT_TC     ? launch_logmel_tc(p, b, st)
             : p->variant == HMFE_VARIANT_PACKED ? launch_power<f32x2, 12, 1>(p, b, st)
             : p->variant == HMFE_VARIANT_PAIR   ? launch_pair(p, b, st)
                                                 : launch_power<float, 8, 2>(p, b, st);
    if (rc != HMFE_OK) return rc;
    if (p->profile) HMFE_CHECK_CUDA(cudaEventRecord(ev[1], st));
    p->last_launches = 2;
    if (out_mode != HMFE_LOGMEL_OUT_POWER) {
        const int grid = (int)std::min<int64_t>(n_clips, (int64_t)p->sm_count * 8);
        logmel_finalize_kernel<<<grid, 256, 0, st>>>(b, p->n_mels, out_mode, 1e-10f, 80.0f);
        HMFE_CHECK_CUDA(cudaGetLastError());
        p->last_launches = 3;
    }
    if (p->profile) HMFE_CHECK_CUDA(cudaEventRecord(ev[2], st));
    return HMFE_OK;
}

extern "C" {

int hmfe_logmel_plan_create(hmfe_logmel_plan** plan, int sample_rate, int n_fft, int hop, int n_mels, double f_min,
                            double f_max, int variant) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    if (n_fft != kNfft && !generic_shape_ok(n_fft)) {
        set_error("n_fft=%d unsupported: powers of two from 64 to 4096 (the reference uses 1024, src/util.py:482)", n_fft);
        return HMFE_ERR_UNSUPPORTED;
    }
    HMFE_REQUIRE(hop >= 1 && hop <= 4096, "hop=%d out of range", hop);
    HMFE_REQUIRE(n_mels >= 4 && n_mels % 4 == 0 && n_mels <= 32 * kMaxSlots, "n_mels=%d must be a multiple of 4 <= %d", n_mels,
                 32 * kMaxSlots);
    const bool plain = n_fft != kNfft || n_mels % 32 != 0;  // the register-resident kernels: n_fft 1024, mel bands by 32
    HMFE_REQUIRE(sample_rate > 0 && f_min >= 0 && f_max > f_min && f_max <= 0.5 * sample_rate + 1e-9,
                 "bad frequency range [%g, %g] for sr=%d", f_min, f_max, sample_rate);
    HMFE_REQUIRE(variant >= 0 && variant <= 4, "bad variant %d", variant);
    hmfe_logmel_plan* p = new (std::nothrow) hmfe_logmel_plan();
    HMFE_REQUIRE(p != nullptr, "out of host memory");
    p->sample_rate = sample_rate;
    p->n_fft = n_fft;
    p->hop = hop;
    p->n_mels = n_mels;
    p->n_bins = n_fft / 2 + 1;
    p->f_min = f_min;
    p->f_max = f_max;
    p->variant = variant == HMFE_VARIANT_AUTO ? HMFE_VARIANT_PACKED : variant;
    p->sm_count = device_sm_count();
    p->mel_dense = mel_filterbank_slaney(sample_rate, n_fft, n_mels, f_min, f_max);
    if (plain) {  // the plain kernel of logmel_generic.cu, whatever variant was asked for
        p->variant = kVariantGeneric;
        std::vector<float> win, tw;
        std::vector<int> lo, hi;
        generic_tables(n_fft, n_mels, p->mel_dense, win, tw, lo, hi);
        int rc = upload_vec(win, &p->d_win);
        if (rc == HMFE_OK) rc = upload_vec(tw, reinterpret_cast<float**>(&p->d_tw));
        if (rc == HMFE_OK) rc = upload_vec(p->mel_dense, &p->d_gen_mel);
        if (rc == HMFE_OK) rc = upload_vec(lo, &p->d_gen_lo);
        if (rc == HMFE_OK) rc = upload_vec(hi, &p->d_gen_hi);
        if (rc != HMFE_OK) {
            hmfe_logmel_plan_destroy(p);
            return rc;
        }
        *plan = p;
        return HMFE_OK;
    }
    const int group = (p->variant == HMFE_VARIANT_PACKED || p->variant == HMFE_VARIANT_TC) ? 8 : 16;  // lanes per shared-memory phase (16 B / 8 B elements)
    const BandedMel bm = build_banded(p->mel_dense, n_mels, p->n_bins, group, kBinsPad);
    if (!verify_banded(bm, p->mel_dense, kBinsPad)) {
        set_error("internal error: banded mel tables do not reproduce the mel basis");
        delete p;
        return HMFE_ERR_INVALID;
    }
    p->meta.n_slots = bm.n_slots;
    p->meta.total_trip = bm.total_trip;
    p->meta.n_mels = n_mels;
    for (int s = 0; s < kMaxSlots; ++s) {
        p->meta.trip[s] = s < bm.n_slots ? bm.trip[s] : 0;
        p->meta.wbase[s] = s < bm.n_slots ? bm.wbase[s] : 0;
    }
    const std::vector<float> win = half_hann_periodic(n_fft);
    const std::vector<float> tw = p->variant == HMFE_VARIANT_PAIR ? twiddle_plane_paired(n_fft) : twiddle_plane(n_fft, 32);
    int rc = upload_vec(win, &p->d_win);
    if (rc == HMFE_OK) rc = upload_vec(tw, reinterpret_cast<float**>(&p->d_tw));
    if (rc == HMFE_OK) rc = upload_vec(bm.w, &p->d_melw);
    if (rc == HMFE_OK) rc = upload_vec(bm.start, &p->d_start);
    if (rc == HMFE_OK) rc = upload_vec(bm.row, &p->d_row);
    p->tc_ok = tc_shape_ok(p);
    if (p->variant == HMFE_VARIANT_TC && !p->tc_ok) {
        set_error("the tensor-core variant needs n_mels <= 64, hop <= 512 and a zero Nyquist weight; use HMFE_VARIANT_PACKED");
        hmfe_logmel_plan_destroy(p);
        return HMFE_ERR_UNSUPPORTED;
    }
    if (rc == HMFE_OK && p->tc_ok) {
        rc = upload_vec(tc_build_a_words(p->mel_dense, n_mels, p->n_bins), &p->d_tc_a);
        if (rc == HMFE_OK) rc = upload_vec(std::vector<uint32_t>(4, 0u), &p->d_tc_status);
        const char* e = getenv("HMFE_TC_FFT_WARPS");
        if (e) p->tc_fft_warps = atoi(e) <= 8 ? 8 : 11;
    }
    if (rc != HMFE_OK) {
        hmfe_logmel_plan_destroy(p);
        return rc;
    }
    size_t tbytes = (size_t)(1024 * 8 + 1024 * 4 + bm.total_trip * 128 + bm.n_slots * 256 + 2 * kMaxSlots * 4);
    p->table_smem = (tbytes + 15) & ~(size_t)15;
    *plan = p;
    return HMFE_OK;
}

void hmfe_logmel_plan_destroy(hmfe_logmel_plan* p) {
    if (!p) return;
    cudaFree(p->d_win);
    cudaFree(p->d_tw);
    cudaFree(p->d_melw);
    cudaFree(p->d_start);
    cudaFree(p->d_row);
    cudaFree(p->d_tc_a);
    cudaFree(p->d_tc_status);
    cudaFree(p->d_gen_mel);
    cudaFree(p->d_gen_lo);
    cudaFree(p->d_gen_hi);
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    delete p;
}

int hmfe_logmel_plan_set_pad_mode(hmfe_logmel_plan* p, int pad_mode) {
    HMFE_REQUIRE(p, "NULL plan");
    HMFE_REQUIRE(pad_mode == HMFE_PAD_CONSTANT || pad_mode == HMFE_PAD_REFLECT, "bad pad_mode %d", pad_mode);
    HMFE_REQUIRE(p->variant != kVariantGeneric || pad_mode == HMFE_PAD_CONSTANT, "reflect padding needs n_fft = 1024");
    if (pad_mode == HMFE_PAD_REFLECT && p->variant == HMFE_VARIANT_TC) p->variant = HMFE_VARIANT_PACKED;  // same tables
    if (pad_mode == HMFE_PAD_REFLECT && p->variant != HMFE_VARIANT_PACKED) {
        set_error("reflect padding is built for the default (packed) variant only");
        return HMFE_ERR_UNSUPPORTED;
    }
    p->pad_mode = pad_mode;
    return HMFE_OK;
}

int64_t hmfe_logmel_num_frames(int64_t n_samples, int hop) { return hop > 0 && n_samples >= 0 ? 1 + n_samples / hop : -1; }

int hmfe_logmel_mel_basis(const hmfe_logmel_plan* p, float* h_out) {
    HMFE_REQUIRE(p && h_out, "NULL argument");
    std::copy(p->mel_dense.begin(), p->mel_dense.end(), h_out);
    return HMFE_OK;
}

int hmfe_logmel_last_launches(const hmfe_logmel_plan* p) { return p ? p->last_launches : 0; }

int hmfe_logmel_tc_status(hmfe_logmel_plan* p, uint32_t* h_status) {
    HMFE_REQUIRE(p && h_status, "NULL argument");
    *h_status = 0;
    if (!p->d_tc_status) return HMFE_OK;
    HMFE_CHECK_CUDA(cudaMemcpy(h_status, p->d_tc_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return HMFE_OK;
}

int hmfe_logmel_set_profile(hmfe_logmel_plan* p, int enable) {
    HMFE_REQUIRE(p, "NULL plan");
    p->profile = enable != 0;
    return HMFE_OK;
}

int hmfe_logmel_profile_ms(hmfe_logmel_plan* p, double* power_ms, double* finalize_ms, int* n_calls) {
    HMFE_REQUIRE(p, "NULL plan");
    double a = 0, f = 0;
    const int n = (int)(p->prof_events.size() / 3);
    for (int i = 0; i < n; ++i) {
        float t0 = 0, t1 = 0;
        HMFE_CHECK_CUDA(cudaEventSynchronize(p->prof_events[3 * i + 2]));
        HMFE_CHECK_CUDA(cudaEventElapsedTime(&t0, p->prof_events[3 * i], p->prof_events[3 * i + 1]));
        HMFE_CHECK_CUDA(cudaEventElapsedTime(&t1, p->prof_events[3 * i + 1], p->prof_events[3 * i + 2]));
        a += t0;
        f += t1;
    }
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    p->prof_events.clear();
    if (power_ms) *power_ms = a;
    if (finalize_ms) *finalize_ms = f;
    if (n_calls) *n_calls = n;
    return HMFE_OK;
}

int hmfe_logmel_batch_views(hmfe_logmel_plan* p, const float* d_wav, const int64_t* h_starts, const int64_t* h_lengths,
                            int64_t n_clips, float* d_out, int out_mode, void* stream) {
    return hmfe_logmel_batch_views2(p, d_wav, nullptr, h_starts, h_lengths, n_clips, d_out, out_mode, stream);
}

int hmfe_logmel_batch_views2(hmfe_logmel_plan* p, const float* d_wav, const float* d_wav_alt, const int64_t* h_starts,
                             const int64_t* h_lengths, int64_t n_clips, float* d_out, int out_mode, void* stream) {
    HMFE_REQUIRE(p && h_starts && h_lengths, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    HMFE_REQUIRE(out_mode >= 0 && out_mode <= 3, "bad out_mode %d", out_mode);
    p->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_wav && d_out, "NULL device pointer");
    HMFE_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "d_out must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int FR = (p->variant == HMFE_VARIANT_PACKED || p->variant == HMFE_VARIANT_TC) ? 4 : 2;

    bool uniform = true;
    const int64_t n0 = h_lengths[0];
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_lengths[i];
        HMFE_REQUIRE(n >= 0 && n < (int64_t)1 << 30 && (h_starts[i] >= 0 || d_wav_alt != nullptr),
                     "clip %lld has invalid start/length %lld/%lld",
                     (long long)i, (long long)h_starts[i], (long long)n);
        HMFE_REQUIRE(p->pad_mode != HMFE_PAD_REFLECT || n > p->n_fft / 2,
                     "clip %lld: reflect padding needs more than n_fft/2 = %d samples, got %lld", (long long)i,
                     p->n_fft / 2, (long long)n);
        uniform = uniform && n == n0 && h_starts[i] == i * n0;
    }

    LogmelBatch b{};
    b.wav = d_wav;
    b.wav_alt = d_wav_alt;
    b.out = d_out;
    b.n_clips = n_clips;
    b.hop = p->hop;
    b.status = p->d_tc_status;
    {
        const char* e = getenv("HMFE_LOGMEL_STAGGER_NS");
        b.stagger_ns = e ? atoi(e) : 500;  // measured on B200, c1: 0 -> 0.323 ms, 200 -> 0.308, 400..2000 -> 0.302-0.303
    }
    const size_t desc_bytes = uniform ? 0 : (4 * (size_t)n_clips + 2) * sizeof(int64_t);
    const size_t stats_bytes = ((size_t)n_clips * 2 * sizeof(unsigned) + 15) & ~(size_t)15;
    const size_t total_bytes = desc_bytes + stats_bytes + 16;
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = p->ring.acquire(total_bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    if (uniform) {
        const int64_t T = 1 + n0 / p->hop;
        b.uniform_n = (int)n0;
        b.uniform_T = (int)T;
        b.uniform_items = (int)((T + FR - 1) / FR);
        b.n_items = (int64_t)b.uniform_items * n_clips;
        b.n_frames_total = T * n_clips;
    } else {
        int64_t* hs = static_cast<int64_t*>(hbuf);
        int64_t* hl = hs + n_clips;
        int64_t* hf = hl + n_clips;
        int64_t* hi = hf + (n_clips + 1);
        hf[0] = hi[0] = 0;
        for (int64_t i = 0; i < n_clips; ++i) {
            const int64_t T = 1 + h_lengths[i] / p->hop;
            hs[i] = h_starts[i];
            hl[i] = h_lengths[i];
            hf[i + 1] = hf[i] + T;
            hi[i + 1] = hi[i] + (T + FR - 1) / FR;
        }
        b.n_items = hi[n_clips];
        b.n_frames_total = hf[n_clips];
        int64_t* dc = static_cast<int64_t*>(dbuf);
        b.clip_start = dc;
        b.clip_len = dc + n_clips;
        b.frame_off = dc + 2 * n_clips;
        b.item_prefix = dc + 3 * n_clips + 1;
        int rc = p->ring.upload(slot, desc_bytes, st);
        if (rc != HMFE_OK) return rc;
    }
    b.stats = reinterpret_cast<unsigned*>(static_cast<unsigned char*>(dbuf) + desc_bytes);
    b.queue = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(dbuf) + desc_bytes + stats_bytes);

    int rc = run_logmel(p, b, out_mode, st);
    if (rc != HMFE_OK) return rc;
    return p->ring.release(slot, st);
}

int64_t hmfe_logmel_device_workspace_bytes(int64_t n_clips) {
    return n_clips < 0 ? -1 : (((int64_t)n_clips * 2 * (int64_t)sizeof(unsigned) + 15) & ~(int64_t)15) + 16;
}

int hmfe_logmel_batch_device(hmfe_logmel_plan* p, const float* d_wav, const float* d_wav_alt, const int64_t* d_desc,
                             int64_t n_clips, float* d_out, int out_mode, void* d_workspace, void* stream) {
    HMFE_REQUIRE(p, "NULL plan");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    HMFE_REQUIRE(out_mode >= 0 && out_mode <= 3, "bad out_mode %d", out_mode);
    p->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_wav && d_desc && d_out && d_workspace, "NULL device pointer");
    HMFE_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0,
                 "d_out and d_workspace must be 16-byte aligned");
    HMFE_REQUIRE(p->pad_mode == HMFE_PAD_CONSTANT, "device-planned batches use constant padding");
    if (p->variant == kVariantGeneric) {
        set_error("device-planned batches need n_fft = 1024 (use hmfe_logmel_batch_views for other frame lengths)");
        return HMFE_ERR_UNSUPPORTED;
    }
    LogmelBatch b{};
    b.wav = d_wav;
    b.wav_alt = d_wav_alt ? d_wav_alt : d_wav;
    b.out = d_out;
    b.n_clips = n_clips;
    b.hop = p->hop;
    b.status = p->d_tc_status;
    {
        const char* e = getenv("HMFE_LOGMEL_STAGGER_NS");
        b.stagger_ns = e ? atoi(e) : 500;
    }
    b.clip_start = d_desc;
    b.clip_len = d_desc + n_clips;
    b.frame_off = d_desc + 2 * n_clips;
    b.item_prefix = d_desc + 3 * n_clips + 1;
    b.n_items_dev = b.item_prefix + n_clips;
    b.n_items = (int64_t)1 << 40;  // unknown to the host: full persistent grid
    b.stats = static_cast<unsigned*>(d_workspace);
    b.queue = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(d_workspace) +
                                                    (((size_t)n_clips * 2 * sizeof(unsigned) + 15) & ~(size_t)15));
    return run_logmel(p, b, out_mode, static_cast<cudaStream_t>(stream));
}

int hmfe_logmel_batch(hmfe_logmel_plan* p, const float* d_wav, const int64_t* h_offsets, int64_t n_clips, float* d_out,
                      int out_mode, void* stream) {
    HMFE_REQUIRE(p && h_offsets, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    std::vector<int64_t> len((size_t)n_clips);
    for (int64_t i = 0; i < n_clips; ++i) len[i] = h_offsets[i + 1] - h_offsets[i];
    return hmfe_logmel_batch_views(p, d_wav, h_offsets, len.data(), n_clips, d_out, out_mode, stream);
}

}  // extern "C"
