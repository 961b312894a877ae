// Biquad-cascade (SOS) IIR over ragged batches as a chunked scan, float64 arithmetic.
// Replaces scipy.signal.lfilter(b, a, x) as called by _butter_bandpass_filter
// (/root/reference/src/util.py:113-126): causal, single pass, zero initial state, float64
// result.  The cascade has the same transfer function as the (b, a) form the reference
// builds; SOS is the better-conditioned realisation (difference ~1e-8).
//
//   pass A  every chunk of C samples is filtered from a ZERO state -> end state z_j
//   pass B  per clip, sequential over chunks: s_{j+1} = M s_j + z_j   (M = zero-input
//           transition of the cascade over C samples, 2S x 2S, computed on the host)
//   pass C  every chunk is filtered again from its true initial state s_j -> output
//
// One lane owns one chunk (sequential recurrence); a warp owns 32 consecutive chunks and
// moves samples through a padded shared-memory tile so that global loads/stores stay
// coalesced 128-byte rows.
//
// Overlap ("warm-up") variant, used when the filter forgets fast enough: the zero-input
// response of a stable cascade decays geometrically, so a chunk that starts W samples early
// from a ZERO state reaches its true state to within ||A^W|| (W is chosen on the host so that
// ||A^W||_inf <= 1e-13, below the rounding noise of the float64 recurrence itself).  That makes
// ONE pass of (1 + W/C) x the work instead of two passes plus the carry scan, and it lets the
// pass also produce the per-hop energy sums the silence trim needs (hmfe_iir_sos_trim_batch),
// so the filtered signal is not read a second time.  Filters with slowly decaying poles fall
// back to the exact three-kernel scan above.
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "api_common.h"
#include "ctx.h"
#include "hmfe_common.cuh"
#include "trim_common.cuh"

namespace hmfe {

constexpr int kIirMaxSections = 8;
constexpr int kIirWarps = 4;

template <int S>
struct IirCoef {
    double b0[S], b1[S], b2[S], a1[S], a2[S];
    // Butterworth band-pass sections all have numerator g_k * (1, 0, -1).  When every section has
    // that form the gains are folded into one input gain and a section costs DADD + 2 DFMA
    // instead of 4 DFMA + DMUL.
    int bandpass_form;
    double gain;
};

struct IirBatch {
    const float* x;
    float* y32;
    double* y64;
    const int64_t* clip_off;      // [n_clips+1]
    const int64_t* chunk_prefix;  // [n_clips+1]
    double* zstate;               // [n_chunks][2S] end states from zero state
    double* init;                 // [n_chunks][2S] true initial states
    const double* M;              // [2S][2S] row-major
    int64_t n_clips, n_chunks;
    int C;
};

// GAIN = false leaves the folded band-pass gain to the caller (applied to the float32 output)
template <int S, bool BP, bool GAIN = true>
HMFE_D double cascade(const IirCoef<S>& cf, double v, double (&s1)[S], double (&s2)[S]) {
    if (BP && GAIN) v *= cf.gain;
#pragma unroll
    for (int k = 0; k < S; ++k) {  // direct form II transposed
        if (BP) {                  // b = (1, 0, -1)
            const double y = v + s1[k];
            s1[k] = fma(-cf.a1[k], y, s2[k]);
            s2[k] = fma(-cf.a2[k], y, -v);
            v = y;
        } else {
            const double y = fma(cf.b0[k], v, s1[k]);
            s1[k] = fma(cf.b1[k], v, fma(-cf.a1[k], y, s2[k]));
            s2[k] = fma(cf.b2[k], v, -cf.a2[k] * y);
            v = y;
        }
    }
    return v;
}

template <int S, bool FINAL, bool BP>
__global__ void __launch_bounds__(kIirWarps * 32) iir_chunk_kernel(const IirBatch b, const IirCoef<S> cf) {
    __shared__ double s_tile[kIirWarps][32][33];
    __shared__ int64_t s_row[kIirWarps][32];
    __shared__ int s_valid[kIirWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * kIirWarps + warp) * 32 + lane;
    double s1[S], s2[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s1[k] = s2[k] = 0.0;
    int64_t row = 0;
    int valid = 0;
    if (g < b.n_chunks) {
        int64_t lo = 0, hi = b.n_clips;  // largest clip with chunk_prefix[clip] <= g
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (b.chunk_prefix[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t j = g - b.chunk_prefix[lo];
        const int64_t c0 = b.clip_off[lo], n = b.clip_off[lo + 1] - c0;
        row = c0 + j * b.C;
        valid = (int)min((int64_t)b.C, n - j * b.C);
        if (FINAL) {
            const double* in = b.init + g * (2 * S);
#pragma unroll
            for (int k = 0; k < S; ++k) {
                s1[k] = in[2 * k];
                s2[k] = in[2 * k + 1];
            }
        }
    }
    s_row[warp][lane] = row;
    s_valid[warp][lane] = valid;
    __syncwarp();
    double(*tile)[33] = s_tile[warp];
    // software pipeline: the 32 row loads of step t0+32 are in flight while step t0 is filtered
    float nxt[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) nxt[r] = lane < s_valid[warp][r] ? __ldg(b.x + s_row[warp][r] + lane) : 0.0f;
    for (int t0 = 0; t0 < b.C; t0 += 32) {
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = (double)nxt[r];
        __syncwarp();
        if (t0 + 32 < b.C) {
            const int i = t0 + 32 + lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) nxt[r] = i < s_valid[warp][r] ? __ldg(b.x + s_row[warp][r] + i) : 0.0f;
        }
        if (t0 < valid) {
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const double y = cascade<S, BP>(cf, tile[lane][k], s1, s2);
                if (FINAL) tile[lane][k] = y;
            }
        }
        __syncwarp();
        if (FINAL) {
            const int i = t0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                if (i < s_valid[warp][r]) {
                    const double y = tile[r][lane];
                    if (b.y32) b.y32[s_row[warp][r] + i] = (float)y;
                    if (b.y64) b.y64[s_row[warp][r] + i] = y;
                }
            }
            __syncwarp();
        }
    }
    if (!FINAL && g < b.n_chunks) {
        double* z = b.zstate + g * (2 * S);
#pragma unroll
        for (int k = 0; k < S; ++k) {
            z[2 * k] = s1[k];
            z[2 * k + 1] = s2[k];
        }
    }
}

// one warp per clip; lane i < 2S owns state component i and row i of M
template <int S>
__global__ void __launch_bounds__(128) iir_carry_kernel(const IirBatch b) {
    constexpr int D = 2 * S;
    constexpr int PF = 8;  // chunks whose end states are fetched ahead of the sequential recurrence
    const int lane = threadIdx.x & 31;
    const int64_t clip = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (clip >= b.n_clips) return;
    double m[D];
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = lane < D ? b.M[lane * D + k] : 0.0;
    double s = 0.0;
    const int64_t g0 = b.chunk_prefix[clip], g1 = b.chunk_prefix[clip + 1];
    for (int64_t gb = g0; gb < g1; gb += PF) {
        double z[PF];
#pragma unroll
        for (int j = 0; j < PF; ++j) z[j] = (lane < D && gb + j < g1) ? b.zstate[(gb + j) * D + lane] : 0.0;
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            if (gb + j < g1) {
                if (lane < D) b.init[(gb + j) * D + lane] = s;
                double acc0 = z[j], acc1 = 0.0;
#pragma unroll
                for (int k = 0; k < D; k += 2) {
                    acc0 = fma(m[k], __shfl_sync(0xffffffffu, s, k), acc0);
                    acc1 = fma(m[k + 1], __shfl_sync(0xffffffffu, s, k + 1), acc1);
                }
                s = acc0 + acc1;
            }
        }
    }
}


// ------------------------------------------------------------------------------ overlap variant
struct IirOverlapBatch {
    const float* x;
    float* y32;
    double* y64;
    const int64_t* clip_off;      // [n_clips+1]
    const int64_t* chunk_prefix;  // [n_clips+1]
    const int64_t* hop_off;       // [n_clips+1] first hop-energy slot of each clip (POWER only)
    float* hop_energy;            // [sum ceil(n/hop)] sum of y^2 (float32 squares) per hop block
    int64_t n_clips, n_chunks;
    int C, W;                     // multiples of 32; POWER: C is a multiple of hop
    int hop_steps;                // hop / 32 (POWER only)
};

struct __align__(16) IirRow {
    long long off;  // element offset of stream position t = 0 (may be negative: never dereferenced there)
    int lo, span;   // stream positions [lo, lo + span) hold clip samples
};

// T = tile element: float (y32 only) or double (y64 requested)
template <int S, bool BP, typename T, bool POWER>
__global__ void __launch_bounds__(kIirWarps * 32, 4) iir_overlap_kernel(const IirOverlapBatch b, const IirCoef<S> cf) {
    __shared__ T s_tile[kIirWarps][32][33];
    __shared__ IirRow s_rowd[kIirWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * kIirWarps + warp) * 32 + lane;
    double s1[S], s2[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s1[k] = s2[k] = 0.0;
    IirRow me{0, 0, 0};
    int valid = 0;
    int64_t hop_base = 0;
    if (g < b.n_chunks) {
        int64_t lo = 0, hi = b.n_clips;  // largest clip with chunk_prefix[clip] <= g
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (b.chunk_prefix[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t j = g - b.chunk_prefix[lo];
        const int64_t c0 = b.clip_off[lo], n = b.clip_off[lo + 1] - c0;
        const int64_t out0 = j * b.C;  // clip-relative first output sample of this chunk
        valid = (int)min((int64_t)b.C, n - out0);
        me.off = c0 + out0 - b.W;
        me.lo = (int)max((int64_t)0, (int64_t)b.W - out0);  // nothing before the clip start: exact zero state
        me.span = b.W + valid - me.lo;
        if (POWER) hop_base = b.hop_off[lo] + j * (b.C / (b.hop_steps * 32));
    }
    s_rowd[warp][lane] = me;
    __syncwarp();
    const int hi_self = b.W + valid;
    T(*tile)[33] = s_tile[warp];
    const IirRow* rows = s_rowd[warp];
    const int t_end = b.W + b.C;
    // the whole warp skips the steps in which no lane has clip samples yet (first chunks of a clip)
    int t_first = me.span > 0 ? (me.lo & ~31) : t_end;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t_first = min(t_first, __shfl_xor_sync(0xffffffffu, t_first, d));

    float nxt[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const IirRow rd = rows[r];
        const int i = t_first + lane;
        nxt[r] = (unsigned)(i - rd.lo) < (unsigned)rd.span ? __ldg(b.x + rd.off + i) : 0.0f;
    }
    float hop_acc = 0.0f;
    int hop_step = 0, hop_idx = 0;
    for (int t0 = t_first; t0 < t_end; t0 += 32) {
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = (T)nxt[r];
        __syncwarp();
        if (t0 + 32 < t_end) {
            const int i = t0 + 32 + lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const IirRow rd = rows[r];
                nxt[r] = (unsigned)(i - rd.lo) < (unsigned)rd.span ? __ldg(b.x + rd.off + i) : 0.0f;
            }
        }
        const bool emit = t0 >= b.W;  // warp uniform (W is a multiple of 32)
        if (t0 < hi_self) {
            if (!emit) {
#pragma unroll 8
                for (int k = 0; k < 32; ++k) cascade<S, BP>(cf, (double)tile[lane][k], s1, s2);
            } else if (t0 + 32 <= hi_self) {
                float e = 0.0f;
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const double y = cascade<S, BP>(cf, (double)tile[lane][k], s1, s2);
                    tile[lane][k] = (T)y;
                    if (POWER) {
                        const float f = (float)y;
                        e = fmaf(f, f, e);
                    }
                }
                hop_acc += e;
            } else {  // the step that holds the end of the clip
                float e = 0.0f;
                const int kmax = hi_self - t0;
#pragma unroll 1
                for (int k = 0; k < kmax; ++k) {
                    const double y = cascade<S, BP>(cf, (double)tile[lane][k], s1, s2);
                    tile[lane][k] = (T)y;
                    if (POWER) {
                        const float f = (float)y;
                        e = fmaf(f, f, e);
                    }
                }
                hop_acc += e;
            }
        }
        __syncwarp();
        if (emit) {
            const int i = t0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const IirRow rd = rows[r];
                if (i < rd.lo + rd.span) {
                    const T y = tile[r][lane];
                    if (sizeof(T) == 4) {
                        b.y32[rd.off + i] = (float)y;
                    } else {
                        if (b.y32) b.y32[rd.off + i] = (float)y;
                        b.y64[rd.off + i] = (double)y;
                    }
                }
            }
            __syncwarp();
            if (POWER && ++hop_step == b.hop_steps) {  // a hop block is complete (warp uniform)
                if (hop_idx * b.hop_steps * 32 < valid) b.hop_energy[hop_base + hop_idx] = hop_acc;
                hop_acc = 0.0f;
                hop_step = 0;
                ++hop_idx;
            }
        }
    }
}

// ------------------------------------------------------------------------------ overlap, 128-bit rows
// Same algorithm with 16-byte global accesses.  Every clip is addressed from the 16-byte aligned
// element at or before its first sample (aligned origin = c0 - s, s in 0..3): chunk boundaries,
// the warm-up length and the 128-sample steps are multiples of 4 in that coordinate, so a lane's
// float4 never straddles two chunks and rows move as LDG.128 / STG.128.  Samples in front of the
// clip (the s alignment slots, and the warm-up of its first chunks) are loaded as zeros, which
// leaves the zero state untouched.  Steps in which some row has a clip edge inside a float4 take
// a scalar (per element predicated) variant of the load / store phase.
//
// Hop energies are accumulated in the aligned coordinate as well: group g holds the samples at
// aligned positions [g*hop, (g+1)*hop) as {sum over all but the first four, the four first
// squares}; trim_index_hop4_kernel re-assembles the clip-relative hop block h from groups h and
// h+1 (additions only).
struct IirOverlap4Batch {
    const float* x;
    float* y32;
    const int64_t* clip_off;      // [n_clips+1]
    const int64_t* chunk_prefix;  // [n_clips+1]  chunks of ceil((s + n) / C)
    const int64_t* group_off;     // [n_clips+1]  groups of ceil((s + n) / hop)    (POWER only)
    float* group_energy;          // [n_groups][8]                                  (POWER only)
    int64_t n_clips, n_chunks;
    int C, W;                     // multiples of 128; POWER: C is a multiple of hop
    int hop;                      // multiple of 32 (POWER only)
    int align;                    // (address of x / 4) mod 4
    int l2_prefetch;              // cp.async with the L2::256B hint (HMFE_IIR_L2PF, A/B)
};

#ifndef HMFE_IIR_CONV_DEFAULT
#define HMFE_IIR_CONV_DEFAULT 0
#endif
constexpr int kRing = 4;                  // 32-sample column blocks per tile row
constexpr int kRow4 = 32 * kRing + 4;     // tile row stride in floats: 16-byte aligned rows, conflict-free LDS.128 by row

struct __align__(16) IirRow4 {
    long long base;  // element index of stream position 0 (16-byte aligned address; may be negative)
    int lo, hi;      // stream positions [lo, hi) hold clip samples
};

// float <-> double without the conversion instructions (F2F.F64.F32 / F2F.F32.F64 issue at a fraction of the DFMA
// rate and sit at both ends of every sample's dependency chain): exponent re-biasing on the integer pipe, which
// this kernel leaves idle.  Exact for normal numbers; float denormals (|x| < 1.2e-38) and doubles below the float
// normal range are flushed to (signed) zero, far below the 1e-10 agreement the tests demand.  CONV bit 0: input,
// bit 1: output (round to nearest even on the 29 dropped mantissa bits, like cvt.rn).
HMFE_D double f32_to_f64_int(float x) {
    const unsigned u = __float_as_uint(x);
    const unsigned a = u & 0x7fffffffu;
    unsigned hi = (u & 0x80000000u) | ((a >> 3) + 0x38000000u);
    unsigned lo = u << 29;
    if (a < 0x00800000u) {
        hi = u & 0x80000000u;
        lo = 0u;
    }
    return __hiloint2double((int)hi, (int)lo);
}
HMFE_D float f64_to_f32_int(double y) {
    const unsigned hi = (unsigned)__double2hiint(y), lo = (unsigned)__double2loint(y);
    const unsigned mh = hi & 0x7fffffffu;
    // round to nearest even at bit 29 of the 63-bit magnitude: add 0x0fffffff + lsb, carry into the high word
    const unsigned lsb = (lo >> 29) & 1u;
    const unsigned long long m = (((unsigned long long)mh << 32) | lo) + 0x0fffffffull + lsb;
    unsigned f = (unsigned)(m >> 29) - (896u << 23);
    if (mh < (897u << 20)) f = 0u;  // below the float normal range
    return __uint_as_float(f | (hi & 0x80000000u));
}
template <int CONV>
HMFE_D double to_f64(float x) {
    if constexpr (CONV & 1) return f32_to_f64_int(x);
    return (double)x;
}
template <int CONV>
HMFE_D float to_f32(double y) {
    if constexpr (CONV & 2) return f64_to_f32_int(y);
    return (float)y;
}

template <int S, bool BP, int CONV>
HMFE_D void iir_block32(const IirCoef<S>& cf, float gain, float* row, double (&s1)[S], double (&s2)[S], bool emit,
                        float& body, float (&head)[4]) {
    // 32 consecutive samples of this lane's row: read as 8 float4, filter, write back in place
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float4 v = *reinterpret_cast<const float4*>(row + 4 * u);
        float* e = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float f = to_f32<CONV>(cascade<S, BP, false>(cf, to_f64<CONV>(e[c]), s1, s2)) * gain;
            e[c] = f;
            if (u == 0)
                head[c] = f * f;
            else
                body = fmaf(f, f, body);
        }
        if (emit) *reinterpret_cast<float4*>(row + 4 * u) = v;
    }
}

HMFE_D void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// the same with an L2 prefetch hint: the miss brings in the 256-byte line pair the row's next visit will ask for
HMFE_D void cp_async16_pf(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
HMFE_D void cp_async4(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
HMFE_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
HMFE_D void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// The tile of a warp is a ring of kRing 32-sample column blocks.  Block k of the stream is
//   loaded   by cp.async (LDGSTS, zero fill outside the clip) kRing blocks ahead, 4 rows per instruction,
//   filtered in place by the lane that owns the row,
//   stored   row-major (STG.128, 4 rows per instruction) and its columns handed to block k + kRing,
// so the global-load latency of a block is covered by the filtering of the kRing - 1 blocks before it.
// The folded band-pass gain is applied to the float32 output (FP32 pipe) instead of the float64 input.
// Measured on B200 (c2, 1.78 G samples): ring 4 x 3 CTAs/SM 3.55 ms, 3 x 3 3.58, 3 x 4 3.72, 2 x 4 3.73, 2 x 5 3.90.
// GRP: 32-sample blocks per load / store group.  With GRP = 2 a row is fetched and written 256 bytes at a time (two column
// blocks = half the ring per visit), i.e. half as many visits to every DRAM page; the ring then runs one group ahead.
template <int S, bool BP, bool POWER, int CONV, int GRP = 1>
__global__ void __launch_bounds__(kIirWarps * 32, 3) iir_overlap4_kernel(const IirOverlap4Batch b, const IirCoef<S> cf) {
    extern __shared__ __align__(16) unsigned char iir_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gain = BP ? (float)cf.gain : 1.0f;
    float(*tile)[kRow4] = reinterpret_cast<float(*)[kRow4]>(iir_smem) + warp * 32;
    IirRow4* rows = reinterpret_cast<IirRow4*>(iir_smem + (size_t)kIirWarps * 32 * kRow4 * sizeof(float)) + warp * 32;
    const int64_t g = ((int64_t)blockIdx.x * kIirWarps + warp) * 32 + lane;
    double s1[S], s2[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s1[k] = s2[k] = 0.0;
    IirRow4 me{0, 0, 0};
    int64_t group_base = 0;
    if (g < b.n_chunks) {
        int64_t lo = 0, hi = b.n_clips;  // largest clip with chunk_prefix[clip] <= g
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (b.chunk_prefix[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t j = g - b.chunk_prefix[lo];
        const int64_t c0 = b.clip_off[lo], n = b.clip_off[lo + 1] - c0;
        const int s = (int)((c0 + b.align) & 3);
        const int64_t p0 = j * b.C;  // aligned-coordinate start of this chunk
        me.base = c0 - s + p0 - b.W;
        me.lo = (int)max((int64_t)0, (int64_t)s + b.W - p0);
        me.hi = b.W + (int)min((int64_t)b.C, (int64_t)s + n - p0);
        if (POWER) group_base = b.group_off[lo] + j * (b.C / b.hop);
    }
    rows[lane] = me;
    __syncwarp();
    const int hi_self = me.hi;
    const int valid_len = me.hi - b.W;  // chunk positions [0, valid_len) exist (<= 0 for idle lanes)
    const int t_end = b.W + b.C;
    int t_first = me.hi > me.lo ? (me.lo & ~(32 * GRP - 1)) : t_end;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t_first = min(t_first, __shfl_xor_sync(0xffffffffu, t_first, d));
    // 32-sample blocks in which this lane's own row has a clip edge that is not a multiple of 4
    const int edge_lo = (me.hi > me.lo && (me.lo & 3)) ? (me.lo >> 5) : -1;
    const int edge_hi = (me.hi > me.lo && (me.hi & 3)) ? (me.hi >> 5) : -1;
    // the rows this lane moves in the row-major phases: 4*i + (lane >> 3), columns 4*(lane & 7) .. +3
    const int rsub = lane >> 3, c4 = 4 * (lane & 7);
    IirRow4 mv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mv[i] = rows[4 * i + rsub];

    auto issue_load = [&](int tb) {  // stream positions [tb, tb + 32) of every row -> column block (tb >> 5) & 3
        const int col = 32 * ((tb >> 5) % kRing) + c4;
        const int t = tb + c4;
        const bool edge = __any_sync(0xffffffffu, edge_lo == (tb >> 5) || edge_hi == (tb >> 5));
        if (!edge) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool in = t >= mv[i].lo && t < mv[i].hi;
                if (b.l2_prefetch)
                    cp_async16_pf(&tile[4 * i + rsub][col], in ? (const void*)(b.x + mv[i].base + t) : (const void*)b.x, in ? 16 : 0);
                else
                    cp_async16(&tile[4 * i + rsub][col], in ? (const void*)(b.x + mv[i].base + t) : (const void*)b.x, in ? 16 : 0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool in = t + c >= mv[i].lo && t + c < mv[i].hi;
                    cp_async4(&tile[4 * i + rsub][col + c], in ? (const void*)(b.x + mv[i].base + t + c) : (const void*)b.x,
                              in ? 4 : 0);
                }
        }
    };

    static_assert(kRing % GRP == 0, "the ring holds whole groups");
#pragma unroll 1
    for (int k = 0; k < kRing; k += GRP) {
#pragma unroll 1
        for (int h = 0; h < GRP; ++h)
            if (t_first + 32 * (k + h) < t_end) issue_load(t_first + 32 * (k + h));
        cp_async_commit();
    }

    float body = 0.0f, head[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q[4] = {0.0f, 0.0f, 0.0f, 0.0f}, grp = 0.0f;
    int gidx = 0, pos_in_group = 0;
#pragma unroll 1
    for (int tg = t_first; tg < t_end; tg += 32 * GRP) {
        cp_async_wait<kRing / GRP - 1>();
        __syncwarp();
#pragma unroll 1
      for (int tb = tg; tb < tg + 32 * GRP; tb += 32) {  // (GRP = 1: one trip)
        const int col0 = 32 * ((tb >> 5) % kRing);
        const bool emit = tb >= b.W;  // warp uniform
        bool group_start = false;
        if (POWER && emit) {
            if (pos_in_group == b.hop) {  // a group is complete (warp uniform)
                if (gidx * b.hop < valid_len) {
                    float4* dst = reinterpret_cast<float4*>(b.group_energy + (group_base + gidx) * 8);
                    dst[0] = make_float4(grp, q[0], q[1], q[2]);
                    dst[1] = make_float4(q[3], 0.0f, 0.0f, 0.0f);
                }
                ++gidx;
                pos_in_group = 0;
                grp = 0.0f;
            }
            group_start = pos_in_group == 0;
            pos_in_group += 32;
        }
        // ---- filter: this lane's row
        const int rem = hi_self - tb;
        if (rem >= 32) {
            body = 0.0f;
            iir_block32<S, BP, CONV>(cf, gain, &tile[lane][col0], s1, s2, emit, body, head);
            if (POWER && emit) {
                if (group_start) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) q[c] = head[c];
                } else {
                    body += (head[0] + head[1]) + (head[2] + head[3]);
                }
                grp += body;
            }
        } else if (rem > 0) {  // the block that holds the end of the clip
            float acc = 0.0f;
            if (POWER && emit && group_start) q[0] = q[1] = q[2] = q[3] = 0.0f;
#pragma unroll 1
            for (int k = 0; k < rem; ++k) {
                const float f = to_f32<CONV>(cascade<S, BP, false>(cf, to_f64<CONV>(tile[lane][col0 + k]), s1, s2)) * gain;
                tile[lane][col0 + k] = f;
                if (POWER && emit) {
                    if (group_start && k < 4) {  // (no dynamic register indexing)
                        const float sq = f * f;
                        if (k == 0) q[0] = sq;
                        if (k == 1) q[1] = sq;
                        if (k == 2) q[2] = sq;
                        if (k == 3) q[3] = sq;
                    } else {
                        acc = fmaf(f, f, acc);
                    }
                }
            }
            grp += acc;
        } else if (POWER && emit && group_start) {
            q[0] = q[1] = q[2] = q[3] = 0.0f;
        }
      }
        __syncwarp();
        // ---- store: rows 4*i + rsub, positions tb + c4 .. + 3
#pragma unroll 1
      for (int tb = tg; tb < tg + 32 * GRP; tb += 32) {
        const int col0 = 32 * ((tb >> 5) % kRing);
        const bool emit = tb >= b.W;
        if (emit) {
            const int t = tb + c4;
            const bool edge = __any_sync(0xffffffffu, edge_lo == (tb >> 5) || edge_hi == (tb >> 5));
            if (!edge) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int olo = max(mv[i].lo, b.W);
                    if (t >= olo && t < mv[i].hi)
                        *reinterpret_cast<float4*>(b.y32 + mv[i].base + t) =
                            *reinterpret_cast<const float4*>(&tile[4 * i + rsub][col0 + c4]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int olo = max(mv[i].lo, b.W);
                    const float4 v = *reinterpret_cast<const float4*>(&tile[4 * i + rsub][col0 + c4]);
                    const float* e = reinterpret_cast<const float*>(&v);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (t + c >= olo && t + c < mv[i].hi) b.y32[mv[i].base + t + c] = e[c];
                }
            }
        }
      }
        __syncwarp();
        // ---- hand the column blocks to the stream blocks one ring further
#pragma unroll 1
        for (int tb = tg; tb < tg + 32 * GRP; tb += 32)
            if (tb + 32 * kRing < t_end) issue_load(tb + 32 * kRing);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (POWER && gidx * b.hop < valid_len) {  // the last group of the chunk
        float4* dst = reinterpret_cast<float4*>(b.group_energy + (group_base + gidx) * 8);
        dst[0] = make_float4(grp, q[0], q[1], q[2]);
        dst[1] = make_float4(q[3], 0.0f, 0.0f, 0.0f);
    }
}

// ------------------------------------------------------------------------------ overlap, 128-bit rows, pipelined cascade
// The same pass with the recurrence re-timed: one TICK advances every section by one sample, section s working on
// the sample section s - 1 finished in the tick before (a systolic cascade; the values travel in the registers
// p[1..S-1]).  Every section sees the same inputs and does the same operations in the same order as cascade(), so the
// result is bit-identical; the output of a sample leaves the last section S - 1 ticks after the sample went in.
// What changes is the shape of the dependency graph: the S section steps of a tick are independent of each other,
// so a ROLLED loop of four ticks per iteration has all the instruction-level parallelism there is, with nothing to
// drain or refill at its back edge.  That rolled loop is what lets the shared-memory reads be software-pipelined by
// hand: ptxas sank the LDS.128 of the unrolled 32-sample block of iir_overlap4_kernel to ~30 instructions before their
// first use whatever the source said, and ncu showed the first conversion of every float4 waiting on the shared-memory
// scoreboard (9.5 % of the kernel's stall samples).  Here the float4 of input unit i + 2 is requested, and the one of
// unit i + 1 converted to double, while the ticks of unit i run.
// To keep outputs and 16-byte tile columns aligned the pipeline delay is padded to a multiple of four ticks
// (D = 4 DU; S = 5: exactly 4), and the INPUT side runs DU units ahead: iteration k of the block loop produces the
// outputs of block k from the input units DU .. 7 + DU of block k, i.e. it needs block k + 1 resident as well.
template <int S, bool BP>
struct IirPipe {
    static constexpr int DU = S > 1 ? (S - 1 + 3) / 4 : 0;  // delay in 4-sample units
    static constexpr int E = 4 * DU - (S - 1);              // extra single-sample delays behind the last section
    double z1[S], z2[S], p[S];
    float dl[E > 0 ? E : 1];
    HMFE_D void reset() {
#pragma unroll
        for (int k = 0; k < S; ++k) z1[k] = z2[k] = p[k] = 0.0;
#pragma unroll
        for (int k = 0; k < (E > 0 ? E : 1); ++k) dl[k] = 0.0f;
    }
    // x enters section 0; returns the output of the sample that entered 4 DU ticks ago (the caller rounds it to float,
    // like the (float) cast of the other one-pass kernels)
    HMFE_D double tick(const IirCoef<S>& cf, double x) {
        double out = 0.0;
#pragma unroll
        for (int s = S - 1; s >= 0; --s) {
            const double v = s == 0 ? x : p[s];
            double y;
            if (BP) {  // b = (1, 0, -1), as in cascade()
                y = v + z1[s];
                z1[s] = fma(-cf.a1[s], y, z2[s]);
                z2[s] = fma(-cf.a2[s], y, -v);
            } else {
                y = fma(cf.b0[s], v, z1[s]);
                z1[s] = fma(cf.b1[s], v, fma(-cf.a1[s], y, z2[s]));
                z2[s] = fma(cf.b2[s], v, -cf.a2[s] * y);
            }
            if (s == S - 1)
                out = y;
            else
                p[s + 1] = y;
        }
        if (E > 0) {  // (only for section counts other than 1 and 5) pad the delay to a multiple of four ticks
            const float f = (float)out;
            const float g = dl[E - 1];
#pragma unroll
            for (int k = E - 1; k > 0; --k) dl[k] = dl[k - 1];
            dl[0] = f;
            return (double)g;  // exact: the value was rounded to float on the way in
        }
        return out;
    }
};

template <int S, bool BP, bool POWER>
__global__ void __launch_bounds__(kIirWarps * 32, 3) iir_pipe4_kernel(const IirOverlap4Batch b, const IirCoef<S> cf) {
    using Pipe = IirPipe<S, BP>;
    constexpr int DU = Pipe::DU;
    constexpr int kPending = DU > 0 ? kRing - 2 : kRing - 1;  // cp.async groups that may still be in flight at a block's start
    extern __shared__ __align__(16) unsigned char iir_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gain = BP ? (float)cf.gain : 1.0f;
    float(*tile)[kRow4] = reinterpret_cast<float(*)[kRow4]>(iir_smem) + warp * 32;
    IirRow4* rows = reinterpret_cast<IirRow4*>(iir_smem + (size_t)kIirWarps * 32 * kRow4 * sizeof(float)) + warp * 32;
    const int64_t g = ((int64_t)blockIdx.x * kIirWarps + warp) * 32 + lane;
    IirRow4 me{0, 0, 0};
    int64_t group_base = 0;
    if (g < b.n_chunks) {
        int64_t lo = 0, hi = b.n_clips;  // largest clip with chunk_prefix[clip] <= g
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (b.chunk_prefix[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t j = g - b.chunk_prefix[lo];
        const int64_t c0 = b.clip_off[lo], n = b.clip_off[lo + 1] - c0;
        const int s = (int)((c0 + b.align) & 3);
        const int64_t p0 = j * b.C;  // aligned-coordinate start of this chunk
        me.base = c0 - s + p0 - b.W;
        me.lo = (int)max((int64_t)0, (int64_t)s + b.W - p0);
        me.hi = b.W + (int)min((int64_t)b.C, (int64_t)s + n - p0);
        if (POWER) group_base = b.group_off[lo] + j * (b.C / b.hop);
    }
    rows[lane] = me;
    __syncwarp();
    const int hi_self = me.hi;
    const int valid_len = me.hi - b.W;  // chunk positions [0, valid_len) exist (<= 0 for idle lanes)
    const int t_end = b.W + b.C;
    int t_first = me.hi > me.lo ? (me.lo & ~31) : t_end;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t_first = min(t_first, __shfl_xor_sync(0xffffffffu, t_first, d));
    const int edge_lo = (me.hi > me.lo && (me.lo & 3)) ? (me.lo >> 5) : -1;
    const int edge_hi = (me.hi > me.lo && (me.hi & 3)) ? (me.hi >> 5) : -1;
    const int rsub = lane >> 3, c4 = 4 * (lane & 7);
    IirRow4 mv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mv[i] = rows[4 * i + rsub];

    auto issue_load = [&](int tb) {  // stream positions [tb, tb + 32) of every row -> column block (tb >> 5) % kRing
        const int col = 32 * ((tb >> 5) % kRing) + c4;
        const int t = tb + c4;
        const bool edge = __any_sync(0xffffffffu, edge_lo == (tb >> 5) || edge_hi == (tb >> 5));
        if (!edge) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool in = t >= mv[i].lo && t < mv[i].hi;
                cp_async16(&tile[4 * i + rsub][col], in ? (const void*)(b.x + mv[i].base + t) : (const void*)b.x, in ? 16 : 0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool in = t + c >= mv[i].lo && t + c < mv[i].hi;
                    cp_async4(&tile[4 * i + rsub][col + c], in ? (const void*)(b.x + mv[i].base + t + c) : (const void*)b.x,
                              in ? 4 : 0);
                }
        }
    };
#pragma unroll 1
    for (int k = 0; k < kRing; ++k) {
        if (t_first + 32 * k < t_end) issue_load(t_first + 32 * k);
        cp_async_commit();
    }

    // this lane's row: the float4 at stream position t (a multiple of 4).  Positions beyond the end of the stream read
    // whatever the ring holds: they only reach outputs beyond the end, which are neither stored nor counted.
    static_assert((kRing & (kRing - 1)) == 0, "the ring offset is a mask");
    const unsigned my_row = (unsigned)__cvta_generic_to_shared(&tile[lane][0]);
    auto fetch = [&](int t) -> float4 {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(my_row + 4u * ((unsigned)t & (32u * kRing - 1u))));
        return v;
    };
    Pipe pipe;
    pipe.reset();
    double xin[4] = {0.0, 0.0, 0.0, 0.0};       // input unit that enters the pipeline next, converted
    float4 raw1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // the unit after it
    int tin = t_first;                           // stream position of the unit in xin
    if (t_first < t_end) {
        cp_async_wait<kPending>();
        __syncwarp();
        if (DU > 0) {  // prime the pipeline: the first DU input units produce no output
#pragma unroll 1
            for (int i = 0; i < DU; ++i) {
                const float4 v = fetch(tin);
                pipe.tick(cf, (double)v.x);
                pipe.tick(cf, (double)v.y);
                pipe.tick(cf, (double)v.z);
                pipe.tick(cf, (double)v.w);
                tin += 4;
            }
        }
        const float4 v = fetch(tin);
        xin[0] = (double)v.x;
        xin[1] = (double)v.y;
        xin[2] = (double)v.z;
        xin[3] = (double)v.w;
        raw1 = fetch(tin + 4);
    }

    float head[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q[4] = {0.0f, 0.0f, 0.0f, 0.0f}, grp = 0.0f;
    int gidx = 0, pos_in_group = 0;
#pragma unroll 1
    for (int tb = t_first; tb < t_end; tb += 32) {
        cp_async_wait<kPending>();  // blocks tb and (input look-ahead) tb + 32 are resident
        __syncwarp();
        const int col0 = 32 * ((tb >> 5) % kRing);
        const bool emit = tb >= b.W;  // warp uniform
        bool group_start = false;
        if (POWER && emit) {
            if (pos_in_group == b.hop) {  // a group is complete (warp uniform)
                if (gidx * b.hop < valid_len) {
                    float4* dst = reinterpret_cast<float4*>(b.group_energy + (group_base + gidx) * 8);
                    dst[0] = make_float4(grp, q[0], q[1], q[2]);
                    dst[1] = make_float4(q[3], 0.0f, 0.0f, 0.0f);
                }
                ++gidx;
                pos_in_group = 0;
                grp = 0.0f;
            }
            group_start = pos_in_group == 0;
            pos_in_group += 32;
        }
        // ---- filter: the eight output units of this lane's row.  The rounding / gain / energy / store of unit u - 1
        // is written between the request for input unit u + 2 and the ticks of unit u, which it does not depend on.
        const int rem = hi_self - tb;  // outputs at block positions >= rem lie beyond the clip: no energy from them
        float body = 0.0f;
        const unsigned out_row = my_row + 4u * (unsigned)col0;
        auto block = [&](auto emit_c, auto mask_c) {
            constexpr bool EMIT = decltype(emit_c)::value, MASK = decltype(mask_c)::value;
            double yq[4];
            auto ticks = [&]() {
                yq[0] = pipe.tick(cf, xin[0]);
                yq[1] = pipe.tick(cf, xin[1]);
                yq[2] = pipe.tick(cf, xin[2]);
                yq[3] = pipe.tick(cf, xin[3]);
            };
            auto emit_unit = [&](int u, auto head_c) {  // output unit u from yq; head_c: the unit whose squares are kept apart
                constexpr bool HEAD = decltype(head_c)::value;
                if (!EMIT) return;
                float4 o;
                o.x = (float)yq[0] * gain;
                o.y = (float)yq[1] * gain;
                o.z = (float)yq[2] * gain;
                o.w = (float)yq[3] * gain;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_row + 16u * (unsigned)u), "f"(o.x), "f"(o.y),
                             "f"(o.z), "f"(o.w)
                             : "memory");
                if (POWER) {
                    float e0 = o.x, e1 = o.y, e2 = o.z, e3 = o.w;
                    if (MASK) {
                        const int r = rem - 4 * u;  // valid outputs of this unit
                        e0 = r > 0 ? e0 : 0.0f;
                        e1 = r > 1 ? e1 : 0.0f;
                        e2 = r > 2 ? e2 : 0.0f;
                        e3 = r > 3 ? e3 : 0.0f;
                    }
                    if (HEAD) {
                        head[0] = e0 * e0;
                        head[1] = e1 * e1;
                        head[2] = e2 * e2;
                        head[3] = e3 * e3;
                    } else {
                        body = fmaf(e0, e0, body);
                        body = fmaf(e1, e1, body);
                        body = fmaf(e2, e2, body);
                        body = fmaf(e3, e3, body);
                    }
                }
            };
            auto finish = [&](int u) { emit_unit(u, std::true_type{}); };        // u == 0
            auto finish_body = [&](int u) { emit_unit(u, std::false_type{}); };  // u >= 1
            auto advance = [&](const float4& raw2) {
                xin[0] = (double)raw1.x;
                xin[1] = (double)raw1.y;
                xin[2] = (double)raw1.z;
                xin[3] = (double)raw1.w;
                raw1 = raw2;
                tin += 4;
            };
            // units 0 and 1 are peeled (unit 0 has nothing to finish, the finish of unit 0 fills `head`), the other six
            // run two per trip: no register rotation and half the loop overhead (103 -> 93 instructions per unit)
            {
                const float4 raw2 = fetch(tin + 8);
                ticks();
                advance(raw2);
            }
            {
                const float4 raw2 = fetch(tin + 8);
                finish(0);
                ticks();
                advance(raw2);
            }
#pragma unroll 1
            for (int u = 2; u < 8; u += 2) {
                const float4 raw2 = fetch(tin + 8);
                finish_body(u - 1);
                ticks();
                advance(raw2);
                const float4 raw3 = fetch(tin + 8);
                finish_body(u);
                ticks();
                advance(raw3);
            }
            finish_body(7);
        };
        // warp-uniform choice (a per-lane branch would run the recurrence twice for a warp that holds the last chunk of a
        // clip): the masked variant whenever some row ends inside or before this block
        if (!emit)
            block(std::false_type{}, std::false_type{});
        else if (!__any_sync(0xffffffffu, rem < 32))
            block(std::true_type{}, std::false_type{});
        else
            block(std::true_type{}, std::true_type{});
        if (POWER && emit) {
            if (group_start) {
#pragma unroll
                for (int c = 0; c < 4; ++c) q[c] = head[c];
            } else {
                body += (head[0] + head[1]) + (head[2] + head[3]);
            }
            grp += body;
        }
        __syncwarp();
        // ---- store: rows 4*i + rsub, positions tb + c4 .. + 3
        if (emit) {
            const int t = tb + c4;
            const bool edge = __any_sync(0xffffffffu, edge_lo == (tb >> 5) || edge_hi == (tb >> 5));
            if (!edge) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int olo = max(mv[i].lo, b.W);
                    if (t >= olo && t < mv[i].hi)
                        *reinterpret_cast<float4*>(b.y32 + mv[i].base + t) =
                            *reinterpret_cast<const float4*>(&tile[4 * i + rsub][col0 + c4]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int olo = max(mv[i].lo, b.W);
                    const float4 v = *reinterpret_cast<const float4*>(&tile[4 * i + rsub][col0 + c4]);
                    const float* e = reinterpret_cast<const float*>(&v);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (t + c >= olo && t + c < mv[i].hi) b.y32[mv[i].base + t + c] = e[c];
                }
            }
            __syncwarp();
        }
        // ---- hand the column block to stream block tb + 32 * kRing
        if (tb + 32 * kRing < t_end) issue_load(tb + 32 * kRing);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (POWER && gidx * b.hop < valid_len) {  // the last group of the chunk
        float4* dst = reinterpret_cast<float4*>(b.group_energy + (group_base + gidx) * 8);
        dst[0] = make_float4(grp, q[0], q[1], q[2]);
        dst[1] = make_float4(q[3], 0.0f, 0.0f, 0.0f);
    }
}

// Same for the aligned-coordinate groups of iir_overlap4_kernel: hop block h of a clip whose first
// sample sits at alignment slot s is {first squares e >= s of group h} + {rest of group h} +
// {first squares e < s of group h+1}.
struct TrimHop4Batch {
    const float* group_energy;  // [n_groups][8] = {rest, q0, q1, q2, q3, -, -, -}
    const int64_t* clip_off;
    const int64_t* group_off;
    int64_t* start_end;
    int64_t n_clips;
    int hop, align;
    float top_db;
};

__global__ void __launch_bounds__(256) trim_index_hop4_kernel(const TrimHop4Batch b) {
    __shared__ float s_f[8];
    __shared__ int s_i[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float frame_length = (float)(2 * b.hop);
    for (int64_t clip = blockIdx.x; clip < b.n_clips; clip += gridDim.x) {
        const int64_t c0 = b.clip_off[clip];
        const int n = (int)(b.clip_off[clip + 1] - c0);
        const int s = (int)((c0 + b.align) & 3);
        const int T = 1 + n / b.hop;
        const int H = (n + b.hop - 1) / b.hop;
        const int G = (int)(b.group_off[clip + 1] - b.group_off[clip]);
        const float* ge = b.group_energy + b.group_off[clip] * 8;
        auto block_energy = [&](int h) {
            if (h < 0 || h >= H) return 0.0f;
            const float* a = ge + (size_t)h * 8;
            float e = 0.0f;
            for (int c = s; c < 4; ++c) e += a[1 + c];
            e += a[0];
            if (h + 1 < G)
                for (int c = 0; c < s; ++c) e += a[8 + 1 + c];
            return e;
        };
        auto power = [&](int t) {
            const float acc = block_energy(t - 1) + block_energy(t);
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

// zero-input transition of the cascade over C samples: column k = state after C steps from e_k
static void transition_matrix(const double* sos, int S, int C, std::vector<double>& M) {
    const int D = 2 * S;
    M.assign((size_t)D * D, 0.0);
    std::vector<double> s1(S), s2(S);
    for (int col = 0; col < D; ++col) {
        std::fill(s1.begin(), s1.end(), 0.0);
        std::fill(s2.begin(), s2.end(), 0.0);
        (col % 2 == 0 ? s1 : s2)[col / 2] = 1.0;
        for (int t = 0; t < C; ++t) {
            double v = 0.0;
            for (int k = 0; k < S; ++k) {
                const double* c = sos + 6 * k;  // b0 b1 b2 1 a1 a2 (normalised)
                const double y = fma(c[0], v, s1[k]);
                s1[k] = fma(c[1], v, fma(-c[4], y, s2[k]));
                s2[k] = fma(c[2], v, -c[5] * y);
                v = y;
            }
        }
        for (int k = 0; k < S; ++k) {
            M[(size_t)(2 * k) * D + col] = s1[k];
            M[(size_t)(2 * k + 1) * D + col] = s2[k];
        }
    }
}

template <int S>
static int run_iir(hmfe_ctx* ctx, IirBatch b, const double* sos, bool bp, double gain, cudaStream_t st) {
    IirCoef<S> cf;
    for (int k = 0; k < S; ++k) {
        cf.b0[k] = sos[6 * k + 0];
        cf.b1[k] = sos[6 * k + 1];
        cf.b2[k] = sos[6 * k + 2];
        cf.a1[k] = sos[6 * k + 4];
        cf.a2[k] = sos[6 * k + 5];
    }
    cf.bandpass_form = bp ? 1 : 0;
    cf.gain = gain;
    const unsigned grid = (unsigned)((b.n_chunks + kIirWarps * 32 - 1) / (kIirWarps * 32));
    ctx->prof_begin(HMFE_K_IIR_ZERO_STATE, st);
    if (bp)
        iir_chunk_kernel<S, false, true><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    else
        iir_chunk_kernel<S, false, false><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->prof_begin(HMFE_K_IIR_CARRY, st);
    iir_carry_kernel<S><<<(unsigned)((b.n_clips * 32 + 127) / 128), 128, 0, st>>>(b);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->prof_begin(HMFE_K_IIR_FINAL, st);
    if (bp)
        iir_chunk_kernel<S, true, true><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    else
        iir_chunk_kernel<S, true, false><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches += 3;
    return HMFE_OK;
}


// smallest multiple of 32 after which the zero-input response of the cascade has decayed below
// tol (infinity norm of the state transition), or -1 if that takes more than max_w samples
static int decay_length(const double* sos, int S, double tol, int max_w) {
    const int D = 2 * S;
    std::vector<double> st((size_t)D * D, 0.0);  // column c = state reached from unit vector e_c
    for (int c = 0; c < D; ++c) st[(size_t)c * D + c] = 1.0;
    for (int t = 1; t <= max_w; ++t) {
        for (int c = 0; c < D; ++c) {
            double* s = st.data() + (size_t)c * D;  // (s1[0], s2[0], s1[1], ...)
            double v = 0.0;
            for (int k = 0; k < S; ++k) {
                const double* q = sos + 6 * k;
                const double y = fma(q[0], v, s[2 * k]);
                s[2 * k] = fma(q[1], v, fma(-q[4], y, s[2 * k + 1]));
                s[2 * k + 1] = fma(q[2], v, -q[5] * y);
                v = y;
            }
        }
        if (t % 32 == 0) {
            double norm = 0.0;
            for (int r = 0; r < D; ++r) {
                double row = 0.0;
                for (int c = 0; c < D; ++c) row += fabs(st[(size_t)c * D + r]);
                norm = std::max(norm, row);
            }
            if (!(norm < 1e300)) return -1;  // unstable
            if (norm <= tol) return t;
        }
    }
    return -1;
}

template <int S, typename T, bool POWER>
static int launch_overlap(const IirOverlapBatch& b, const IirCoef<S>& cf, bool bp, cudaStream_t st) {
    const unsigned grid = (unsigned)((b.n_chunks + kIirWarps * 32 - 1) / (kIirWarps * 32));
    if (bp)
        iir_overlap_kernel<S, true, T, POWER><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    else
        iir_overlap_kernel<S, false, T, POWER><<<grid, kIirWarps * 32, 0, st>>>(b, cf);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

template <int S>
static int run_iir_overlap(hmfe_ctx* ctx, const IirOverlapBatch& b, const double* sos, bool bp, double gain, bool power,
                           cudaStream_t st) {
    IirCoef<S> cf;
    for (int k = 0; k < S; ++k) {
        cf.b0[k] = sos[6 * k + 0];
        cf.b1[k] = sos[6 * k + 1];
        cf.b2[k] = sos[6 * k + 2];
        cf.a1[k] = sos[6 * k + 4];
        cf.a2[k] = sos[6 * k + 5];
    }
    cf.bandpass_form = bp ? 1 : 0;
    cf.gain = gain;
    ctx->prof_begin(HMFE_K_IIR_OVERLAP, st);
    int rc;
    if (b.y64)
        rc = power ? launch_overlap<S, double, true>(b, cf, bp, st) : launch_overlap<S, double, false>(b, cf, bp, st);
    else
        rc = power ? launch_overlap<S, float, true>(b, cf, bp, st) : launch_overlap<S, float, false>(b, cf, bp, st);
    if (rc != HMFE_OK) return rc;
    ctx->prof_end(st);
    ctx->last_launches += 1;
    return HMFE_OK;
}

constexpr size_t kOverlap4Smem = (size_t)kIirWarps * 32 * kRow4 * sizeof(float) + (size_t)kIirWarps * 32 * sizeof(IirRow4);

// HMFE_IIR_GROUP: 32-sample blocks per load / store group of iir_overlap4_kernel (1 or 2; default 2).  Measured on B200,
// c2: 3.60 ms with 128-byte visits per row, 3.46 ms with 256-byte visits - the 57 000 concurrent row streams of this
// kernel touch a different DRAM page with every visit, and twice the bytes per visit is half the page activations.
static int iir_group_mode() {
    static const int mode = [] {
        const char* e = getenv("HMFE_IIR_GROUP");
        return e && atoi(e) == 1 ? 1 : 2;
    }();
    return mode;
}

template <int S, bool BP, bool POWER, int CONV, int GRP = 1>
static int launch_overlap4_c(const IirOverlap4Batch& b, const IirCoef<S>& cf, cudaStream_t st) {
    auto kern = iir_overlap4_kernel<S, BP, POWER, CONV, GRP>;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOverlap4Smem));
    const unsigned grid = (unsigned)((b.n_chunks + kIirWarps * 32 - 1) / (kIirWarps * 32));
    kern<<<grid, kIirWarps * 32, kOverlap4Smem, st>>>(b, cf);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

// float <-> double conversions of the one-pass kernel: 0 = conversion instructions, 1 = integer re-biasing on the way
// in, 3 = both ways (HMFE_IIR_CONV overrides; measured on B200)
[[maybe_unused]] static int iir_conv_mode() {
    static const int mode = [] {
        const char* e = getenv("HMFE_IIR_CONV");
        const int m = e ? atoi(e) : HMFE_IIR_CONV_DEFAULT;
        return (m == 1 || m == 3) ? m : 0;
    }();
    return mode;
}

// 0 (default): iir_overlap4_kernel; 1: iir_pipe4_kernel, the pipelined cascade (HMFE_IIR_PIPE overrides, for A/B runs).
// Measured on B200, c2 (1.78 G samples): 3.54 ms against 3.61 ms.  The pipelined form removed what ncu had blamed in the
// unrolled one (shared-memory scoreboard stalls on the first conversion of every float4: 9.5 % -> 0) and runs 29
// instead of 27.7 instructions per sample; both end at 51 % of the FP64 pipe, with ~35 % of the warps' time in the
// per-block load / store / hand-over code that neither form changes.  A third form moved every lane's row with one
// cp.async.bulk per block and direction (no cross-lane phase at all, 101 registers): 4.36 ms - the bulk-copy instruction
// takes its operands from uniform registers, so 32 different per-lane copies become a 32-trip ELECT / R2UR / UBLKCP
// loop per warp (ncu: 31.5 trips per block, 41 % of the stall samples fixed-latency waits in it); removed.
static int iir_pipe_mode() {
    static const int mode = [] {
        const char* e = getenv("HMFE_IIR_PIPE");
        return e ? atoi(e) : 0;
    }();
    return mode;
}

template <int S, bool BP, bool POWER>
static int launch_pipe4(const IirOverlap4Batch& b, const IirCoef<S>& cf, cudaStream_t st) {
    auto kern = iir_pipe4_kernel<S, BP, POWER>;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOverlap4Smem));
    const unsigned grid = (unsigned)((b.n_chunks + kIirWarps * 32 - 1) / (kIirWarps * 32));
    kern<<<grid, kIirWarps * 32, kOverlap4Smem, st>>>(b, cf);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

template <int S, bool BP, bool POWER>
static int launch_overlap4_k(const IirOverlap4Batch& b, const IirCoef<S>& cf, cudaStream_t st) {
    if (iir_pipe_mode() != 0) return launch_pipe4<S, BP, POWER>(b, cf, st);
#ifdef HMFE_IIR_BUILD_CONV_VARIANTS  // the integer-conversion A/B (3.58 -> 3.61 / 3.65 ms): 2/3 of this file's build time
    switch (iir_conv_mode()) {
        case 1: return launch_overlap4_c<S, BP, POWER, 1>(b, cf, st);
        case 3: return launch_overlap4_c<S, BP, POWER, 3>(b, cf, st);
        default: break;
    }
#endif
    if (iir_group_mode() == 2) return launch_overlap4_c<S, BP, POWER, 0, 2>(b, cf, st);
    return launch_overlap4_c<S, BP, POWER, 0>(b, cf, st);
}

template <int S>
static int run_iir_overlap4(hmfe_ctx* ctx, const IirOverlap4Batch& b, const double* sos, bool bp, double gain, bool power,
                            cudaStream_t st) {
    IirCoef<S> cf;
    for (int k = 0; k < S; ++k) {
        cf.b0[k] = sos[6 * k + 0];
        cf.b1[k] = sos[6 * k + 1];
        cf.b2[k] = sos[6 * k + 2];
        cf.a1[k] = sos[6 * k + 4];
        cf.a2[k] = sos[6 * k + 5];
    }
    cf.bandpass_form = bp ? 1 : 0;
    cf.gain = gain;
    ctx->prof_begin(HMFE_K_IIR_OVERLAP, st);
    int rc;
    if (bp)
        rc = power ? launch_overlap4_k<S, true, true>(b, cf, st) : launch_overlap4_k<S, true, false>(b, cf, st);
    else
        rc = power ? launch_overlap4_k<S, false, true>(b, cf, st) : launch_overlap4_k<S, false, false>(b, cf, st);
    if (rc != HMFE_OK) return rc;
    ctx->prof_end(st);
    ctx->last_launches += 1;
    return HMFE_OK;
}

struct TrimArgs {
    int frame_length, hop_length;
    float top_db;
    int64_t* d_start_end;
};

// Shared implementation of hmfe_iir_sos_batch (trim == nullptr) and hmfe_iir_sos_trim_batch.
static int iir_impl(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips, const double* h_sos,
                    int n_sections, float* d_y32, double* d_y64, const TrimArgs* trim, void* stream) {
    HMFE_REQUIRE(ctx && h_offsets && h_sos, "NULL argument");
    HMFE_REQUIRE(n_sections >= 1 && n_sections <= kIirMaxSections, "n_sections=%d not in [1, %d]", n_sections,
                 kIirMaxSections);
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    ctx->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_x && (d_y32 || d_y64), "NULL device pointer");
    if (trim) {
        HMFE_REQUIRE(d_y32 && trim->d_start_end, "the fused trim needs the float32 output and d_start_end");
        HMFE_REQUIRE(trim->frame_length >= 2 && trim->hop_length >= 1, "bad trim arguments");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = n_sections, D = 2 * S;
    std::vector<double> sos((size_t)6 * S);
    for (int k = 0; k < S; ++k) {
        const double a0 = h_sos[6 * k + 3];
        HMFE_REQUIRE(a0 != 0.0, "section %d has a0 == 0", k);
        for (int c = 0; c < 6; ++c) sos[6 * k + c] = h_sos[6 * k + c] / a0;
    }
    // band-pass form: every numerator is g_k * (1, 0, -1) -> fold the gains into one input gain and
    // run (and build the transition matrix for) the normalised cascade
    bool bp = true;
    double gain = 1.0;
    for (int k = 0; k < S; ++k) {
        bp = bp && sos[6 * k + 1] == 0.0 && sos[6 * k + 2] == -sos[6 * k + 0] && sos[6 * k + 0] != 0.0;
        gain *= sos[6 * k + 0];
    }
    if (bp)
        for (int k = 0; k < S; ++k) {
            sos[6 * k + 0] = 1.0;
            sos[6 * k + 1] = 0.0;
            sos[6 * k + 2] = -1.0;
        }
    else
        gain = 1.0;
    int64_t total = 0, max_len = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_offsets[i + 1] - h_offsets[i];
        HMFE_REQUIRE(n >= 0 && n < (int64_t)1 << 30, "clip %lld has invalid length %lld", (long long)i, (long long)n);
        total += n;
        max_len = std::max(max_len, n);
    }

    // ---- choose the algorithm: overlap (one pass, W warm-up samples per chunk) or exact scan
    const int C_exact = total >= ((int64_t)32 << 20) ? 512 : 128;
    if (ctx->iir_cache_S != S || memcmp(ctx->iir_cache_sos, sos.data(), sizeof(double) * 6 * S) != 0) {
        ctx->iir_cache_W = decay_length(sos.data(), S, 1e-13, 8192);
        ctx->iir_cache_S = S;
        memcpy(ctx->iir_cache_sos, sos.data(), sizeof(double) * 6 * S);
    }
    int W = ctx->iir_cache_W;
    const bool fuse_trim = trim && trim->frame_length == 2 * trim->hop_length && trim->hop_length % 32 == 0;
    const int hop = trim ? trim->hop_length : 1;
    // 128-bit rows need x and y32 on the same 16-byte phase and no float64 output
    const int align = (int)((reinterpret_cast<uintptr_t>(d_x) >> 2) & 3);
    const bool vec = d_y64 == nullptr && (reinterpret_cast<uintptr_t>(d_x) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(d_y32) & 3) == 0 &&
                     (int)((reinterpret_cast<uintptr_t>(d_y32) >> 2) & 3) == align && ctx->iir_rows != HMFE_IIR_ROWS_SCALAR;
    auto shift_of = [&](int64_t i) { return vec ? (int)((h_offsets[i] + align) & 3) : 0; };
    int unit = 800;
    if (vec) {
        W = W > 0 ? (W + 127) / 128 * 128 : W;
        unit = 128;
        if (fuse_trim) {
            int a = 128, bb = hop;  // lcm(128, hop)
            while (bb) {
                const int r = a % bb;
                a = bb;
                bb = r;
            }
            unit = 128 / a * hop;
        } else {
            unit = 3200;
        }
    } else if (fuse_trim) {
        unit = hop;
    }
    int C_overlap = 0;
    if (W > 0 && ctx->iir_algo != HMFE_IIR_ALGO_SCAN) {
        const double slots_o = (double)ctx->sm_count * (vec ? 3 : 4) * kIirWarps * 32;  // CTAs per SM by launch bounds
        const double slots_e = (double)ctx->sm_count * 6 * kIirWarps * 32;
        auto chunks_at = [&](int C, bool shifted) {
            int64_t c = 0;
            for (int64_t i = 0; i < n_clips; ++i)
                c += (h_offsets[i + 1] - h_offsets[i] + (shifted ? shift_of(i) : 0) + C - 1) / C;
            return c;
        };
        const double cost_exact = 2.2 * C_exact * ceil((double)chunks_at(C_exact, false) / slots_e) * slots_e;
        double best = 0.0;
        for (int m = 1; m <= 64; ++m) {
            const int C = unit * m;
            if (C < 512) continue;
            if (C > 32768) break;
            const double cost = (double)(W + C) * ceil((double)chunks_at(C, true) / slots_o) * slots_o;
            if (C_overlap == 0 || cost < best) {
                best = cost;
                C_overlap = C;
            }
        }
        if (ctx->iir_algo == HMFE_IIR_ALGO_AUTO && !(best < cost_exact)) C_overlap = 0;
    }
    HMFE_REQUIRE(C_overlap > 0 || ctx->iir_algo != HMFE_IIR_ALGO_OVERLAP,
                 "overlap IIR requested but the filter does not decay within 8192 samples");
    const bool overlap = C_overlap > 0;
    const int C = overlap ? C_overlap : C_exact;
    const bool power = overlap && fuse_trim;
    ctx->iir_last_algo = overlap ? HMFE_IIR_ALGO_OVERLAP : HMFE_IIR_ALGO_SCAN;
    ctx->iir_last_C = C;
    ctx->iir_last_W = overlap ? W : 0;
    ctx->iir_last_rows = overlap && vec ? HMFE_IIR_ROWS_VECTOR : HMFE_IIR_ROWS_SCALAR;

    std::vector<double> M;
    if (!overlap) transition_matrix(sos.data(), S, C, M);
    const size_t idx_bytes = 3 * (size_t)(n_clips + 1) * sizeof(int64_t);
    const size_t m_bytes = overlap ? 0 : (size_t)D * D * sizeof(double);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(idx_bytes + m_bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    int64_t* hc = static_cast<int64_t*>(hbuf);
    int64_t* hp = hc + (n_clips + 1);
    int64_t* hh = hp + (n_clips + 1);
    hp[0] = hh[0] = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_offsets[i + 1] - h_offsets[i] + (overlap ? shift_of(i) : 0);
        hc[i] = h_offsets[i];
        hp[i + 1] = hp[i] + (n + C - 1) / C;
        hh[i + 1] = hh[i] + (power ? (n + hop - 1) / hop : 0);
    }
    hc[n_clips] = h_offsets[n_clips];
    if (m_bytes) memcpy(static_cast<unsigned char*>(hbuf) + idx_bytes, M.data(), m_bytes);
    int rc = ctx->ring.upload(slot, idx_bytes + m_bytes, st);
    if (rc != HMFE_OK) return rc;
    const int64_t* d_clip_off = static_cast<int64_t*>(dbuf);
    const int64_t* d_chunk_prefix = d_clip_off + (n_clips + 1);
    const int64_t* d_hop_off = d_chunk_prefix + (n_clips + 1);
    const int64_t n_chunks = hp[n_clips];

    if (overlap && vec) {
        IirOverlap4Batch b{};
        b.x = d_x;
        b.y32 = d_y32;
        b.clip_off = d_clip_off;
        b.chunk_prefix = d_chunk_prefix;
        b.group_off = d_hop_off;
        b.n_clips = n_clips;
        b.n_chunks = n_chunks;
        b.C = C;
        b.W = W;
        b.hop = power ? hop : C;
        b.align = align;
        {
            static const int l2pf = [] {
                const char* e = getenv("HMFE_IIR_L2PF");  // default on: c2 3.48 -> 3.30 ms (256-byte visits), 3.60 -> 3.33 (128-byte)
                return e ? atoi(e) : 1;
            }();
            b.l2_prefetch = l2pf;
        }
        if (power) {
            rc = ctx->reserve_scratch((size_t)std::max<int64_t>(1, hh[n_clips]) * 8 * sizeof(float));
            if (rc != HMFE_OK) return rc;
            b.group_energy = static_cast<float*>(ctx->scratch);
        }
        if (n_chunks > 0) {
            switch (S) {
                case 1: rc = run_iir_overlap4<1>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 2: rc = run_iir_overlap4<2>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 3: rc = run_iir_overlap4<3>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 4: rc = run_iir_overlap4<4>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 5: rc = run_iir_overlap4<5>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 6: rc = run_iir_overlap4<6>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 7: rc = run_iir_overlap4<7>(ctx, b, sos.data(), bp, gain, power, st); break;
                default: rc = run_iir_overlap4<8>(ctx, b, sos.data(), bp, gain, power, st); break;
            }
            if (rc != HMFE_OK) return rc;
        }
        if (power) {
            TrimHop4Batch tb{b.group_energy, d_clip_off, d_hop_off, trim->d_start_end, n_clips, hop, align, trim->top_db};
            ctx->prof_begin(HMFE_K_TRIM_INDEX, st);
            trim_index_hop4_kernel<<<(int)std::min<int64_t>(n_clips, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(tb);
            HMFE_CHECK_CUDA(cudaGetLastError());
            ctx->prof_end(st);
            ctx->last_launches += 1;
        }
    } else if (overlap) {
        IirOverlapBatch b{};
        b.x = d_x;
        b.y32 = d_y32;
        b.y64 = d_y64;
        b.clip_off = d_clip_off;
        b.chunk_prefix = d_chunk_prefix;
        b.hop_off = d_hop_off;
        b.n_clips = n_clips;
        b.n_chunks = n_chunks;
        b.C = C;
        b.W = W;
        b.hop_steps = power ? hop / 32 : 1;
        if (power) {
            rc = ctx->reserve_scratch((size_t)std::max<int64_t>(1, hh[n_clips]) * sizeof(float));
            if (rc != HMFE_OK) return rc;
            b.hop_energy = static_cast<float*>(ctx->scratch);
        }
        if (n_chunks > 0) {
            switch (S) {
                case 1: rc = run_iir_overlap<1>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 2: rc = run_iir_overlap<2>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 3: rc = run_iir_overlap<3>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 4: rc = run_iir_overlap<4>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 5: rc = run_iir_overlap<5>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 6: rc = run_iir_overlap<6>(ctx, b, sos.data(), bp, gain, power, st); break;
                case 7: rc = run_iir_overlap<7>(ctx, b, sos.data(), bp, gain, power, st); break;
                default: rc = run_iir_overlap<8>(ctx, b, sos.data(), bp, gain, power, st); break;
            }
            if (rc != HMFE_OK) return rc;
        }
        if (power) {
            TrimHopBatch tb{b.hop_energy, d_clip_off, d_hop_off, trim->d_start_end, n_clips, hop, trim->top_db};
            ctx->prof_begin(HMFE_K_TRIM_INDEX, st);
            trim_index_hop_kernel<<<(int)std::min<int64_t>(n_clips, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(tb);
            HMFE_CHECK_CUDA(cudaGetLastError());
            ctx->prof_end(st);
            ctx->last_launches += 1;
        }
    } else {
        IirBatch b{};
        b.x = d_x;
        b.y32 = d_y32;
        b.y64 = d_y64;
        b.clip_off = d_clip_off;
        b.chunk_prefix = d_chunk_prefix;
        b.M = reinterpret_cast<const double*>(static_cast<const unsigned char*>(dbuf) + idx_bytes);
        b.n_clips = n_clips;
        b.n_chunks = n_chunks;
        b.C = C;
        if (n_chunks > 0) {
            rc = ctx->reserve_scratch(2 * (size_t)n_chunks * D * sizeof(double));
            if (rc != HMFE_OK) return rc;
            b.zstate = static_cast<double*>(ctx->scratch);
            b.init = b.zstate + n_chunks * D;
            switch (S) {
                case 1: rc = run_iir<1>(ctx, b, sos.data(), bp, gain, st); break;
                case 2: rc = run_iir<2>(ctx, b, sos.data(), bp, gain, st); break;
                case 3: rc = run_iir<3>(ctx, b, sos.data(), bp, gain, st); break;
                case 4: rc = run_iir<4>(ctx, b, sos.data(), bp, gain, st); break;
                case 5: rc = run_iir<5>(ctx, b, sos.data(), bp, gain, st); break;
                case 6: rc = run_iir<6>(ctx, b, sos.data(), bp, gain, st); break;
                case 7: rc = run_iir<7>(ctx, b, sos.data(), bp, gain, st); break;
                default: rc = run_iir<8>(ctx, b, sos.data(), bp, gain, st); break;
            }
            if (rc != HMFE_OK) return rc;
        }
    }
    rc = ctx->ring.release(slot, st);
    if (rc != HMFE_OK) return rc;
    if (trim && !power) {  // exact scan (or a hop the fused kernel does not cover): separate trim kernels
        const int launches = ctx->last_launches;
        rc = hmfe_trim_batch(ctx, d_y32, h_offsets, n_clips, trim->frame_length, trim->hop_length, trim->top_db,
                             trim->d_start_end, stream);
        ctx->last_launches += launches;
    }
    return rc;
}

}  // namespace hmfe

using namespace hmfe;

extern "C" int hmfe_iir_sos_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                                  const double* h_sos, int n_sections, float* d_y32, double* d_y64, void* stream) {
    return iir_impl(ctx, d_x, h_offsets, n_clips, h_sos, n_sections, d_y32, d_y64, nullptr, stream);
}

extern "C" int hmfe_iir_sos_trim_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                                       const double* h_sos, int n_sections, float* d_y32, double* d_y64,
                                       int frame_length, int hop_length, float top_db, int64_t* d_start_end,
                                       void* stream) {
    const TrimArgs t{frame_length, hop_length, top_db, d_start_end};
    return iir_impl(ctx, d_x, h_offsets, n_clips, h_sos, n_sections, d_y32, d_y64, &t, stream);
}

extern "C" int hmfe_ctx_set_iir_algo(hmfe_ctx* ctx, int algo) {
    HMFE_REQUIRE(ctx, "NULL ctx");
    HMFE_REQUIRE(algo >= HMFE_IIR_ALGO_AUTO && algo <= HMFE_IIR_ALGO_OVERLAP, "bad IIR algorithm id %d", algo);
    ctx->iir_algo = algo;
    return HMFE_OK;
}

extern "C" int hmfe_ctx_last_iir_plan(const hmfe_ctx* ctx, int* algo, int* chunk, int* warmup, int* rows) {
    HMFE_REQUIRE(ctx, "NULL ctx");
    if (algo) *algo = ctx->iir_last_algo;
    if (chunk) *chunk = ctx->iir_last_C;
    if (warmup) *warmup = ctx->iir_last_W;
    if (rows) *rows = ctx->iir_last_rows;
    return HMFE_OK;
}

extern "C" int hmfe_ctx_set_iir_rows(hmfe_ctx* ctx, int rows) {
    HMFE_REQUIRE(ctx, "NULL ctx");
    HMFE_REQUIRE(rows == HMFE_IIR_ROWS_AUTO || rows == HMFE_IIR_ROWS_SCALAR, "bad IIR row mode %d", rows);
    ctx->iir_rows = rows;
    return HMFE_OK;
}
