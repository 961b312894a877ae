// Kaldi-style log mel filterbank (Audio-MAE front-end).
// Replaces torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=16000,
// use_energy=False, window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10,
// frame_length=25) as called at /root/reference/src/util.py:845-856 and
// src/benchmark/baseline/extract_feature.py:232-243.
//
// Per frame (snip_edges): 400 samples at shift 160 -> subtract frame mean -> pre-emphasis 0.97
// (replicate pad) -> symmetric Hann -> zero pad to 512 -> |rFFT|^2 -> 128 HTK-mel triangles
// (20 Hz .. Nyquist, Nyquist column zero) -> log(max(., FLT_EPSILON)).
//
// A warp handles 4 consecutive frames as two complex 512-point FFTs (frame pairs packed as
// re/im), 512 = 32 (lanes) x 16 (registers): pass 1 is a packed (FFMA2) 16-point FFT over the
// two transforms, pass 2 one scalar 32-point FFT per lane (lane = transform*16 + k2).
// The per-chunk waveform mean subtraction the reference does before fbank cancels exactly
// against the per-frame DC removal and is not materialised.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "api_common.h"
#include "logmel_core.cuh"
#include "tables.h"

namespace hmfe {

constexpr int kFbPad = 512;
constexpr int kFbMaxSlots = 8;
constexpr int kFbPStride = 320;  // power tile stride per transform (>= 257 + banded slack)
constexpr int kFbWarps = 8;
constexpr int kFbMelGroup = 8;

struct FbMeta {
    int n_slots, total_trip, n_mels;
    int trip[kFbMaxSlots], wbase[kFbMaxSlots];
    int win, shift;
    float preemph;
    int flags;         // HMFE_FB_*
    float log_offset;  // HMFE_FB_LOG_OFFSET: log(x + log_offset) instead of log(max(x, FLT_EPSILON))
};

struct FbBatch {
    const float* wav;
    float* out;
    const int64_t* clip_start;   // [n_clips]
    const int64_t* clip_len;     // [n_clips]
    const int64_t* frame_off;    // [n_clips+1] output row offsets
    const int64_t* item_prefix;  // [n_clips+1]
    int64_t n_clips, n_items;
    int uniform_n, uniform_m, uniform_items, uniform_rows;
    int row_cap;  // > 0: frames beyond this row count are dropped (padded layout)
    unsigned long long* queue;  // next unclaimed work item (warps claim blocks of items dynamically)
};

struct FbTables {
    const float* win;   // [win] 0.5 * window
    const float2* tw;   // [16][32]
    const float* melw;  // [total_trip][32]
    const int* start;
    const int* row;
};

// Work item: 4 consecutive frames of one clip.
struct FbItem {
    const float* x;
    float* o;
    int nsamp, m, f0;
    bool valid;
};

HMFE_D FbItem fb_locate(const FbBatch& b, const FbMeta& mm, int64_t item, int64_t it_end, int64_t& clip) {
    FbItem c;
    c.valid = item < it_end;
    if (!c.valid) {
        c.x = b.wav;
        c.o = b.out;
        c.nsamp = c.m = c.f0 = 0;
        return c;
    }
    int64_t q;
    if (b.uniform_items > 0) {
        clip = item / b.uniform_items;
        q = item - clip * b.uniform_items;
        c.nsamp = b.uniform_n;
        c.m = b.uniform_m;
        c.x = b.wav + clip * (int64_t)c.nsamp;
        c.o = b.out + clip * (int64_t)b.uniform_rows * mm.n_mels;
    } else {
        if (clip < 0 || item >= b.item_prefix[min(clip + 4, b.n_clips)]) {
            int64_t lo = max(clip, (int64_t)0), hi = b.n_clips;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (b.item_prefix[mid] <= item)
                    lo = mid;
                else
                    hi = mid;
            }
            clip = lo;
        }
        while (item >= b.item_prefix[clip + 1]) ++clip;
        q = item - b.item_prefix[clip];
        c.nsamp = (int)b.clip_len[clip];
        c.m = c.nsamp >= mm.win ? 1 + (c.nsamp - mm.win) / mm.shift : 0;
        if (b.row_cap > 0) c.m = min(c.m, b.row_cap);
        c.x = b.wav + b.clip_start[clip];
        c.o = b.out + b.frame_off[clip] * mm.n_mels;
    }
    c.f0 = (int)q * 4;
    return c;
}

// FAST (25 ms / 10 ms at 16 kHz: win 400, shift 160 = 5 * 32): the 4 frames of an item cover 880
// consecutive samples, and sample n = lane + 32*n2 of frame t is span element lane + 32*(n2 + 5t):
// each lane reads its 28 span samples once (prefetched one item ahead) and every frame is built
// from those registers; the previous sample of the pre-emphasis comes from the neighbouring lane.
constexpr int kFbSpanRegs = 28;

constexpr float kLogFltEpsilon = -15.942384719848633f;  // float32(log(1.1920929e-7)), torch.log(eps) (kaldi.py: _get_log_energy / fbank floor)
// ln(x) for normal positive x
HMFE_D float fast_ln(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return 0.693147180559945309f * r;
}

HMFE_D void fb_load_span(const FbItem& c, int lane, float (&raw)[kFbSpanRegs]) {
    const int base = c.f0 * 160 + lane;
    if (c.valid && base - lane + 32 * kFbSpanRegs <= c.nsamp) {
        const float* p = c.x + base;
#pragma unroll
        for (int j = 0; j < kFbSpanRegs; ++j) raw[j] = __ldg(p + 32 * j);
    } else {
#pragma unroll
        for (int j = 0; j < kFbSpanRegs; ++j) {
            const int i = base + 32 * j;
            raw[j] = (c.valid && i < c.nsamp) ? __ldg(c.x + i) : 0.0f;
        }
    }
}

// CUSTOM = false: Kaldi's switches (DC removal, pre-emphasis, power, log floor) are compile-time facts;
// CUSTOM = true: they come from the plan (hmfe_fbank_plan_create_custom).
template <bool FAST, bool CUSTOM>
__global__ void __launch_bounds__(kFbWarps * 32, 2)
fbank_kernel(const FbBatch b, const FbTables tb, const FbMeta mm) {
    const bool remove_dc = CUSTOM ? (mm.flags & HMFE_FB_REMOVE_DC) != 0 : true;
    const bool use_preemph = CUSTOM ? mm.preemph != 0.0f : true;
    const bool magnitude = CUSTOM ? (mm.flags & HMFE_FB_MAGNITUDE) != 0 : false;
    const bool log_offset = CUSTOM ? (mm.flags & HMFE_FB_LOG_OFFSET) != 0 : false;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_tw = reinterpret_cast<float2*>(smem);                   // 512
    float* s_win = reinterpret_cast<float*>(s_tw + 512);              // 512 (zero padded)
    float* s_melw = s_win + 512;
    int* s_start = reinterpret_cast<int*>(s_melw + mm.total_trip * 32);
    int* s_row = s_start + mm.n_slots * 32;
    size_t tbytes = (size_t)(512 * 8 + 512 * 4 + mm.total_trip * 128 + mm.n_slots * 256);
    tbytes = (tbytes + 15) & ~(size_t)15;
    xelem<float>* tile = reinterpret_cast<xelem<float>*>(smem + tbytes) + (threadIdx.x >> 5) * (32 * kXStride);

    for (int i = threadIdx.x; i < 512; i += kFbWarps * 32) {
        s_tw[i] = tb.tw[i];
        s_win[i] = i < mm.win ? tb.win[i] : 0.0f;
    }
    for (int i = threadIdx.x; i < mm.total_trip * 32; i += kFbWarps * 32) s_melw[i] = tb.melw[i];
    for (int i = threadIdx.x; i < mm.n_slots * 32; i += kFbWarps * 32) {
        s_start[i] = tb.start[i];
        s_row[i] = tb.row[i];
    }
    // clear the tile's padding column once (read under zero mel weights, never written by the
    // exchange): stale NaN bit patterns in shared memory must not poison 0 * x
    tile[(threadIdx.x & 31) * kXStride + 32] = xelem<float>{};
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warps claim blocks of 8 consecutive items from a global counter (as in logmel.cu) and start staggered
    constexpr int kItemBlock = 8;
    const int64_t it_end = b.n_items;
    auto claim = [&]() -> int64_t {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(b.queue, (unsigned long long)kItemBlock);
        return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    int64_t blk_end = 0;
    auto next_item = [&](int64_t it) -> int64_t {
        if (it + 1 < blk_end) return it + 1;
        if (it >= it_end) return it;
        const int64_t nb = claim();
        blk_end = nb + kItemBlock;
        return nb;
    };
    int64_t clip = -1;
    __nanosleep((unsigned)(warp * 500));

    float raw[FAST ? kFbSpanRegs : 1];
    int64_t item = claim();
    blk_end = item + kItemBlock;
    FbItem cur = fb_locate(b, mm, item, it_end, clip);
    if constexpr (FAST) fb_load_span(cur, lane, raw);

    while (item < it_end) {
        const float* x = cur.x;
        float* o = cur.o;
        const int m = cur.m, f0 = cur.f0;

        // ---- 4 frames: DC removal, pre-emphasis, window; pack (A = f0,f0+1 | B = f0+2,f0+3)
        f32x2 re[16], im[16];
        if constexpr (FAST) {
            // d[j] = x[i] - preemph * x[i-1] over the span (frame independent), then per frame
            // (x[n] - mean) - preemph * (x[n-1] - mean) = d - (1 - preemph) * mean
            float mean[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (remove_dc) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float acc = 0.0f;
#pragma unroll
                    for (int n2 = 0; n2 < 13; ++n2)
                        if (n2 < 12 || lane < 16) acc += raw[n2 + 5 * t];
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    mean[t] = acc * (1.0f / 400.0f);  // within 1 ulp of sum / 400: far below the output tolerance
                }
            }
            const float x0[4] = {raw[0], raw[5], raw[10], raw[15]};  // first sample of each frame (lane 0)
            const float cm = 1.0f - mm.preemph;
            if (use_preemph) {
                float prev_row = 0.0f;  // raw[j-1] of this lane
#pragma unroll
                for (int j = 0; j < kFbSpanRegs; ++j) {
                    const float give = lane == 31 ? prev_row : raw[j];
                    const float pv = __shfl_sync(0xffffffffu, give, (lane + 31) & 31);
                    prev_row = raw[j];
                    raw[j] = fmaf(-mm.preemph, pv, raw[j]);
                }
            }
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                float y[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v = 0.0f;
                    if (n2 < 13) {
                        const float w = s_win[lane + 32 * n2];  // zero for n >= 400
                        v = fmaf(-cm, mean[t], raw[n2 + 5 * t]);
                        if (n2 == 0 && lane == 0) {  // replicate padding: x[-1] := x[0]
                            const float a = x0[t] - mean[t];
                            v = a - mm.preemph * a;
                        }
                        v *= w;
                    }
                    y[t] = v;
                }
                re[brev(n2, 4)] = f32x2{y[0], y[2]};
                im[brev(n2, 4)] = f32x2{y[1], y[3]};
            }
            // the span of the warp's next item is fetched while this item is transformed
            item = next_item(item);
            cur = fb_locate(b, mm, item, it_end, clip);
            fb_load_span(cur, lane, raw);
        } else {
            // pass 1 over the frame: mean (DC removal); pass 2 re-reads the samples (L1 hits)
            float mean[4];
            const float* xf[4];
            bool ok[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int f = f0 + t;
                ok[t] = f < m;
                xf[t] = x + (int64_t)(ok[t] ? f : 0) * mm.shift;
                float acc = 0.0f;
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    const int n = lane + 32 * n2;
                    if (ok[t] && n < mm.win) acc += __ldg(xf[t] + n);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                mean[t] = remove_dc ? acc / (float)mm.win : 0.0f;
            }
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int n = lane + 32 * n2;
                const float w = s_win[n];
                float y[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v = 0.0f;
                    if (ok[t] && n < mm.win) {
                        const float a = __ldg(xf[t] + n) - mean[t];
                        const float pv = __ldg(xf[t] + max(n - 1, 0)) - mean[t];
                        v = (a - mm.preemph * pv) * w;
                    }
                    y[t] = v;
                }
                re[brev(n2, 4)] = f32x2{y[0], y[2]};
                im[brev(n2, 4)] = f32x2{y[1], y[3]};
            }
            item = next_item(item);
            cur = fb_locate(b, mm, item, it_end, clip);
        }
        fft_dit<16, f32x2>(re, im);
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) {
            const float2 w = s_tw[k2 * 32 + lane];
            const f32x2 nr = vfnmas(im[k2], w.y, vmuls(re[k2], w.x));
            const f32x2 ni = vfmas(im[k2], w.x, vmuls(re[k2], w.y));
            re[k2] = nr;
            im[k2] = ni;
        }
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            tile[k2 * kXStride + lane] = xelem<float>{re[k2].x, im[k2].x};
            tile[(16 + k2) * kXStride + lane] = xelem<float>{re[k2].y, im[k2].y};
        }
        __syncwarp();
        float zr[32], zi[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const xelem<float> e = tile[lane * kXStride + n1];
            zr[brev(n1, 5)] = e.a;
            zi[brev(n1, 5)] = e.b;
        }
        __syncwarp();
        fft_dit<32, float>(zr, zi);  // lane = t*16 + k2 ; bin k = 16*k1 + k2

        {
            const int k2 = lane & 15;
            const int src = (lane & 16) | ((16 - k2) & 15);
            xelem<float>* pt = tile + (lane >> 4) * kFbPStride;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const float give_r = k2 == 0 ? zr[(32 - k1) & 31] : zr[31 - k1];
                const float give_i = k2 == 0 ? zi[(32 - k1) & 31] : zi[31 - k1];
                const float pr = __shfl_sync(0xffffffffu, give_r, src), pi = __shfl_sync(0xffffffffu, give_i, src);
                xelem<float> pw = frame_powers<float>(zr[k1], zi[k1], pr, pi);
                if (magnitude) {
                    pw.a = sqrtf(pw.a);
                    pw.b = sqrtf(pw.b);
                }
                pt[16 * k1 + k2] = pw;
            }
            if (k2 == 0) {
                xelem<float> pw = frame_powers<float>(zr[16], zi[16], zr[16], zi[16]);
                if (magnitude) {
                    pw.a = sqrtf(pw.a);
                    pw.b = sqrtf(pw.b);
                }
                pt[256] = pw;
            }
        }
        __syncwarp();

        for (int s = 0; s < mm.n_slots; ++s) {
            const int start = s_start[s * 32 + lane], trip = mm.trip[s];
            const float* wl = s_melw + mm.wbase[s] * 32 + lane;
            const xelem<float>* pa = tile + start;
            const xelem<float>* pb = tile + kFbPStride + start;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int i = 0; i < trip; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float w0 = wl[(i + j) * 32], w1 = wl[(i + j + 1) * 32];
                    const xelem<float> ea = pa[i + j], eb = pb[i + j], fa = pa[i + j + 1], fb = pb[i + j + 1];
                    a0 = fmaf(ea.a, w0, a0);
                    a1 = fmaf(ea.b, w0, a1);
                    a2 = fmaf(eb.a, w0, a2);
                    a3 = fmaf(eb.b, w0, a3);
                    c0 = fmaf(fa.a, w1, c0);
                    c1 = fmaf(fa.b, w1, c1);
                    c2 = fmaf(fb.a, w1, c2);
                    c3 = fmaf(fb.b, w1, c3);
                }
            }
            a0 += c0;
            a1 += c1;
            a2 += c2;
            a3 += c3;
            const int row = s_row[s * 32 + lane];
            if (row >= 0) {
                const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (f0 + t < m) {
                        // natural log as ln2 * lg2.approx on the special-function unit (~1e-6 nat from the rounded value,
                        // budget 2.3e-3 nat = 1e-2 dB; logf was 13 % of the kernel's instructions).  Entries at the floor
                        // are written as the constant float32 log(FLT_EPSILON) the reference produces, bit for bit.
                        const float x = log_offset ? acc[t] + mm.log_offset : acc[t];
                        const float v = (!log_offset && x <= FLT_EPSILON) ? kLogFltEpsilon : fast_ln(x);
                        o[(int64_t)(f0 + t) * mm.n_mels + row] = v;
                    }
            }
        }
        __syncwarp();
    }
}

// rows_per_clip layout: rows [m_i, rows_per_clip) of every clip are zero (the model-side pad);
// only those rows are cleared, the frames themselves are written once by fbank_kernel.
__global__ void __launch_bounds__(256) fbank_zero_pad_kernel(const FbBatch b, int n_mels, int win, int shift) {
    for (int64_t clip = blockIdx.x; clip < b.n_clips; clip += gridDim.x) {
        int m;
        float* o;
        if (b.uniform_items > 0) {
            m = b.uniform_m;
            o = b.out + clip * (int64_t)b.uniform_rows * n_mels;
        } else {
            const int nsamp = (int)b.clip_len[clip];
            m = min(nsamp >= win ? 1 + (nsamp - win) / shift : 0, b.row_cap);
            o = b.out + b.frame_off[clip] * n_mels;
        }
        float4* z = reinterpret_cast<float4*>(o + (int64_t)m * n_mels);
        const int64_t n4 = (int64_t)(b.row_cap - m) * n_mels / 4;
        for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) z[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

}  // namespace hmfe

using namespace hmfe;

struct hmfe_fbank_plan {
    int sample_rate, win, shift, n_mels;
    std::vector<float> mel_dense;
    FbMeta meta;
    float *d_win = nullptr, *d_melw = nullptr;
    float2* d_tw = nullptr;
    int *d_start = nullptr, *d_row = nullptr;
    size_t table_smem = 0;
    DescRing ring;
    int last_launches = 0, sm_count = 148;
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;  // pairs around fbank_kernel
};

template <typename T>
static int fb_upload(const std::vector<T>& v, T** dptr) {
    HMFE_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), std::max<size_t>(1, v.size()) * sizeof(T)));
    HMFE_CHECK_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return HMFE_OK;
}

extern "C" {

void hmfe_fbank_plan_destroy(hmfe_fbank_plan* p) {
    if (!p) return;
    cudaFree(p->d_win);
    cudaFree(p->d_tw);
    cudaFree(p->d_melw);
    cudaFree(p->d_start);
    cudaFree(p->d_row);
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    delete p;
}

static int fbank_plan_init(hmfe_fbank_plan** plan, int sample_rate, int win, int shift, int n_mels,
                           const std::vector<float>& half_window, std::vector<float> mel_dense, int flags, double preemph,
                           double log_offset) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    if (win <= 256 || win > kFbPad || shift < 1) {
        set_error("window of %d samples unsupported: the kernel is specialised for a 512-point padded FFT "
                  "(25 ms at 16 kHz, src/util.py:845-856)", win);
        return HMFE_ERR_UNSUPPORTED;
    }
    HMFE_REQUIRE(n_mels >= 32 && n_mels % 32 == 0 && n_mels <= 32 * kFbMaxSlots, "n_mels=%d must be a multiple of 32 <= %d",
                 n_mels, 32 * kFbMaxSlots);
    hmfe_fbank_plan* p = new (std::nothrow) hmfe_fbank_plan();
    HMFE_REQUIRE(p != nullptr, "out of host memory");
    p->sample_rate = sample_rate;
    p->win = win;
    p->shift = shift;
    p->n_mels = n_mels;
    p->sm_count = device_sm_count();
    p->mel_dense = std::move(mel_dense);
    // Alignment of every band's window start to its lane (mod group): 16 makes the 8-byte tile reads conflict free but
    // pads the 128 Kaldi bands (<= 10 bins wide) to 72 trips; 8 needs 40 trips and accepts two-way conflicts between
    // lanes l and l + 8 (measured on c3: 16 -> 0.789 ms, 8 -> 0.745 ms, 4 -> 0.787 ms).  HMFE_FBANK_MEL_GROUP overrides for measurements.
    int group = kFbMelGroup;
    if (const char* e = getenv("HMFE_FBANK_MEL_GROUP")) {
        const int g = atoi(e);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16) group = g;
    }
    // ... and the lanes are then re-assigned (minimum-cost assignment, tables.h) so that the windows are congruent to their
    // lane modulo 16 wherever that fits the same trip counts: for the Kaldi bank all of them do (40 trips, conflict free).
    int prefer = 16;
    if (const char* e = getenv("HMFE_FBANK_MEL_PREFER")) prefer = atoi(e);
    const BandedMel bm = build_banded(p->mel_dense, n_mels, kFbPad / 2 + 1, group, kFbPStride, prefer);
    if (!verify_banded(bm, p->mel_dense, kFbPStride)) {
        set_error("the mel matrix is not banded (every row must have one contiguous support that fits the tile)");
        delete p;
        return HMFE_ERR_INVALID;
    }
    p->meta.n_slots = bm.n_slots;
    p->meta.total_trip = bm.total_trip;
    p->meta.n_mels = n_mels;
    p->meta.win = win;
    p->meta.shift = shift;
    p->meta.preemph = (float)preemph;
    p->meta.flags = flags;
    p->meta.log_offset = (float)log_offset;
    for (int s = 0; s < bm.n_slots; ++s) {
        p->meta.trip[s] = bm.trip[s];
        p->meta.wbase[s] = bm.wbase[s];
    }
    int rc = fb_upload(half_window, &p->d_win);
    if (rc == HMFE_OK) rc = fb_upload(twiddle_plane(kFbPad, 16), reinterpret_cast<float**>(&p->d_tw));
    if (rc == HMFE_OK) rc = fb_upload(bm.w, &p->d_melw);
    if (rc == HMFE_OK) rc = fb_upload(bm.start, &p->d_start);
    if (rc == HMFE_OK) rc = fb_upload(bm.row, &p->d_row);
    if (rc != HMFE_OK) {
        hmfe_fbank_plan_destroy(p);
        return rc;
    }
    size_t tbytes = (size_t)(512 * 8 + 512 * 4 + bm.total_trip * 128 + bm.n_slots * 256);
    p->table_smem = (tbytes + 15) & ~(size_t)15;
    *plan = p;
    return HMFE_OK;
}

int hmfe_fbank_plan_create(hmfe_fbank_plan** plan, int sample_rate, double frame_length_ms, double frame_shift_ms,
                           int n_mels, double low_freq, double high_freq, double preemph) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    HMFE_REQUIRE(sample_rate > 0 && frame_length_ms > 0 && frame_shift_ms > 0, "bad frame parameters");
    const int win = (int)(sample_rate * frame_length_ms * 0.001), shift = (int)(sample_rate * frame_shift_ms * 0.001);
    if (win <= 256 || win > kFbPad || shift < 1 || n_mels < 32 || n_mels % 32 || n_mels > 32 * kFbMaxSlots)
        return fbank_plan_init(plan, sample_rate, win, shift, n_mels, {}, {}, 0, 0.0, 0.0);  // reports the error
    return fbank_plan_init(plan, sample_rate, win, shift, n_mels, half_hann_symmetric(win),
                           mel_banks_kaldi(n_mels, kFbPad, sample_rate, low_freq, high_freq), HMFE_FB_REMOVE_DC, preemph, 0.0);
}

int hmfe_fbank_plan_create_custom(hmfe_fbank_plan** plan, int sample_rate, int win, int shift, int n_mels,
                                  const float* h_window, const float* h_mel, int flags, double preemph, double log_offset) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    HMFE_REQUIRE(h_window && h_mel, "NULL argument");
    HMFE_REQUIRE((flags & ~(HMFE_FB_REMOVE_DC | HMFE_FB_MAGNITUDE | HMFE_FB_LOG_OFFSET)) == 0, "unknown flag bits 0x%x", flags);
    if (win <= 256 || win > kFbPad || shift < 1 || n_mels < 32 || n_mels % 32 || n_mels > 32 * kFbMaxSlots)
        return fbank_plan_init(plan, sample_rate, win, shift, n_mels, {}, {}, 0, 0.0, 0.0);  // reports the error
    std::vector<float> hw((size_t)win);
    for (int i = 0; i < win; ++i) hw[i] = 0.5f * h_window[i];  // 1/4 of the two-frames-per-FFT separation, exact
    std::vector<float> mel(h_mel, h_mel + (size_t)n_mels * (kFbPad / 2 + 1));
    return fbank_plan_init(plan, sample_rate, win, shift, n_mels, hw, std::move(mel), flags, preemph, log_offset);
}

int64_t hmfe_fbank_num_frames(const hmfe_fbank_plan* p, int64_t n_samples) {
    if (!p || n_samples < 0) return -1;
    return n_samples >= p->win ? 1 + (n_samples - p->win) / p->shift : 0;
}

int hmfe_fbank_mel_basis(const hmfe_fbank_plan* p, float* h_out) {
    HMFE_REQUIRE(p && h_out, "NULL argument");
    std::copy(p->mel_dense.begin(), p->mel_dense.end(), h_out);
    return HMFE_OK;
}

int hmfe_fbank_last_launches(const hmfe_fbank_plan* p) { return p ? p->last_launches : 0; }

int hmfe_fbank_set_profile(hmfe_fbank_plan* p, int enable) {
    HMFE_REQUIRE(p, "NULL plan");
    p->profile = enable != 0;
    return HMFE_OK;
}

int hmfe_fbank_profile_ms(hmfe_fbank_plan* p, double* kernel_ms, int* n_calls) {
    HMFE_REQUIRE(p, "NULL plan");
    double a = 0;
    const int n = (int)(p->prof_events.size() / 2);
    for (int i = 0; i < n; ++i) {
        float t = 0;
        HMFE_CHECK_CUDA(cudaEventSynchronize(p->prof_events[2 * i + 1]));
        HMFE_CHECK_CUDA(cudaEventElapsedTime(&t, p->prof_events[2 * i], p->prof_events[2 * i + 1]));
        a += t;
    }
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    p->prof_events.clear();
    if (kernel_ms) *kernel_ms = a;
    if (n_calls) *n_calls = n;
    return HMFE_OK;
}

// rows_per_clip == 0: clips' frames are packed back to back ([sum m_i, n_mels]);
// rows_per_clip  > 0: clip i owns rows [i*rows_per_clip, (i+1)*rows_per_clip), frames beyond
// rows_per_clip are dropped and missing rows are zero (the model-side pad to 1024 rows,
// audioMAE/models_mae.py:1178-1181).
int hmfe_fbank_batch_views(hmfe_fbank_plan* p, const float* d_wav, const int64_t* h_starts, const int64_t* h_lengths,
                           int64_t n_clips, float* d_out, int rows_per_clip, void* stream) {
    HMFE_REQUIRE(p && h_starts && h_lengths, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0 && rows_per_clip >= 0, "bad arguments");
    p->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_wav && d_out, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool uniform = true;
    const int64_t n0 = h_lengths[0];
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_lengths[i];
        HMFE_REQUIRE(n >= 0 && n < (int64_t)1 << 30 && h_starts[i] >= 0, "clip %lld has invalid start/length", (long long)i);
        uniform = uniform && n == n0 && h_starts[i] == i * n0;
    }
    FbBatch b{};
    b.wav = d_wav;
    b.out = d_out;
    b.n_clips = n_clips;
    b.row_cap = rows_per_clip;
    auto frames_of = [&](int64_t n) {
        int64_t m = n >= p->win ? 1 + (n - p->win) / p->shift : 0;
        return rows_per_clip > 0 ? std::min<int64_t>(m, rows_per_clip) : m;
    };
    const size_t desc_bytes = uniform ? 0 : (4 * (size_t)n_clips + 2) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = p->ring.acquire(desc_bytes + 16, &hbuf, &dbuf);
    if (slot < 0) return slot;
    b.queue = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(dbuf) + desc_bytes);
    HMFE_CHECK_CUDA(cudaMemsetAsync(b.queue, 0, sizeof(unsigned long long), st));
    int64_t total_rows = 0;
    if (uniform) {
        const int64_t m = frames_of(n0);
        b.uniform_n = (int)n0;
        b.uniform_m = (int)m;
        b.uniform_items = (int)std::max<int64_t>(1, (m + 3) / 4);
        b.uniform_rows = rows_per_clip > 0 ? rows_per_clip : (int)m;
        b.n_items = m > 0 ? (int64_t)b.uniform_items * n_clips : 0;
        total_rows = (int64_t)b.uniform_rows * n_clips;
    } else {
        int64_t* hs = static_cast<int64_t*>(hbuf);
        int64_t* hl = hs + n_clips;
        int64_t* hf = hl + n_clips;
        int64_t* hi = hf + (n_clips + 1);
        hf[0] = hi[0] = 0;
        for (int64_t i = 0; i < n_clips; ++i) {
            const int64_t m = frames_of(h_lengths[i]);
            hs[i] = h_starts[i];
            hl[i] = h_lengths[i];
            hf[i + 1] = hf[i] + (rows_per_clip > 0 ? rows_per_clip : m);
            hi[i + 1] = hi[i] + (m + 3) / 4;
        }
        b.n_items = hi[n_clips];
        total_rows = hf[n_clips];
        int64_t* dc = static_cast<int64_t*>(dbuf);
        b.clip_start = dc;
        b.clip_len = dc + n_clips;
        b.frame_off = dc + 2 * n_clips;
        b.item_prefix = dc + 3 * n_clips + 1;
        int rc = p->ring.upload(slot, desc_bytes, st);
        if (rc != HMFE_OK) return rc;
    }
    if (rows_per_clip > 0 && total_rows > 0) {
        HMFE_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "d_out must be 16-byte aligned");
        fbank_zero_pad_kernel<<<(int)std::min<int64_t>(n_clips, (int64_t)p->sm_count * 8), 256, 0, st>>>(b, p->n_mels, p->win,
                                                                                                       p->shift);
        HMFE_CHECK_CUDA(cudaGetLastError());
        p->last_launches += 1;
    }
    if (b.n_items > 0) {
        FbMeta mm = p->meta;
        const size_t smem = p->table_smem + (size_t)kFbWarps * 32 * kXStride * sizeof(xelem<float>);
        const bool fast = p->win == 400 && p->shift == 160;
        const bool custom = p->meta.flags != HMFE_FB_REMOVE_DC || p->meta.preemph == 0.0f;
        auto kern = fast ? (custom ? fbank_kernel<true, true> : fbank_kernel<true, false>)
                         : (custom ? fbank_kernel<false, true> : fbank_kernel<false, false>);
        HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t want = (b.n_items + kFbWarps - 1) / kFbWarps;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count * 2));
        FbTables tb{p->d_win, p->d_tw, p->d_melw, p->d_start, p->d_row};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        if (p->profile) {
            for (int i = 0; i < 2; ++i) {
                HMFE_CHECK_CUDA(cudaEventCreate(&ev[i]));
                p->prof_events.push_back(ev[i]);
            }
            HMFE_CHECK_CUDA(cudaEventRecord(ev[0], st));
        }
        kern<<<grid, kFbWarps * 32, smem, st>>>(b, tb, mm);
        HMFE_CHECK_CUDA(cudaGetLastError());
        if (p->profile) HMFE_CHECK_CUDA(cudaEventRecord(ev[1], st));
        p->last_launches += 1;
    }
    return p->ring.release(slot, st);
}

int hmfe_fbank_batch(hmfe_fbank_plan* p, const float* d_wav, const int64_t* h_offsets, int64_t n_clips, float* d_out,
                     int rows_per_clip, void* stream) {
    HMFE_REQUIRE(p && h_offsets, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    std::vector<int64_t> len((size_t)n_clips);
    for (int64_t i = 0; i < n_clips; ++i) len[i] = h_offsets[i + 1] - h_offsets[i];
    return hmfe_fbank_batch_views(p, d_wav, h_offsets, len.data(), n_clips, d_out, rows_per_clip, stream);
}

}  // extern "C"
