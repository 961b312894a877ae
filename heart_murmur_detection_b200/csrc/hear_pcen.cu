// HeAR mel-PCEN front-end (sibling front-ends): replaces preprocess_audio
// (/root/reference/src/benchmark/baseline/hear/python/data_processing/audio_utils.py:448-476) =
//   :361-365  batch-wide min / max scaling of the audio to [-1, 1]
//   :367-378  STFT, 400-sample periodic Hann frames every 160 samples, fft_length 400, zero padded at the end
//   :379-382  |X|^2 @ [201, 128] HTK mel matrix (:264-358)
//   :383      PCEN (:193-246): EMA smoother (:121-190, coefficient 0.04, initial state = first frame),
//             (x / (1e-8 + ema)^0.8 + 2)^(1/2) - 2^(1/2)
//   :475      bilinear resize (align_corners=False) of the [200, 128] image to [192, 128]
// Three kernels: (1) batch min / max, (2) scaling + framing + window + 400-point FFT + power + banded mel
// (hear_core.cuh), (3) PCEN recursion fused with the row interpolation.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "api_common.h"
#include "hear_core.cuh"
#include "tables.h"

namespace hmfe {

constexpr int kHearWarps = 8;
constexpr int kHearMaxSlots = 4;
constexpr int kHearSpan = 3 * kHearShift + kHearN;        // samples under the 4 frames of an item
constexpr int kHearSpanPad = kHearSpan + 16 * (kHearSpan / 320 + 1);
static_assert(kHearSpanPad * 4 <= kHearPRows * 16, "the sample span overlays the power tile");
constexpr size_t kHearWarpBytes = ((size_t)kHearPRows * 16 + (size_t)2 * kHearXTile * 8 + 15) & ~(size_t)15;

struct HearMeta {
    int n_slots, total_trip, n_mels;
    int trip[kHearMaxSlots], wbase[kHearMaxSlots];
};

struct HearBatch {
    const float* audio;  // [n_clips][n_samples]
    float* mel;          // [n_clips][T][n_mels] mel power
    unsigned* mm;        // ordered-integer codes of the batch minimum and maximum
    int64_t n_clips, n_items;
    int n_samples, n_padded, T, items_per_clip;
};

struct HearTables {
    const float* win;     // [400] 0.5 * window
    const float2* plane;  // [25][16] (cos, -sin)(2 pi n2 k1 / 400)
    const float* melw;    // [total_trip][32]
    const int* start;     // [n_slots][32]
    const int* row;       // [n_slots][32]
};

// order-preserving map float -> unsigned (so atomicMin / atomicMax on the codes order like the floats)
HMFE_HD unsigned f2ord(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
HMFE_HD float ord2f(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

__global__ void hear_minmax_init_kernel(unsigned* mm, int with_zero) {
    // samples the reference pads with zeros before scaling (audio_utils.py:466-468) take part in min / max
    mm[0] = with_zero ? f2ord(0.0f) : 0xffffffffu;
    mm[1] = with_zero ? f2ord(0.0f) : 0u;
}

__global__ void __launch_bounds__(256) hear_minmax_kernel(const float* __restrict__ x, int64_t n, unsigned* mm) {
    float lo = INFINITY, hi = -INFINITY;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int64_t n4 = n >> 2;
        for (int64_t j = i; j < n4; j += stride) {
            const float4 v = __ldg(x4 + j);
            lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
            hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
        i += n4 * 4;
    }
    for (; i < n; i += stride) {
        const float v = __ldg(x + i);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(mm, f2ord(lo));
        atomicMax(mm + 1, f2ord(hi));
    }
}

// physical position of span sample i: 16 extra floats per 320 samples keep the two half-warps of pass 1
// (which read 320 samples apart) on different shared-memory banks
HMFE_HD int hear_skew(int i) { return i + 16 * (i / 320); }

__global__ void __launch_bounds__(kHearWarps * 32)
hear_mel_kernel(const HearBatch b, const HearTables tb, const HearMeta mm) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_plane = reinterpret_cast<float2*>(smem);                   // 400
    float* s_win = reinterpret_cast<float*>(s_plane + 400);             // 400
    float* s_melw = s_win + 400;                                        // total_trip * 32
    int* s_start = reinterpret_cast<int*>(s_melw + mm.total_trip * 32);  // n_slots * 32
    int* s_row = s_start + mm.n_slots * 32;
    size_t tbytes = (size_t)(400 * 8 + 400 * 4 + mm.total_trip * 128 + mm.n_slots * 256);
    tbytes = (tbytes + 15) & ~(size_t)15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + tbytes + (size_t)warp * kHearWarpBytes;
    xelem<f32x2>* ptile = reinterpret_cast<xelem<f32x2>*>(wbase);
    float2* tile = reinterpret_cast<float2*>(wbase + (size_t)kHearPRows * 16);
    // the scaled samples of an item are dead once pass 1 has read them: they share the power tile's memory
    float* span = reinterpret_cast<float*>(ptile);

    for (int i = threadIdx.x; i < 400; i += kHearWarps * 32) {
        s_plane[i] = tb.plane[i];
        s_win[i] = tb.win[i];
    }
    for (int i = threadIdx.x; i < mm.total_trip * 32; i += kHearWarps * 32) s_melw[i] = tb.melw[i];
    for (int i = threadIdx.x; i < mm.n_slots * 32; i += kHearWarps * 32) {
        s_start[i] = tb.start[i];
        s_row[i] = tb.row[i];
    }
    __syncthreads();

    // scaling of audio_utils.py:361-365 in float32, operation by operation
    const HearScale sc = hear_make_scale(ord2f(b.mm[0]), ord2f(b.mm[1]));
    auto scale = [&](float x) { return hear_scale(sc, x); };
    const float scaled_zero = scale(0.0f);
    float* pf = reinterpret_cast<float*>(ptile);
    const int n_mels = mm.n_mels;

    // The samples under an item (880 = 27.5 rows of 32) are fetched one item ahead into registers, so the global
    // loads of item i+1 are in flight while item i is transformed.
    constexpr int kRows = (kHearSpan + 31) / 32;
    float raw[kRows];
    auto load_raw = [&](int64_t it) {
        const bool ok = it < b.n_items;
        const int64_t cl = ok ? it / b.items_per_clip : 0;
        const int s0n = ok ? (int)(it - cl * b.items_per_clip) * 4 * kHearShift : 0;
        const float* xs = b.audio + cl * (int64_t)b.n_samples + s0n;
        const int avail = ok ? b.n_samples - s0n : 0;  // samples of the clip from s0 on
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
            const int i = lane + 32 * j;
            raw[j] = i < avail ? __ldg(xs + i) : 0.0f;
        }
    };
    const int64_t item_stride = (int64_t)gridDim.x * kHearWarps;
    int64_t item = (int64_t)blockIdx.x * kHearWarps + warp;
    load_raw(item);
    for (; item < b.n_items; item += item_stride) {
        const int64_t clip = item / b.items_per_clip;
        const int f0 = (int)(item - clip * b.items_per_clip) * 4;
        const int s0 = f0 * kHearShift;
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
            const int i = lane + 32 * j, idx = s0 + i;
            // inside the clip: scaled sample; zero padding of preprocess_audio (:466-468): scaled zero;
            // zero padding of the STFT (:83-85, after the scaling): literal zero
            const float v = idx < b.n_samples ? scale(raw[j]) : (idx < b.n_padded ? scaled_zero : 0.0f);
            if (i < kHearSpan) span[hear_skew(i)] = v;
        }
        __syncwarp();
        auto fetch = [&](int tr, bool second, int n) -> float {
            const int t = 2 * tr + (second ? 1 : 0);
            return f0 + t < b.T ? span[hear_skew(kHearShift * t + n)] : 0.0f;
        };
        hear_pass1(lane, s_win, s_plane, fetch, tile);
        __syncwarp();
        load_raw(item + item_stride);
        // power tile rows above bin 200 are read under zero weights and never written by the separation: clear
        // what the span (or another kernel) left there, so that 0 * x cannot be NaN
        for (int i = kHearBins + lane; i < kHearPRows; i += 32) ptile[i] = xelem<f32x2>{};
        const int k1 = min(lane, 24);
        const int src = lane < 25 ? (25 - lane) % 25 : lane;
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            float zr[16], zi[16];
            hear_pass2(k1, tile, r, zr, zi);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const float gr = lane == 0 ? zr[hear_give_reg(true, k2)] : zr[hear_give_reg(false, k2)];
                const float gi = lane == 0 ? zi[hear_give_reg(true, k2)] : zi[hear_give_reg(false, k2)];
                const float pr = __shfl_sync(0xffffffffu, gr, src), pi = __shfl_sync(0xffffffffu, gi, src);
                const xelem<float> pw = frame_powers<float>(zr[k2], zi[k2], pr, pi);
                // power tile element k = (frame f0, f0+1 | f0+2, f0+3): transform r owns one 8-byte half
                if (lane < 25) *reinterpret_cast<float2*>(pf + 4 * (lane + 25 * k2) + 2 * r) = make_float2(pw.a, pw.b);
            }
            if (lane == 0) {
                const xelem<float> pw = frame_powers<float>(zr[8], zi[8], zr[8], zi[8]);
                *reinterpret_cast<float2*>(pf + 4 * 200 + 2 * r) = make_float2(pw.a, pw.b);
            }
        }
        __syncwarp();
        float* o = b.mel + (clip * (int64_t)b.T + f0) * n_mels;
        for (int s = 0; s < mm.n_slots; ++s) {
            f32x2 aa, ab;
            mel_slot<f32x2>(lane, ptile, s_melw + mm.wbase[s] * 32, s_start[s * 32 + lane], mm.trip[s], aa, ab);
            const int row = s_row[s * 32 + lane];
            if (row >= 0) {
                if (f0 < b.T) o[row] = aa.x;
                if (f0 + 1 < b.T) o[n_mels + row] = aa.y;
                if (f0 + 2 < b.T) o[2 * n_mels + row] = ab.x;
                if (f0 + 3 < b.T) o[3 * n_mels + row] = ab.y;
            }
        }
        __syncwarp();
    }
}

// One thread per (clip, mel channel): EMA + PCEN down the frames, emitting the bilinear rows as soon as both
// source rows exist (hear_pcen_column, hear_core.cuh).
__global__ void __launch_bounds__(128)
hear_pcen_resize_kernel(const float* __restrict__ mel, float* __restrict__ out, int T, int n_mels, int out_rows,
                        const PcenParams pp) {
    const int64_t clip = blockIdx.x;
    const int c = threadIdx.x;
    if (c >= n_mels) return;
    const float* x = mel + clip * (int64_t)T * n_mels + c;
    float* o = out + clip * (int64_t)out_rows * n_mels + c;
    hear_pcen_column(pp, T, out_rows, [&](int t) { return __ldg(x + (int64_t)t * n_mels); },
                     [&](int i, float v) { o[(int64_t)i * n_mels] = v; });
}

}  // namespace hmfe

using namespace hmfe;

struct hmfe_hear_plan {
    int n_mels = 0;
    HearMeta meta{};
    PcenParams pp{};
    std::vector<float> mel_dense;  // [n_mels][201]
    float *d_win = nullptr, *d_melw = nullptr;
    float2* d_plane = nullptr;
    int *d_start = nullptr, *d_row = nullptr;
    size_t smem = 0;
    int sm_count = 148;
    int last_launches = 0;
};

template <typename T>
static int hear_upload(const std::vector<T>& v, T** dptr) {
    HMFE_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), std::max<size_t>(1, v.size()) * sizeof(T)));
    HMFE_CHECK_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return HMFE_OK;
}

extern "C" {

void hmfe_hear_plan_destroy(hmfe_hear_plan* p) {
    if (!p) return;
    cudaFree(p->d_win);
    cudaFree(p->d_melw);
    cudaFree(p->d_plane);
    cudaFree(p->d_start);
    cudaFree(p->d_row);
    delete p;
}

int hmfe_hear_plan_create(hmfe_hear_plan** plan, const float* h_window, const float* h_mel, int n_mels, double alpha,
                          double smooth_coef, double delta, double root, double floor) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    HMFE_REQUIRE(h_window && h_mel, "NULL argument");
    HMFE_REQUIRE(n_mels >= 32 && n_mels % 32 == 0 && n_mels <= 32 * kHearMaxSlots, "n_mels=%d must be a multiple of 32 <= %d",
                 n_mels, 32 * kHearMaxSlots);
    HMFE_REQUIRE(root > 0 && smooth_coef >= 0 && smooth_coef <= 1, "bad PCEN parameters");
    hmfe_hear_plan* p = new (std::nothrow) hmfe_hear_plan();
    HMFE_REQUIRE(p != nullptr, "out of host memory");
    p->n_mels = n_mels;
    p->sm_count = device_sm_count();
    // audio_utils.py:208-216: alpha is clipped to <= 1, root to >= 1; :151-152: the two EMA gains
    p->pp.alpha = (float)std::min(alpha, 1.0);
    p->pp.c_in = (float)smooth_coef;
    p->pp.c_state = (float)(1.0 - smooth_coef);
    p->pp.delta = (float)delta;
    p->pp.inv_root = 1.0f / (float)std::max(root, 1.0);
    p->pp.floor = (float)floor;
    p->pp.delta_root = powf(p->pp.delta, p->pp.inv_root);
    p->mel_dense.assign((size_t)n_mels * kHearBins, 0.0f);
    for (int k = 0; k < kHearBins; ++k)
        for (int m = 0; m < n_mels; ++m) p->mel_dense[(size_t)m * kHearBins + k] = h_mel[(size_t)k * n_mels + m];
    const BandedMel bm = build_banded(p->mel_dense, n_mels, kHearBins, 8, kHearPRows);
    if (!verify_banded(bm, p->mel_dense, kHearPRows)) {
        set_error("internal error: banded mel tables do not reproduce the mel matrix");
        delete p;
        return HMFE_ERR_INVALID;
    }
    p->meta.n_slots = bm.n_slots;
    p->meta.total_trip = bm.total_trip;
    p->meta.n_mels = n_mels;
    for (int s = 0; s < kHearMaxSlots; ++s) {
        p->meta.trip[s] = s < bm.n_slots ? bm.trip[s] : 0;
        p->meta.wbase[s] = s < bm.n_slots ? bm.wbase[s] : 0;
    }
    std::vector<float> win(kHearN);
    for (int n = 0; n < kHearN; ++n) win[n] = 0.5f * h_window[n];  // the separation omits the 1/2 of (Z + conj Z') / 2
    std::vector<float2> plane(25 * 16);
    for (int k1 = 0; k1 < 25; ++k1)
        for (int n2 = 0; n2 < 16; ++n2) {
            const double a = 2.0 * kPi * (double)((n2 * k1) % kHearN) / (double)kHearN;
            plane[k1 * 16 + n2] = make_float2((float)cos(a), (float)-sin(a));
        }
    int rc = hear_upload(win, &p->d_win);
    if (rc == HMFE_OK) rc = hear_upload(plane, &p->d_plane);
    if (rc == HMFE_OK) rc = hear_upload(bm.w, &p->d_melw);
    if (rc == HMFE_OK) rc = hear_upload(bm.start, &p->d_start);
    if (rc == HMFE_OK) rc = hear_upload(bm.row, &p->d_row);
    if (rc != HMFE_OK) {
        hmfe_hear_plan_destroy(p);
        return rc;
    }
    size_t tbytes = (size_t)(400 * 8 + 400 * 4 + bm.total_trip * 128 + bm.n_slots * 256);
    tbytes = (tbytes + 15) & ~(size_t)15;
    p->smem = tbytes + kHearWarps * kHearWarpBytes;
    *plan = p;
    return HMFE_OK;
}

int hmfe_hear_num_frames(int n_padded) { return n_padded > 0 ? (n_padded + kHearShift - 1) / kHearShift : 0; }

size_t hmfe_hear_workspace_bytes(const hmfe_hear_plan* p, int64_t n_clips, int n_padded) {
    if (!p || n_clips < 0 || n_padded <= 0) return 0;
    return (size_t)n_clips * (size_t)hmfe_hear_num_frames(n_padded) * (size_t)p->n_mels * sizeof(float) + 16;
}

int hmfe_hear_last_launches(const hmfe_hear_plan* p) { return p ? p->last_launches : 0; }

static int hear_mel_stage(hmfe_hear_plan* p, const float* d_audio, int64_t n_clips, int n_samples, int n_padded, float* d_mel,
                          unsigned* d_mm, cudaStream_t st) {
    const int T = hmfe_hear_num_frames(n_padded);
    hear_minmax_init_kernel<<<1, 1, 0, st>>>(d_mm, n_samples < n_padded ? 1 : 0);
    HMFE_CHECK_CUDA(cudaGetLastError());
    const int64_t total = n_clips * (int64_t)n_samples;
    if (total > 0) {
        const int grid = (int)std::min<int64_t>((total + 1023) / 1024, (int64_t)p->sm_count * 8);
        hear_minmax_kernel<<<grid, 256, 0, st>>>(d_audio, total, d_mm);
        HMFE_CHECK_CUDA(cudaGetLastError());
    }
    HearBatch b{};
    b.audio = d_audio;
    b.mel = d_mel;
    b.mm = d_mm;
    b.n_clips = n_clips;
    b.n_samples = n_samples;
    b.n_padded = n_padded;
    b.T = T;
    b.items_per_clip = (T + 3) / 4;
    b.n_items = n_clips * (int64_t)b.items_per_clip;
    HearTables tb{p->d_win, p->d_plane, p->d_melw, p->d_start, p->d_row};
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(hear_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    const int64_t want = (b.n_items + kHearWarps - 1) / kHearWarps;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count * 2));  // 2 CTAs per SM fit
    hear_mel_kernel<<<grid, kHearWarps * 32, p->smem, st>>>(b, tb, p->meta);
    HMFE_CHECK_CUDA(cudaGetLastError());
    p->last_launches = total > 0 ? 3 : 2;
    return HMFE_OK;
}

static int hear_check(hmfe_hear_plan* p, const void* d_audio, int64_t n_clips, int n_samples, int n_padded, const void* d_out,
                      const void* d_workspace, size_t workspace_bytes, bool mel_only) {
    HMFE_REQUIRE(p, "NULL plan");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    HMFE_REQUIRE(n_samples >= 0 && n_padded >= 1 && n_samples <= n_padded && n_padded < (1 << 24),
                 "bad sample counts %d / %d (the reference rejects clips longer than the padded length, audio_utils.py:469-472)",
                 n_samples, n_padded);
    p->last_launches = 0;
    if (n_clips == 0) return 1;
    HMFE_REQUIRE((d_audio || n_samples == 0) && d_out && d_workspace, "NULL device pointer");
    const size_t need = mel_only ? 16 : hmfe_hear_workspace_bytes(p, n_clips, n_padded);
    HMFE_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
    HMFE_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0, "workspace must be 16-byte aligned");
    return HMFE_OK;
}

int hmfe_hear_mel_batch(hmfe_hear_plan* p, const float* d_audio, int64_t n_clips, int n_samples, int n_padded, float* d_mel,
                        void* d_workspace, size_t workspace_bytes, void* stream) {
    const int rc = hear_check(p, d_audio, n_clips, n_samples, n_padded, d_mel, d_workspace, workspace_bytes, true);
    if (rc != HMFE_OK) return rc == 1 ? HMFE_OK : rc;
    return hear_mel_stage(p, d_audio, n_clips, n_samples, n_padded, d_mel, static_cast<unsigned*>(d_workspace),
                          static_cast<cudaStream_t>(stream));
}

int hmfe_hear_mel_pcen_batch(hmfe_hear_plan* p, const float* d_audio, int64_t n_clips, int n_samples, int n_padded, int out_rows,
                             float* d_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    const int rc0 = hear_check(p, d_audio, n_clips, n_samples, n_padded, d_out, d_workspace, workspace_bytes, false);
    if (rc0 != HMFE_OK) return rc0 == 1 ? HMFE_OK : rc0;
    HMFE_REQUIRE(out_rows >= 1, "out_rows < 1");
    HMFE_REQUIRE(n_clips < (int64_t)INT32_MAX, "too many clips for one call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* d_mm = static_cast<unsigned*>(d_workspace);
    float* d_mel = reinterpret_cast<float*>(static_cast<unsigned char*>(d_workspace) + 16);
    const int rc = hear_mel_stage(p, d_audio, n_clips, n_samples, n_padded, d_mel, d_mm, st);
    if (rc != HMFE_OK) return rc;
    hear_pcen_resize_kernel<<<(unsigned)n_clips, 128, 0, st>>>(d_mel, d_out, hmfe_hear_num_frames(n_padded), p->n_mels, out_rows,
                                                               p->pp);
    HMFE_CHECK_CUDA(cudaGetLastError());
    p->last_launches += 1;
    return HMFE_OK;
}

}  // extern "C"
