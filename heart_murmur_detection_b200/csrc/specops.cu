// Spectrogram-domain dataset ops on the GPU: crop / Markov frame mask / gain / zero pad.
// Mirrors crop_first, random_crop, random_mask, random_multiply (/root/reference/src/util.py:26-51)
// and the pad-or-crop to a fixed number of frames in the Dataset classes
// (src/pretrain/cola_training.py:56-80, src/pretrain/mae_training.py:88-109,
//  src/benchmark/baseline/audioMAE/models_mae.py:1178-1181).
// All random draws are made on the host with Python's `random` in the reference's order
// (frontend.py), so crop starts and masked frames are bit-exact; this file only applies them.
#include <algorithm>

#include "api_common.h"
#include "ctx.h"

namespace hmfe {

// mean over all elements of each ragged spectrogram (float64 accumulation, float32 result)
// rows [row_lo[s], row_hi[s]) of spectrogram s (row_hi = row_lo + 1 for a packed ragged batch)
__global__ void __launch_bounds__(256) spec_mean_kernel(const float* __restrict__ spec, const int64_t* __restrict__ row_lo,
                                                        const int64_t* __restrict__ row_hi, int n_cols,
                                                        float* __restrict__ mean) {
    __shared__ double s_acc[8];
    const int64_t s = blockIdx.x;
    const int64_t e0 = row_lo[s] * n_cols, e1 = row_hi[s] * n_cols;
    double acc = 0.0;
    for (int64_t i = e0 + threadIdx.x; i < e1; i += 256) acc += (double)spec[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_acc[w];
        mean[s] = e1 > e0 ? (float)(t / (double)(e1 - e0)) : 0.0f;
    }
}

// one CTA per (output item, tile of rows)
constexpr int kCropRowsPerCta = 16;

__global__ void __launch_bounds__(256)
spec_crop_kernel(const float* __restrict__ spec, float* __restrict__ out, const hmfe_crop_desc* __restrict__ descs,
                 const uint8_t* __restrict__ mask, const float* __restrict__ mean, int n_cols, int out_rows, int tiles) {
    const int64_t item = blockIdx.x / tiles;
    const int r0 = (int)(blockIdx.x % tiles) * kCropRowsPerCta;
    const hmfe_crop_desc d = descs[item];
    const float fill = (mean != nullptr) ? mean[d.spec_id] : 0.0f;
    float* o = out + item * (int64_t)out_rows * n_cols;
    const int n = min(kCropRowsPerCta, out_rows - r0) * n_cols;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int r = r0 + i / n_cols, c = i - (i / n_cols) * n_cols;
        float v = 0.0f;
        if (r < d.n_rows) {
            const int64_t src_row = d.src_row + r;
            const bool masked = mask != nullptr && mask[(int64_t)d.mask_off + r] != 0;
            v = masked ? fill : __ldg(spec + src_row * n_cols + c);
            v *= d.gain;
        }
        o[(int64_t)r * n_cols + c] = v;
    }
}

// SpecAugmentation stripes (finetuning.py:104-116): zero a rectangle of an output item
__global__ void __launch_bounds__(256)
spec_zero_rects_kernel(float* __restrict__ out, const hmfe_rect_desc* __restrict__ rects, int out_rows, int n_cols) {
    const hmfe_rect_desc r = rects[blockIdx.x];
    float* o = out + r.item * (int64_t)out_rows * n_cols;
    const int n = r.n_rows * r.n_cols;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int rr = i / r.n_cols, cc = i - rr * r.n_cols;
        o[(int64_t)(r.row0 + rr) * n_cols + r.col0 + cc] = 0.0f;
    }
}

}  // namespace hmfe

using namespace hmfe;

extern "C" {

int hmfe_spec_mean_batch(hmfe_ctx* ctx, const float* d_spec, const int64_t* h_row_offsets, int64_t n_specs, int n_cols,
                         float* d_mean, void* stream) {
    HMFE_REQUIRE(ctx && h_row_offsets, "NULL argument");
    HMFE_REQUIRE(n_specs >= 0 && n_cols > 0, "bad arguments");
    ctx->last_launches = 0;
    if (n_specs == 0) return HMFE_OK;
    HMFE_REQUIRE(d_spec && d_mean, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)(n_specs + 1) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_row_offsets, bytes);
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    ctx->prof_begin(HMFE_K_SPEC_MEAN, st);
    spec_mean_kernel<<<(unsigned)n_specs, 256, 0, st>>>(d_spec, static_cast<int64_t*>(dbuf), static_cast<int64_t*>(dbuf) + 1,
                                                         n_cols, d_mean);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

int hmfe_spec_mean_ranges(hmfe_ctx* ctx, const float* d_spec, const int64_t* h_row_lo, const int64_t* h_row_hi, int64_t n,
                          int n_cols, float* d_mean, void* stream) {
    HMFE_REQUIRE(ctx && h_row_lo && h_row_hi, "NULL argument");
    HMFE_REQUIRE(n >= 0 && n_cols > 0, "bad arguments");
    ctx->last_launches = 0;
    if (n == 0) return HMFE_OK;
    HMFE_REQUIRE(d_spec && d_mean, "NULL device pointer");
    for (int64_t i = 0; i < n; ++i)
        HMFE_REQUIRE(h_row_lo[i] >= 0 && h_row_hi[i] >= h_row_lo[i], "row range %lld is inverted", (long long)i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)(2 * n) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_row_lo, (size_t)n * sizeof(int64_t));
    memcpy(static_cast<int64_t*>(hbuf) + n, h_row_hi, (size_t)n * sizeof(int64_t));
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    ctx->prof_begin(HMFE_K_SPEC_MEAN, st);
    spec_mean_kernel<<<(unsigned)n, 256, 0, st>>>(d_spec, static_cast<int64_t*>(dbuf), static_cast<int64_t*>(dbuf) + n, n_cols,
                                                  d_mean);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

int hmfe_spec_crop_batch(hmfe_ctx* ctx, const float* d_spec, int n_cols, const hmfe_crop_desc* h_descs, int64_t n_items,
                         const uint8_t* d_row_mask, const float* d_mean, float* d_out, int out_rows, void* stream) {
    HMFE_REQUIRE(ctx && (h_descs || n_items == 0), "NULL argument");
    HMFE_REQUIRE(n_items >= 0 && n_cols > 0 && out_rows > 0, "bad arguments");
    HMFE_REQUIRE(d_row_mask == nullptr || d_mean != nullptr, "a row mask needs the per-spectrogram means");
    ctx->last_launches = 0;
    if (n_items == 0) return HMFE_OK;
    HMFE_REQUIRE(d_spec && d_out, "NULL device pointer");
    for (int64_t i = 0; i < n_items; ++i)
        HMFE_REQUIRE(h_descs[i].src_row >= 0 && h_descs[i].n_rows >= 0 && h_descs[i].n_rows <= out_rows &&
                         h_descs[i].spec_id >= 0 && h_descs[i].mask_off >= 0,
                     "crop descriptor %lld is inconsistent", (long long)i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)n_items * sizeof(hmfe_crop_desc);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_descs, bytes);
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    const int tiles = (out_rows + kCropRowsPerCta - 1) / kCropRowsPerCta;
    HMFE_REQUIRE(n_items * tiles < (int64_t)INT32_MAX, "crop grid too large");
    ctx->prof_begin(HMFE_K_SPEC_CROP, st);
    spec_crop_kernel<<<(unsigned)(n_items * tiles), 256, 0, st>>>(d_spec, d_out, static_cast<hmfe_crop_desc*>(dbuf),
                                                                    d_row_mask, d_mean, n_cols, out_rows, tiles);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

int hmfe_spec_zero_rects(hmfe_ctx* ctx, float* d_out, int out_rows, int n_cols, int64_t n_items, const hmfe_rect_desc* h_rects,
                         int64_t n_rects, void* stream) {
    HMFE_REQUIRE(ctx && (h_rects || n_rects == 0), "NULL argument");
    HMFE_REQUIRE(n_rects >= 0 && n_rects < (int64_t)INT32_MAX && out_rows > 0 && n_cols > 0, "bad arguments");
    ctx->last_launches = 0;
    if (n_rects == 0) return HMFE_OK;
    HMFE_REQUIRE(d_out, "NULL device pointer");
    for (int64_t i = 0; i < n_rects; ++i) {
        const hmfe_rect_desc& r = h_rects[i];
        HMFE_REQUIRE(r.item >= 0 && r.item < n_items && r.row0 >= 0 && r.n_rows >= 0 && r.row0 + r.n_rows <= out_rows &&
                         r.col0 >= 0 && r.n_cols >= 0 && r.col0 + r.n_cols <= n_cols,
                     "rectangle %lld lies outside its item", (long long)i);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)n_rects * sizeof(hmfe_rect_desc);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_rects, bytes);
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    spec_zero_rects_kernel<<<(unsigned)n_rects, 256, 0, st>>>(d_out, static_cast<hmfe_rect_desc*>(dbuf), out_rows, n_cols);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

}  // extern "C"
