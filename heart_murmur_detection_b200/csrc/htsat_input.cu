// HTS-AT input stage: what the OPERA-CT encoder does to the path's log-mel before its first
// layer (/root/reference/src/model/htsat/htsat.py):
//   :889-891  bn0 - BatchNorm2d over the mel bins, inference form  y = x * a_f + b_f
//   :829-858  reshape_wav2img - bicubic (align_corners=True, A = -0.75, clamped taps) resize of the time axis to
//             spec_size * freq_ratio frames, then the fold  out[n*F + f][t'] = y[n * (T'/ratio) + t'][f]
// fused into one kernel that reads the ragged [T_i, F] spectrograms once and writes the [spec, spec] images.
#include <math.h>

#include "api_common.h"
#include "ctx.h"

namespace hmfe {

struct HtsatBatch {
    const float* spec;       // [rows, F]
    float* out;              // [n_items, spec_size, spec_size]
    const int64_t* src_row;  // [n_items]
    const int* n_rows;       // [n_items]
    const float* scale;      // [F]  a_f
    const float* shift;      // [F]  b_f
    int F, ratio, target_T, spec_size;
};

// F == 64 (the reference's mel_bins): 256 threads = 64 bins x 4 time phases, 32 output frames per CTA
__global__ void __launch_bounds__(256) htsat_input_kernel(const HtsatBatch b) {
    __shared__ float tile[64][33];
    const int tpb = b.target_T / 32;  // 32-frame blocks per item
    const int64_t item = blockIdx.x / tpb;
    const int t_block = (int)(blockIdx.x % tpb) * 32;
    const int T = b.n_rows[item];
    const float* x = b.spec + b.src_row[item] * b.F;
    const int f = threadIdx.x & 63, tq = threadIdx.x >> 6;
    const float a = b.scale[f], c = b.shift[f];
    const float A = -0.75f;
    const float step = b.target_T > 1 ? (float)(T - 1) / (float)(b.target_T - 1) : 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int tl = tq + 4 * j, t_out = t_block + tl;
        float v;
        if (T == b.target_T) {  // the reference does not interpolate a full-length input
            v = fmaf(x[(int64_t)t_out * b.F + f], a, c);
        } else {
            const float real = step * (float)t_out;
            const float fl = floorf(real);
            const int i0 = (int)fl;
            const float t = real - fl;
            const float u = 1.0f - t;
            const float c0 = ((A * (t + 1.0f) - 5.0f * A) * (t + 1.0f) + 8.0f * A) * (t + 1.0f) - 4.0f * A;
            const float c1 = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
            const float c2 = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
            const float c3 = ((A * (u + 1.0f) - 5.0f * A) * (u + 1.0f) + 8.0f * A) * (u + 1.0f) - 4.0f * A;
            const int r0 = min(max(i0 - 1, 0), T - 1), r1 = min(max(i0, 0), T - 1);
            const int r2 = min(max(i0 + 1, 0), T - 1), r3 = min(max(i0 + 2, 0), T - 1);
            const float y0 = fmaf(__ldg(x + (int64_t)r0 * b.F + f), a, c), y1 = fmaf(__ldg(x + (int64_t)r1 * b.F + f), a, c);
            const float y2 = fmaf(__ldg(x + (int64_t)r2 * b.F + f), a, c), y3 = fmaf(__ldg(x + (int64_t)r3 * b.F + f), a, c);
            v = y0 * c0 + y1 * c1 + y2 * c2 + y3 * c3;
        }
        tile[f][tl] = v;
    }
    __syncthreads();
    // fold: frame t_out = n * per + t'  ->  row n*F + f, column t'
    const int per = b.target_T / b.ratio;
    const int n = t_block / per, tcol0 = t_block - n * per;
    float* o = b.out + item * (int64_t)b.spec_size * b.spec_size;
    const int tl = threadIdx.x & 31, fr = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ff = fr + 8 * j;
        o[(int64_t)(n * b.F + ff) * b.spec_size + tcol0 + tl] = tile[ff][tl];
    }
}

}  // namespace hmfe

using namespace hmfe;

extern "C" int hmfe_htsat_input_batch(hmfe_ctx* ctx, const float* d_spec, int n_cols, const int64_t* h_src_row,
                                      const int32_t* h_n_rows, int64_t n_items, const float* h_scale, const float* h_shift,
                                      int spec_size, float* d_out, void* stream) {
    HMFE_REQUIRE(ctx && (n_items == 0 || (h_src_row && h_n_rows)) && h_scale && h_shift, "NULL argument");
    HMFE_REQUIRE(n_items >= 0, "n_items < 0");
    if (n_cols != 64 || spec_size <= 0 || spec_size % n_cols != 0) {
        set_error("htsat input stage is specialised for 64 mel bins and spec_size a multiple of 64 (htsat.py:545-573)");
        return HMFE_ERR_UNSUPPORTED;
    }
    ctx->last_launches = 0;
    if (n_items == 0) return HMFE_OK;
    HMFE_REQUIRE(d_spec && d_out, "NULL device pointer");
    const int ratio = spec_size / n_cols, target_T = spec_size * ratio;
    HMFE_REQUIRE(target_T % (32 * ratio) == 0, "spec_size=%d unsupported", spec_size);
    for (int64_t i = 0; i < n_items; ++i)
        HMFE_REQUIRE(h_src_row[i] >= 0 && h_n_rows[i] >= 1 && h_n_rows[i] <= target_T,
                     "item %lld: %d frames do not fit the %d-frame image (htsat.py:833-835 asserts the same)", (long long)i,
                     h_n_rows[i], target_T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t row_bytes = (size_t)n_items * sizeof(int64_t), cnt_bytes = ((size_t)n_items * sizeof(int32_t) + 15) & ~(size_t)15;
    const size_t bytes = row_bytes + cnt_bytes + 2 * (size_t)n_cols * sizeof(float);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    unsigned char* hb = static_cast<unsigned char*>(hbuf);
    memcpy(hb, h_src_row, row_bytes);
    memcpy(hb + row_bytes, h_n_rows, (size_t)n_items * sizeof(int32_t));
    memcpy(hb + row_bytes + cnt_bytes, h_scale, (size_t)n_cols * sizeof(float));
    memcpy(hb + row_bytes + cnt_bytes + (size_t)n_cols * sizeof(float), h_shift, (size_t)n_cols * sizeof(float));
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    unsigned char* db = static_cast<unsigned char*>(dbuf);
    HtsatBatch b{};
    b.spec = d_spec;
    b.out = d_out;
    b.src_row = reinterpret_cast<const int64_t*>(db);
    b.n_rows = reinterpret_cast<const int*>(db + row_bytes);
    b.scale = reinterpret_cast<const float*>(db + row_bytes + cnt_bytes);
    b.shift = b.scale + n_cols;
    b.F = n_cols;
    b.ratio = ratio;
    b.target_T = target_T;
    b.spec_size = spec_size;
    const int64_t grid = n_items * (target_T / 32);
    HMFE_REQUIRE(grid < (int64_t)INT32_MAX, "grid too large");
    htsat_input_kernel<<<(unsigned)grid, 256, 0, st>>>(b);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}
