// Zero-phase band-pass: scipy.signal.sosfiltfilt(sos, x) (padtype="odd", padlen = 3 * ntaps) on ragged batches.
// BASELINE.json's north_star names sosfiltfilt; the reference itself calls the causal lfilter
// (/root/reference/src/util.py:113-126), which hmfe_iir_sos_batch reproduces.  This entry is the
// additional zero-phase mode, built from the same IIR kernels:
//   ext   = odd extension of every clip by `edge` samples on both sides            (float32, like numpy)
//   y1    = sosfilt(ext, zi * ext[0])      = zero-state response + ext[0] * h_zi   (float64)
//   y2    = sosfilt(reverse(y1), zi * y1[-1])                                      (float64)
//   y     = reverse(y2)[edge : -edge]
// where zi is scipy's sosfilt_zi (step-response steady state) and h_zi[n] the cascade's zero-input response from
// that state, tabulated on the host until it has decayed.  The reversed intermediate is fed back as float32
// (the IIR kernels take float32 input): relative error <= 6e-8, far inside the 1e-4 budget.
#include <math.h>

#include <algorithm>
#include <vector>

#include "api_common.h"
#include "ctx.h"
#include "hmfe_common.cuh"

namespace hmfe {

struct FiltfiltBatch {
    const int64_t* clip_off;  // [n_clips+1] offsets of the clips in x / y
    const int64_t* ext_off;   // [n_clips+1] offsets of the extended clips
    int64_t n_clips;
    int edge;
};

HMFE_D int64_t find_clip(const int64_t* off, int64_t n_clips, int64_t i) {
    int64_t lo = 0, hi = n_clips;  // largest c with off[c] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= i)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// scipy.signal._arraytools.odd_ext, in the input's float32 like numpy
__global__ void __launch_bounds__(256) odd_ext_kernel(const float* __restrict__ x, float* __restrict__ ext, const FiltfiltBatch b,
                                                      int64_t total_ext) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_ext; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = find_clip(b.ext_off, b.n_clips, i);
        const int64_t k = i - b.ext_off[c];
        const float* xc = x + b.clip_off[c];
        const int64_t n = b.clip_off[c + 1] - b.clip_off[c];
        float v;
        if (k < b.edge)
            v = 2.0f * xc[0] - xc[b.edge - k];
        else if (k < b.edge + n)
            v = xc[k - b.edge];
        else
            v = 2.0f * xc[n - 1] - xc[n - 2 - (k - b.edge - n)];
        ext[i] = v;
    }
}

// y[n] += first * h[n] for the first `len` samples of every extended clip (first = ext[0] of the pass input)
__global__ void __launch_bounds__(256) add_zi_response_kernel(double* __restrict__ y, const float* __restrict__ pass_in,
                                                              const double* __restrict__ h, int len, const FiltfiltBatch b) {
    for (int64_t c = blockIdx.x; c < b.n_clips; c += gridDim.x) {  // one CTA per clip
        const int64_t e0 = b.ext_off[c], ne = b.ext_off[c + 1] - e0;
        const double first = (double)pass_in[e0];
        const int64_t m = min((int64_t)len, ne);
        for (int64_t n = threadIdx.x; n < m; n += blockDim.x) y[e0 + n] = fma(first, h[n], y[e0 + n]);
    }
}

// out32[k] = (float) y[ne - 1 - k] per extended clip
__global__ void __launch_bounds__(256) reverse_cast_kernel(const double* __restrict__ y, float* __restrict__ out, const FiltfiltBatch b,
                                                           int64_t total_ext) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_ext; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = find_clip(b.ext_off, b.n_clips, i);
        const int64_t e0 = b.ext_off[c], ne = b.ext_off[c + 1] - e0;
        out[i] = (float)y[e0 + (ne - 1 - (i - e0))];
    }
}

// y[j] = y2[ne - 1 - (edge + j)] for j < n
__global__ void __launch_bounds__(256) reverse_crop_kernel(const double* __restrict__ y2, float* __restrict__ y32, double* __restrict__ y64,
                                                           const FiltfiltBatch b, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = find_clip(b.clip_off, b.n_clips, i);
        const int64_t j = i - b.clip_off[c];
        const int64_t e0 = b.ext_off[c], ne = b.ext_off[c + 1] - e0;
        const double v = y2[e0 + (ne - 1 - (b.edge + j))];
        if (y32) y32[i] = (float)v;
        if (y64) y64[i] = v;
    }
}

}  // namespace hmfe

using namespace hmfe;

static int default_padlen(const double* sos, int S) {  // scipy: 3 * (2*S + 1 - min(#(b2 == 0), #(a2 == 0)))
    int zb = 0, za = 0;
    for (int k = 0; k < S; ++k) {
        zb += sos[6 * k + 2] == 0.0;
        za += sos[6 * k + 5] == 0.0;
    }
    return 3 * (2 * S + 1 - std::min(zb, za));
}

extern "C" int hmfe_sosfiltfilt_padlen(const double* h_sos, int n_sections) {
    if (!h_sos || n_sections < 1) return -1;
    return default_padlen(h_sos, n_sections);
}

extern "C" int64_t hmfe_sosfiltfilt_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int padlen) {
    if (!h_offsets || n_clips < 0 || padlen < 0) return -1;
    const int64_t total_ext = (h_offsets[n_clips] - h_offsets[0]) + 2 * (int64_t)padlen * n_clips;
    // float32 pass input + float64 pass output, each 256-byte aligned
    return ((total_ext * 4 + 255) & ~(int64_t)255) + ((total_ext * 8 + 255) & ~(int64_t)255) + 256;
}

extern "C" int hmfe_sosfiltfilt_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                                      const double* h_sos, int n_sections, int padlen, void* d_workspace,
                                      int64_t workspace_bytes, float* d_y32, double* d_y64, void* stream) {
    HMFE_REQUIRE(ctx && h_offsets && h_sos, "NULL argument");
    HMFE_REQUIRE(n_sections >= 1 && n_sections <= 8, "n_sections=%d not in [1, 8]", n_sections);
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_x && (d_y32 || d_y64) && d_workspace, "NULL device pointer");
    const int S = n_sections;
    const int edge = padlen >= 0 ? padlen : default_padlen(h_sos, S);
    for (int64_t i = 0; i < n_clips; ++i)
        HMFE_REQUIRE(h_offsets[i + 1] - h_offsets[i] > edge,
                     "clip %lld: the length of the input vector x must be greater than padlen, which is %d.", (long long)i, edge);
    HMFE_REQUIRE(workspace_bytes >= hmfe_sosfiltfilt_workspace_bytes(h_offsets, n_clips, edge), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // scipy.signal.sosfilt_zi and the zero-input response from that state (normalised sections)
    std::vector<double> sos((size_t)6 * S), zi((size_t)2 * S);
    double scale = 1.0;
    for (int k = 0; k < S; ++k) {
        const double a0 = h_sos[6 * k + 3];
        HMFE_REQUIRE(a0 != 0.0, "section %d has a0 == 0", k);
        for (int c = 0; c < 6; ++c) sos[6 * k + c] = h_sos[6 * k + c] / a0;
        const double* q = &sos[6 * k];
        const double sa = 1.0 + q[4] + q[5];
        HMFE_REQUIRE(sa != 0.0, "section %d has a pole at z = 1: no step-response steady state", k);
        const double g = (q[0] + q[1] + q[2]) / sa;
        const double z2 = q[2] - q[5] * g, z1 = q[1] - q[4] * g + z2;
        zi[2 * k] = scale * z1;
        zi[2 * k + 1] = scale * z2;
        scale *= g;
    }
    std::vector<double> h;
    {
        std::vector<double> s = zi;
        double peak = 0.0;
        for (int n = 0; n < 1 << 20; ++n) {
            double v = 0.0;
            for (int k = 0; k < S; ++k) {
                const double* q = &sos[6 * k];
                const double y = fma(q[0], v, s[2 * k]);
                s[2 * k] = fma(q[1], v, fma(-q[4], y, s[2 * k + 1]));
                s[2 * k + 1] = fma(q[2], v, -q[5] * y);
                v = y;
            }
            h.push_back(v);
            peak = std::max(peak, fabs(v));
            double smax = 0.0;
            for (double e : s) smax = std::max(smax, fabs(e));
            if (n >= 32 && smax <= 1e-17 * std::max(peak, 1e-300)) break;
        }
        HMFE_REQUIRE(h.size() < (size_t)1 << 20, "the filter's transient does not decay: zero-phase mode unavailable");
    }

    const size_t idx_bytes = 2 * (size_t)(n_clips + 1) * sizeof(int64_t);
    const size_t h_bytes = h.size() * sizeof(double);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(idx_bytes + h_bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    int64_t* hc = static_cast<int64_t*>(hbuf);
    int64_t* he = hc + (n_clips + 1);
    std::vector<int64_t> ext_off((size_t)n_clips + 1);
    he[0] = 0;
    for (int64_t i = 0; i <= n_clips; ++i) hc[i] = h_offsets[i] - h_offsets[0];
    for (int64_t i = 0; i < n_clips; ++i) he[i + 1] = he[i] + (h_offsets[i + 1] - h_offsets[i]) + 2 * (int64_t)edge;
    for (int64_t i = 0; i <= n_clips; ++i) ext_off[i] = he[i];
    memcpy(static_cast<unsigned char*>(hbuf) + idx_bytes, h.data(), h_bytes);
    int rc = ctx->ring.upload(slot, idx_bytes + h_bytes, st);
    if (rc != HMFE_OK) return rc;
    FiltfiltBatch b{static_cast<int64_t*>(dbuf), static_cast<int64_t*>(dbuf) + (n_clips + 1), n_clips, edge};
    const double* d_h = reinterpret_cast<const double*>(static_cast<unsigned char*>(dbuf) + idx_bytes);
    const int64_t total = hc[n_clips], total_ext = he[n_clips];
    unsigned char* ws = static_cast<unsigned char*>(d_workspace);
    ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
    float* pass_in = reinterpret_cast<float*>(ws);
    double* pass_out = reinterpret_cast<double*>(ws + ((total_ext * 4 + 255) & ~(int64_t)255));
    const float* x0 = d_x + h_offsets[0];
    const int grid = (int)std::min<int64_t>((total_ext + 255) / 256, (int64_t)ctx->sm_count * 16);
    const int zi_grid = (int)std::min<int64_t>(n_clips, (int64_t)ctx->sm_count * 16);

    odd_ext_kernel<<<grid, 256, 0, st>>>(x0, pass_in, b, total_ext);
    HMFE_CHECK_CUDA(cudaGetLastError());
    int launches = 1;
    for (int pass = 0; pass < 2; ++pass) {
        rc = hmfe_iir_sos_batch(ctx, pass_in, ext_off.data(), n_clips, h_sos, S, nullptr, pass_out, stream);
        if (rc != HMFE_OK) return rc;
        launches += ctx->last_launches;
        add_zi_response_kernel<<<zi_grid, 256, 0, st>>>(pass_out, pass_in, d_h, (int)h.size(), b);
        HMFE_CHECK_CUDA(cudaGetLastError());
        ++launches;
        if (pass == 0) {
            reverse_cast_kernel<<<grid, 256, 0, st>>>(pass_out, pass_in, b, total_ext);
            HMFE_CHECK_CUDA(cudaGetLastError());
            ++launches;
        }
    }
    reverse_crop_kernel<<<(int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(
        pass_out, d_y32 ? d_y32 + h_offsets[0] : nullptr, d_y64 ? d_y64 + h_offsets[0] : nullptr, b, total);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->last_launches = launches + 1;
    return ctx->ring.release(slot, st);
}
