// Batched polyphase FIR resampler (to 16 kHz).
//
// The reference resamples inside librosa.load(path, sr=16000) (src/util.py:153,222,323,391,805;
// extract_feature.py:214) with libsoxr's HQ filter - a third-party C library that is not
// vendored and not installable here (parity unpinned) - and with
// torchaudio.transforms.Resample at src/model/models_eval.py:964-968.  This kernel implements
// the latter's published algorithm (Hann-windowed sinc, lowpass_filter_width 6, rolloff 0.99,
// or the Kaiser variant), which is the pinned oracle:
//   out[j] = sum_k  kern[j mod U][k] * xpad[(j div U) * D + k],   k in [0, 2*width + D)
//   U = new/gcd, D = orig/gcd, xpad = x left-padded by `width` zeros, out length ceil(U*n/D).
#include <math.h>

#include <algorithm>
#include <new>
#include <vector>

#include "api_common.h"
#include "tables.h"

namespace hmfe {

#define HMFE_RS_D __device__ __forceinline__

struct ResampleBatch {
    const void* x;           // float32 samples, or int16 PCM (decoded as x / 32768 on the fly)
    float* y;
    const int64_t* in_off;   // [n_clips+1]
    const int64_t* out_off;  // [n_clips+1]
    const float* kern;       // [U][W]
    int64_t n_clips, n_out_total;
    int U, D, W, width;
    int taps_in_smem;        // 0: the tap table is too large for shared memory (e.g. 44.1 kHz -> 16 kHz: 160 x 475), read it through L1
};

constexpr int kRsTile = 2048;  // outputs per CTA

HMFE_RS_D float rs_sample(const float* x, int64_t i) { return __ldg(x + i); }
HMFE_RS_D float rs_sample(const int16_t* x, int64_t i) { return (float)__ldg(x + i) * (1.0f / 32768.0f); }  // exact

// One CTA = one tile of kRsTile consecutive outputs of one clip.  The taps [U][W] and the input span of the tile
// ((kRsTile / U) * D + W samples, zero padded at the clip edges) are staged in shared memory once; consecutive
// threads compute consecutive outputs (for the integer up-factors of CirCor / PhysioNet audio, D = 1: the U
// threads of one input position read the same samples - a broadcast - and U different tap rows).
template <typename T>
__global__ void __launch_bounds__(256) resample_kernel(const ResampleBatch b, const int64_t* tile_prefix) {
    extern __shared__ float rs_smem[];
    const float* s_k = b.taps_in_smem ? rs_smem : b.kern;           // [U * W]
    float* s_x = rs_smem + (b.taps_in_smem ? b.U * b.W : 0);        // [span]
    const int64_t tile = blockIdx.x;
    int64_t lo = 0, hi = b.n_clips;
    while (hi - lo > 1) {  // tile -> clip (binary search over the per-clip tile prefix)
        const int64_t mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= tile)
            lo = mid;
        else
            hi = mid;
    }
    const int64_t clip = lo;
    const int64_t i0 = b.in_off[clip];
    const int n_in = (int)(b.in_off[clip + 1] - i0);
    const int64_t o0 = b.out_off[clip];
    const int n_out = (int)(b.out_off[clip + 1] - o0);
    const T* x = static_cast<const T*>(b.x) + i0;
    const int j0 = (int)(tile - tile_prefix[clip]) * kRsTile;
    const int j1 = min(n_out, j0 + kRsTile);
    const int q0 = j0 / b.U;
    const int span = ((j1 - 1) / b.U - q0) * b.D + b.W;  // input positions q0 * D - width ... of this tile
    const int base0 = q0 * b.D - b.width;
    if (b.taps_in_smem)
        for (int i = threadIdx.x; i < b.U * b.W; i += 256) rs_smem[i] = __ldg(b.kern + i);
    for (int i = threadIdx.x; i < span; i += 256) {
        const int g = base0 + i;
        s_x[i] = (g >= 0 && g < n_in) ? rs_sample(x, g) : 0.0f;
    }
    __syncthreads();
    for (int j = j0 + threadIdx.x; j < j1; j += 256) {
        const int p = j % b.U, q = j / b.U;
        const float* k = s_k + p * b.W;
        const float* xs = s_x + (q - q0) * b.D;
        // same order of accumulation as the per-output loop it replaces (t ascending; zero-padded samples add
        // exact zeros), so results are unchanged bit for bit
        float acc = 0.0f;
        const int base = q * b.D - b.width;
        const int k_lo = max(0, -base), k_hi = min(b.W, n_in - base);
        for (int t = k_lo; t < k_hi; ++t) acc = fmaf(k[t], xs[t], acc);
        b.y[o0 + j] = acc;
    }
}

// Integer up-factors (D = 1: 2 / 4 / 8 kHz -> 16 kHz, the rates of the heart-sound corpora).  One thread = 4 consecutive
// input positions = 4 U consecutive outputs.  The input window is read as 16-byte shared-memory vectors (conflict
// free; scalar reads at a stride of 4 floats would be 4-way bank conflicts) and slides through registers; the U
// phases of a tap are one broadcast vector read, used for all 4 positions: (1 + 4) vector reads per 16 U FMAs.
// Outputs leave as 16-byte stores.  Taps are padded with zeros to a multiple of 4 (an FMA with a zero tap leaves the
// accumulator unchanged), accumulation runs over the taps in ascending order like the generic kernel: same results.
constexpr int kUpTileQ = 1024;     // input positions per tile (256 threads x 4)
constexpr int kUpTilesPerCta = 8;  // consecutive tiles per CTA: one clip lookup, then a linear walk

template <typename T, int U>
__global__ void __launch_bounds__(256) resample_up_kernel(const ResampleBatch b, const int64_t* tile_prefix, int64_t n_tiles) {
    extern __shared__ __align__(16) float rs_smem[];
    const int W = b.W, W4 = (W + 3) & ~3;
    float* s_k = rs_smem;            // [W4][U]: the U phases of one tap are adjacent
    float* s_x = rs_smem + U * W4;   // [kUpTileQ + W4 + 4]
    for (int i = threadIdx.x; i < U * W4; i += 256) {
        const int t = i / U, p = i - t * U;
        s_k[i] = t < W ? __ldg(b.kern + p * W + t) : 0.0f;
    }
    const int64_t tile0 = (int64_t)blockIdx.x * kUpTilesPerCta;
    int64_t lo = 0, hi = b.n_clips;
    while (hi - lo > 1) {  // tile -> clip
        const int64_t mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= tile0)
            lo = mid;
        else
            hi = mid;
    }
    int64_t clip = lo;
    for (int64_t tile = tile0; tile < min(tile0 + kUpTilesPerCta, n_tiles); ++tile) {
        while (tile >= tile_prefix[clip + 1]) ++clip;
        const int64_t i0 = b.in_off[clip];
        const int n_in = (int)(b.in_off[clip + 1] - i0);
        const int64_t o0 = b.out_off[clip];
        const T* x = static_cast<const T*>(b.x) + i0;
        const int q0 = (int)(tile - tile_prefix[clip]) * kUpTileQ;
        const int span = kUpTileQ + W4 + 4, base0 = q0 - b.width;
        __syncthreads();  // the previous tile's window has been consumed (and the taps are in place)
        for (int i = threadIdx.x; i < span; i += 256) {
            const int g = base0 + i;
            s_x[i] = (g >= 0 && g < n_in) ? rs_sample(x, g) : 0.0f;
        }
        __syncthreads();
        const int ql = threadIdx.x * 4;
        if (q0 + ql >= n_in) continue;
        float acc[4][U];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int p = 0; p < U; ++p) acc[r][p] = 0.0f;
        float4 xa = *reinterpret_cast<const float4*>(s_x + ql);
        for (int t0 = 0; t0 < W4; t0 += 4) {
            const float4 xb = *reinterpret_cast<const float4*>(s_x + ql + t0 + 4);
            const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                float kv[U];
#pragma unroll
                for (int p = 0; p < U; ++p) kv[p] = s_k[(t0 + tt) * U + p];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int p = 0; p < U; ++p) acc[r][p] = fmaf(kv[p], xs[tt + r], acc[r][p]);
            }
            xa = xb;
        }
        float* y = b.y + o0 + (int64_t)(q0 + ql) * U;
        if ((reinterpret_cast<uintptr_t>(y) & 15) == 0 && q0 + ql + 4 <= n_in) {
#pragma unroll
            for (int i = 0; i < U; ++i) {  // 4 U floats = U float4
                const int e = 4 * i;
                reinterpret_cast<float4*>(y)[i] = make_float4(acc[e / U][e % U], acc[(e + 1) / U][(e + 1) % U],
                                                              acc[(e + 2) / U][(e + 2) % U], acc[(e + 3) / U][(e + 3) % U]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (q0 + ql + r >= n_in) break;
#pragma unroll
                for (int p = 0; p < U; ++p) y[r * U + p] = acc[r][p];
            }
        }
    }
}

static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        term *= q / ((double)k * k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

}  // namespace hmfe

using namespace hmfe;

struct hmfe_resample_plan {
    int orig, target, U, D, W, width;
    std::vector<float> kern;
    float* d_kern = nullptr;
    DescRing ring;
    int last_launches = 0;
};

static int gcd_int(int a, int b) {
    while (b) {
        const int t = a % b;
        a = b;
        b = t;
    }
    return a;
}

extern "C" {

void hmfe_resample_plan_destroy(hmfe_resample_plan* p) {
    if (!p) return;
    cudaFree(p->d_kern);
    delete p;
}

// method 0: sinc_interp_hann, 1: sinc_interp_kaiser (beta <= 0 -> torchaudio default 14.769656459379492)
int hmfe_resample_plan_create(hmfe_resample_plan** plan, int orig_freq, int new_freq, int lowpass_filter_width,
                              double rolloff, int method, double beta) {
    HMFE_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    HMFE_REQUIRE(orig_freq > 0 && new_freq > 0 && lowpass_filter_width > 0 && rolloff > 0 && rolloff <= 1.0,
                 "bad resample parameters");
    HMFE_REQUIRE(method == 0 || method == 1, "bad method %d", method);
    hmfe_resample_plan* p = new (std::nothrow) hmfe_resample_plan();
    HMFE_REQUIRE(p != nullptr, "out of host memory");
    const int g = gcd_int(orig_freq, new_freq);
    p->orig = orig_freq;
    p->target = new_freq;
    p->D = orig_freq / g;
    p->U = new_freq / g;
    const double base_freq = std::min(p->D, p->U) * rolloff;
    p->width = (int)ceil(lowpass_filter_width * (double)p->D / base_freq);
    p->W = 2 * p->width + p->D;
    p->kern.assign((size_t)p->U * p->W, 0.0f);
    if (beta <= 0) beta = 14.769656459379492;
    const double scale = base_freq / p->D, i0b = bessel_i0(beta);
    for (int ph = 0; ph < p->U; ++ph)
        for (int k = 0; k < p->W; ++k) {
            double t = (-(double)ph / p->U + (double)(k - p->width) / p->D) * base_freq;
            t = std::max(-(double)lowpass_filter_width, std::min((double)lowpass_filter_width, t));
            double window;
            if (method == 0) {
                const double c = cos(t * kPi / lowpass_filter_width / 2.0);
                window = c * c;
            } else {
                const double r = t / lowpass_filter_width;
                window = bessel_i0(beta * sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
            }
            t *= kPi;
            const double sinc = t == 0.0 ? 1.0 : sin(t) / t;
            p->kern[(size_t)ph * p->W + k] = (float)(sinc * window * scale);
        }
    if (cudaMalloc(reinterpret_cast<void**>(&p->d_kern), p->kern.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(p->d_kern, p->kern.data(), p->kern.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("cudaMalloc/cudaMemcpy of the resampling kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        hmfe_resample_plan_destroy(p);
        return HMFE_ERR_CUDA;
    }
    *plan = p;
    return HMFE_OK;
}

// librosa / torchaudio output length: ceil(n * new / orig)
int64_t hmfe_resample_out_len(const hmfe_resample_plan* p, int64_t n_in) {
    if (!p || n_in < 0) return -1;
    return (n_in * p->U + p->D - 1) / p->D;
}

int hmfe_resample_last_launches(const hmfe_resample_plan* p) { return p ? p->last_launches : 0; }

int hmfe_resample_taps(const hmfe_resample_plan* p, int* n_phases, int* n_taps, float* h_out) {
    HMFE_REQUIRE(p, "NULL plan");
    if (n_phases) *n_phases = p->U;
    if (n_taps) *n_taps = p->W;
    if (h_out) std::copy(p->kern.begin(), p->kern.end(), h_out);
    return HMFE_OK;
}

static int resample_launch(hmfe_resample_plan* p, const void* d_in, bool pcm16, const int64_t* h_in_offsets, int64_t n_clips,
                           float* d_out, void* stream) {
    HMFE_REQUIRE(p && h_in_offsets, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0, "n_clips < 0");
    p->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_in && d_out, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = 3 * (size_t)(n_clips + 1) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = p->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    int64_t* hi = static_cast<int64_t*>(hbuf);
    int64_t* ho = hi + (n_clips + 1);
    int64_t* ht = ho + (n_clips + 1);
    ho[0] = ht[0] = 0;
    // integer up-factors take the register-tiled kernel (tiles count input positions there)
    const bool up = p->D == 1 && (p->U == 2 || p->U == 4 || p->U == 8) &&
                    ((size_t)(p->U + 1) * (p->W + 4) + kUpTileQ + 8) * sizeof(float) <= 96 * 1024;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_in_offsets[i + 1] - h_in_offsets[i];
        HMFE_REQUIRE(n >= 0 && n < (int64_t)1 << 28, "clip %lld has invalid length %lld", (long long)i, (long long)n);
        const int64_t m = hmfe_resample_out_len(p, n);
        hi[i] = h_in_offsets[i];
        ho[i + 1] = ho[i] + m;
        ht[i + 1] = ht[i] + (up ? (n + kUpTileQ - 1) / kUpTileQ : (m + kRsTile - 1) / kRsTile);
    }
    hi[n_clips] = h_in_offsets[n_clips];
    int rc = p->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    ResampleBatch b{};
    b.x = d_in;
    b.y = d_out;
    b.in_off = static_cast<int64_t*>(dbuf);
    b.out_off = b.in_off + (n_clips + 1);
    b.kern = p->d_kern;
    b.n_clips = n_clips;
    b.n_out_total = ho[n_clips];
    b.U = p->U;
    b.D = p->D;
    b.W = p->W;
    b.width = p->width;
    const int64_t tiles = ht[n_clips];
    if (tiles > 0) {
        HMFE_REQUIRE(tiles < (int64_t)INT32_MAX, "resample grid too large");
        b.taps_in_smem = (size_t)p->U * p->W * sizeof(float) <= 64 * 1024;
        const size_t smem = ((b.taps_in_smem ? (size_t)p->U * p->W : 0) + (size_t)((kRsTile + p->U - 1) / p->U + 1) * p->D + p->W) *
                            sizeof(float);
        HMFE_REQUIRE(smem <= 200 * 1024, "resampling ratio %d/%d with %d taps needs %zu bytes of shared memory", p->U, p->D, p->W,
                     smem);
        const int64_t* prefix = b.in_off + 2 * (n_clips + 1);
        if (up) {
            const size_t smem_up = ((size_t)(p->U + 1) * ((p->W + 3) / 4 * 4) + kUpTileQ + 8) * sizeof(float);
            const unsigned ctas = (unsigned)((tiles + kUpTilesPerCta - 1) / kUpTilesPerCta);
#define HMFE_RS_UP(TYPE, UF)                                                                                                \
    {                                                                                                                       \
        HMFE_CHECK_CUDA(cudaFuncSetAttribute(resample_up_kernel<TYPE, UF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_up)); \
        resample_up_kernel<TYPE, UF><<<ctas, 256, smem_up, st>>>(b, prefix, tiles);                                          \
    }
            if (pcm16) {
                if (p->U == 2) HMFE_RS_UP(int16_t, 2) else if (p->U == 4) HMFE_RS_UP(int16_t, 4) else HMFE_RS_UP(int16_t, 8)
            } else {
                if (p->U == 2) HMFE_RS_UP(float, 2) else if (p->U == 4) HMFE_RS_UP(float, 4) else HMFE_RS_UP(float, 8)
            }
#undef HMFE_RS_UP
        } else if (pcm16) {
            HMFE_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            resample_kernel<int16_t><<<(unsigned)tiles, 256, smem, st>>>(b, prefix);
        } else {
            HMFE_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            resample_kernel<float><<<(unsigned)tiles, 256, smem, st>>>(b, prefix);
        }
        HMFE_CHECK_CUDA(cudaGetLastError());
        p->last_launches = 1;
    }
    return p->ring.release(slot, st);
}

int hmfe_resample_batch(hmfe_resample_plan* p, const float* d_in, const int64_t* h_in_offsets, int64_t n_clips,
                        float* d_out, void* stream) {
    return resample_launch(p, d_in, false, h_in_offsets, n_clips, d_out, stream);
}

int hmfe_resample_batch_pcm16(hmfe_resample_plan* p, const int16_t* d_pcm, const int64_t* h_in_offsets, int64_t n_clips,
                              float* d_out, void* stream) {
    return resample_launch(p, d_pcm, true, h_in_offsets, n_clips, d_out, stream);
}

}  // extern "C"
