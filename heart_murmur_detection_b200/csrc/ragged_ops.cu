// Ragged-batch utility kernels: silence trim (frame RMS + first/last index) and the
// pad / split / tile / repeat gather.  All index work is integer and bit-exact; the index
// tables for the gather are computed on the host (frontend.py) and only applied here.
#include <math.h>

#include <algorithm>
#include <new>

#include "api_common.h"
#include "ctx.h"
#include "trim_common.cuh"

namespace hmfe {

// ---------------------------------------------------------------------------------- trim
// librosa.effects.trim(y, top_db, ref=np.max, frame_length=L, hop_length=h)
// (/root/reference/src/util.py:170-172,237-244,338-340,820-822; extract_feature.py:219-221):
// centred frames (zero pad L/2 each side), p_t = mean(x^2) over frame t (float32),
// rms = sqrt(p), db = 10 log10(max(amin^2, rms^2)) - 10 log10(max(amin^2, max rms ^2)),
// non-silent = db > -top_db; start = h * first, end = min(N, h * (last + 1)); none -> (0, 0).

struct TrimBatch {
    const float* wav;
    const int64_t* clip_off;   // [n_clips+1]
    const int64_t* frame_off;  // [n_clips+1]
    float* power;              // [total frames] rms^2 per frame
    int64_t* start_end;        // [n_clips][2]
    int64_t n_clips, n_frames_total;
    int frame_length, hop;
    float top_db;
};

// one warp per frame
__global__ void __launch_bounds__(256) trim_frame_power_kernel(const TrimBatch b) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (b.n_frames_total + n_warps - 1) / n_warps;
    const int64_t f_begin = warp_global * per, f_end = min(b.n_frames_total, f_begin + per);
    if (f_begin >= f_end) return;
    int64_t lo = 0, hi = b.n_clips;  // largest clip with frame_off[clip] <= f_begin
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (b.frame_off[mid] <= f_begin)
            lo = mid;
        else
            hi = mid;
    }
    int64_t clip = lo;
    const int half = b.frame_length / 2;
    for (int64_t f = f_begin; f < f_end; ++f) {
        while (f >= b.frame_off[clip + 1]) ++clip;
        const int64_t c0 = b.clip_off[clip];
        const int n = (int)(b.clip_off[clip + 1] - c0);
        const int t = (int)(f - b.frame_off[clip]);
        const int s0 = t * b.hop - half;
        const float* x = b.wav + c0;
        float acc = 0.0f;
        for (int i = lane; i < b.frame_length; i += 32) {
            const int j = s0 + i;
            const float v = (j >= 0 && j < n) ? __ldg(x + j) : 0.0f;
            acc = fmaf(v, v, acc);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) {
            const float rms = sqrtf(acc / (float)b.frame_length);
            b.power[f] = rms * rms;
        }
    }
}

// one CTA per clip: clip max, then first / last frame above the threshold
__global__ void __launch_bounds__(256) trim_index_kernel(const TrimBatch b) {
    __shared__ float s_f[8];
    __shared__ int s_i[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t clip = blockIdx.x; clip < b.n_clips; clip += gridDim.x) {
        const int64_t f0 = b.frame_off[clip];
        const int T = (int)(b.frame_off[clip + 1] - f0);
        const int n = (int)(b.clip_off[clip + 1] - b.clip_off[clip]);
        const float* p = b.power + f0;
        float m = 0.0f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) m = fmaxf(m, p[t]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) s_f[warp] = m;
        __syncthreads();
        m = s_f[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_f[w]);
        // amplitude_to_db(rms, ref=max, amin=1e-5): power_to_db(rms^2, ref=max^2, amin=1e-10).
        // The scalar reference term is evaluated in float64 and rounded once (numpy scalar rules).
        const float amin2 = 1e-10f;
        const float ref_db = (float)(10.0 * log10(fmax(1e-10, (double)m)));
        int first = INT_MAX, last = -1;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const float db = 10.0f * log10f(fmaxf(amin2, p[t])) - ref_db;
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

// frame_length == 2 * hop (the reference's 1600 / 800): every sample is read ONCE into per-hop energy sums
// e_h = sum x^2 over [h*hop, (h+1)*hop); frame t is (e_{t-1} + e_t) / frame_length (trim_index_hop_kernel).
// One warp per hop block.
struct HopEnergyBatch {
    const float* wav;
    const int64_t* clip_off;  // [n_clips+1]
    const int64_t* hop_off;   // [n_clips+1]
    float* energy;            // [total hop blocks]
    int64_t n_clips, n_blocks_total;
    int hop;
};

__global__ void __launch_bounds__(256) hop_energy_kernel(const HopEnergyBatch b) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (b.n_blocks_total + n_warps - 1) / n_warps;
    const int64_t h_begin = warp_global * per, h_end = min(b.n_blocks_total, h_begin + per);
    if (h_begin >= h_end) return;
    int64_t lo = 0, hi = b.n_clips;  // largest clip with hop_off[clip] <= h_begin
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (b.hop_off[mid] <= h_begin)
            lo = mid;
        else
            hi = mid;
    }
    int64_t clip = lo;
    for (int64_t h = h_begin; h < h_end; ++h) {
        while (h >= b.hop_off[clip + 1]) ++clip;
        const int64_t c0 = b.clip_off[clip];
        const int n = (int)(b.clip_off[clip + 1] - c0);
        const int s0 = (int)(h - b.hop_off[clip]) * b.hop;
        const int len = min(b.hop, n - s0);
        const float* x = b.wav + c0 + s0;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        int i = lane;
        for (; i + 96 < len; i += 128) {  // four independent loads in flight
            const float v0 = __ldg(x + i), v1 = __ldg(x + i + 32), v2 = __ldg(x + i + 64), v3 = __ldg(x + i + 96);
            a0 = fmaf(v0, v0, a0);
            a1 = fmaf(v1, v1, a1);
            a2 = fmaf(v2, v2, a2);
            a3 = fmaf(v3, v3, a3);
        }
        for (; i < len; i += 32) {
            const float v = __ldg(x + i);
            a0 = fmaf(v, v, a0);
        }
        float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) b.energy[h] = acc;
    }
}

// ---------------------------------------------------------------------------------- gather
// out chunk c, element i:
//   i <  a_end : src[src_off + (a_phase + i) mod period]           (tiled / repeated part)
//   i <  b_end : src[src_off + b_start + (i - a_end)]              (straight copy)
//   else       : 0
// covers _zero_padding, _duplicate_padding, 50 %-overlap framing, truncation and
// non-overlapping chunking (/root/reference/src/util.py:504-620, 257-259; extract_feature.py:250-259).
constexpr int kGatherTile = 4096;

__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     const hmfe_gather_desc* __restrict__ descs, int tiles_per_chunk) {
    const int64_t chunk = blockIdx.x / tiles_per_chunk;
    const int tile = (int)(blockIdx.x - chunk * tiles_per_chunk);
    const hmfe_gather_desc d = descs[chunk];
    const int t0 = tile * kGatherTile;
    if (t0 >= d.len) return;
    const float* s = src + d.src_off;
    float* o = dst + d.dst_off;
    const int t1 = min(d.len, t0 + kGatherTile);
    // position inside the repeated source: advanced by 256 per element instead of a division per element;
    // four independent loads are issued before their stores
    const unsigned period = (unsigned)d.period;
    unsigned ph = (unsigned)(((unsigned long long)d.a_phase + (unsigned)(t0 + threadIdx.x)) % period);
    const unsigned step = 256u % period;
    for (int i0 = t0 + threadIdx.x; i0 < t1; i0 += 1024) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 256 * u;
            v[u] = 0.0f;
            if (i < t1) {
                if (i < d.a_end)
                    v[u] = __ldg(s + ph);
                else if (i < d.b_end)
                    v[u] = __ldg(s + d.b_start + (i - d.a_end));
            }
            ph += step;
            if (ph >= period) ph -= period;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 256 * u;
            if (i < t1) o[i] = v[u];
        }
    }
}

// ---------------------------------------------------------------------------------- device-side planner
// Control flow of get_entire_signal_librosa after the silence trim (/root/reference/src/util.py:248-259), evaluated
// on the device from the trim indices so that the host never waits for them: duration test, "too short" (dropped or
// padded to input_sec), cut at max_sec, pad descriptors (_zero_padding / _duplicate_padding, src/util.py:504-575),
// then the exclusive scans that place every clip's feature rows and work items.  Same integer formulas as
// frontend.plan_* / pipeline._entire_signal_fast (bit-exact index work); one CTA, n_clips is a few thousand.
struct EntirePlanArgs {
    const int64_t* clip_off;   // [n+1] device copy of the host offsets
    const int64_t* start_end;  // [n][2] trim indices, clip relative
    int64_t n_clips;
    double sample_rate, input_sec, max_sec;  // max_sec <= 0: no cut
    int pad, pad_zero;                       // pad: pad short clips; pad_zero: types == "zero" (else "repeat"); 2: zero padding,
                                             // and clips that only get trailing zeros (>= half the length) are NOT copied:
                                             // out_len keeps their own length, frame_off counts the padded one, and the
                                             // log-mel fetch reads zeros behind the clip end anyway
    int hop, item_frames;
    int64_t dst_base;  // first element of the padded copies in their buffer
    int alt;           // padded copies live in a second buffer: start = -(offset + 1)
    int64_t *out_start, *out_len, *frame_off, *item_prefix;  // [n], [n], [n+1], [n+1]
    hmfe_gather_desc* descs;                                 // [n] compact list of the padded copies to materialise
    int64_t* n_descs;                                        // how many of them
};

__global__ void __launch_bounds__(1024) entire_plan_kernel(const EntirePlanArgs a) {
    __shared__ long long s_scan[3][32];
    __shared__ long long s_carry[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long L = (long long)(a.input_sec * a.sample_rate);  // int(input_sec * sample_rate)
    if (threadIdx.x < 3) s_carry[threadIdx.x] = 0;
    if (threadIdx.x == 0) a.frame_off[0] = a.item_prefix[0] = 0;
    __syncthreads();
    for (int64_t base = 0; base < a.n_clips; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        long long T = 0, items = 0, padded = 0, n = 0, len = 0, start = 0;
        bool valid = false, is_short = false;
        if (i < a.n_clips) {
            const long long s0 = a.start_end[2 * i], s1 = a.start_end[2 * i + 1];
            n = s1 - s0;
            const double dur = (double)n / a.sample_rate;
            is_short = dur < a.input_sec;
            valid = !is_short || (a.pad && n > 0);
            start = a.clip_off[i] + s0;
            len = n;
            if (a.max_sec > 0 && dur > a.max_sec) len = min(len, (long long)(a.max_sec * a.sample_rate));
            padded = (valid && is_short) ? 1 : 0;
            long long len_frames = len;
            if (padded) {
                len_frames = L;
                // _zero_padding with frac >= 0.5: the clip followed by zeros (src/util.py:518-520)
                if (a.pad_zero == 2 && !((double)n / (double)L < 0.5))
                    padded = 0;  // a view: out_len = n
                else
                    len = L;
            }
            T = valid ? 1 + len_frames / a.hop : 0;
            items = (T + a.item_frames - 1) / a.item_frames;
        }
        // block-wide exclusive scans of (T, items, padded)
        long long v[3] = {T, items, padded}, incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            long long x = v[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            incl[k] = x;
            if (lane == 31) s_scan[k][warp] = x;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                long long x = lane < (int)(blockDim.x >> 5) ? s_scan[k][lane] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const long long y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                s_scan[k][lane] = x;  // inclusive over warps
            }
        }
        __syncthreads();
        long long excl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) excl[k] = s_carry[k] + (warp ? s_scan[k][warp - 1] : 0) + incl[k] - v[k];
        if (i < a.n_clips) {
            hmfe_gather_desc d{};
            if (padded) {
                const long long slot = a.dst_base + L * excl[2];
                d.src_off = start;
                d.dst_off = slot;
                d.len = (int)L;
                d.period = (int)n;
                if (a.pad_zero) {  // _equally_slice_pad_sample -> one slice -> _zero_padding
                    const bool tile = (double)n / (double)L < 0.5;
                    const long long copies = (L - 1) / n;
                    d.a_end = tile ? (int)(copies * n) : 0;
                    d.b_end = tile ? (int)(copies * n) : (int)n;
                } else {  // _duplicate_padding: source at the end, tail of the doubled clip in front
                    const long long left = L - n;
                    long long len_aug = n;
                    while (len_aug < left) len_aug *= 2;
                    d.a_end = (int)left;
                    d.a_phase = (int)((len_aug - left) % n);
                    d.b_end = (int)L;
                }
                start = a.alt ? -(slot + 1) : slot;
            }
            if (padded) a.descs[excl[2]] = d;
            a.out_start[i] = start;
            a.out_len[i] = valid ? len : 0;
            a.frame_off[i + 1] = excl[0] + T;
            a.item_prefix[i + 1] = excl[1] + items;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) s_carry[k] = excl[k] + v[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_descs = s_carry[2];
}

// gather from a compact descriptor list in device memory whose length is only known on the device: a fixed grid
// walks the (descriptor, tile) pairs
__global__ void __launch_bounds__(256) gather_device_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            const hmfe_gather_desc* __restrict__ descs,
                                                            const int64_t* __restrict__ n_descs, int tiles_per_chunk) {
    const int64_t n_work = *n_descs * tiles_per_chunk;
    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int64_t chunk = w / tiles_per_chunk;
        const int tile = (int)(w - chunk * tiles_per_chunk);
        const hmfe_gather_desc d = descs[chunk];
        const int t0 = tile * kGatherTile;
        if (t0 >= d.len) continue;
        const float* s = src + d.src_off;
        float* o = dst + d.dst_off;
        const int t1 = min(d.len, t0 + kGatherTile);
        const unsigned period = (unsigned)d.period;
        unsigned ph = (unsigned)(((unsigned long long)d.a_phase + (unsigned)(t0 + threadIdx.x)) % period);
        const unsigned step = 256u % period;
        for (int i0 = t0 + threadIdx.x; i0 < t1; i0 += 1024) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 256 * u;
                v[u] = 0.0f;
                if (i < t1) {
                    if (i < d.a_end)
                        v[u] = __ldg(s + ph);
                    else if (i < d.b_end)
                        v[u] = __ldg(s + d.b_start + (i - d.a_end));
                }
                ph += step;
                if (ph >= period) ph -= period;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 256 * u;
                if (i < t1) o[i] = v[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------- PCM16 decode
// The sample conversion inside librosa.load / soundfile.read(dtype="float32") for 16-bit PCM WAV
// payloads (/root/reference/src/util.py:153,222,323,391,805): x = int16 / 32768 (exact in float32).
__global__ void __launch_bounds__(256) pcm16_decode_kernel(const int16_t* __restrict__ pcm, float* __restrict__ out,
                                                           int64_t n, int64_t n8) {
    const float k = 1.0f / 32768.0f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(pcm) + i);  // 8 samples
        const int w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = (float)(short)(w[j] & 0xffff) * k;
            f[2 * j + 1] = (float)(short)(w[j] >> 16) * k;
        }
        float4* o = reinterpret_cast<float4*>(out) + 2 * i;
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int64_t i = 8 * n8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)pcm[i] * k;
}

// ---------------------------------------------------------------------------------- multicast push
// One store instruction, N destinations: `mc_dst` is the NVLink-switch multicast mapping of a buffer
// that exists on every GPU of the node (torch symmetric memory); multimem.st makes the switch
// replicate the 16 bytes into all of them.  An all-gather then costs every GPU ONE egress copy of
// its block instead of N-1.  Few CTAs are enough (the transfer is NVLink-bound), so the kernels of
// the next step keep the SMs.
__global__ void __launch_bounds__(512) multicast_push_kernel(const float4* __restrict__ src, float4* mc_dst, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(src + i);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    }
    __threadfence_system();
}

}  // namespace hmfe

using namespace hmfe;

extern "C" {

int hmfe_multicast_push(const float* d_src, float* mc_dst, int64_t n, int n_ctas, void* stream) {
    HMFE_REQUIRE(n >= 0 && n % 4 == 0, "n must be a non-negative multiple of 4");
    if (n == 0) return HMFE_OK;
    HMFE_REQUIRE(d_src && mc_dst, "NULL pointer");
    HMFE_REQUIRE(((reinterpret_cast<uintptr_t>(d_src) | reinterpret_cast<uintptr_t>(mc_dst)) & 15) == 0,
                 "pointers must be 16-byte aligned");
    HMFE_REQUIRE(n_ctas >= 1 && n_ctas <= 1024, "n_ctas out of range");
    multicast_push_kernel<<<n_ctas, 512, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(d_src), reinterpret_cast<float4*>(mc_dst), n / 4);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

int hmfe_pcm16_decode(const int16_t* d_pcm, int64_t n, float* d_out, void* stream) {
    HMFE_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return HMFE_OK;
    HMFE_REQUIRE(d_pcm && d_out, "NULL device pointer");
    // 16-byte accesses when both pointers allow it, scalar otherwise
    const bool aligned = (reinterpret_cast<uintptr_t>(d_pcm) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
    const int64_t n8 = aligned ? n / 8 : 0;
    const int64_t threads = std::max<int64_t>(1, aligned ? n8 : n);
    const int grid = (int)std::min<int64_t>((threads + 255) / 256, (int64_t)device_sm_count() * 16);
    pcm16_decode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_pcm, d_out, n, n8);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}


int hmfe_ctx_create(hmfe_ctx** ctx) {
    HMFE_REQUIRE(ctx != nullptr, "ctx is NULL");
    *ctx = new (std::nothrow) hmfe_ctx();
    HMFE_REQUIRE(*ctx != nullptr, "out of host memory");
    (*ctx)->sm_count = device_sm_count();
    return HMFE_OK;
}

// ---- caller-provided workspace ("never allocate, caller workspace")
int hmfe_ctx_set_workspace(hmfe_ctx* ctx, void* d_workspace, size_t bytes) {
    HMFE_REQUIRE(ctx, "NULL ctx");
    HMFE_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (ctx->scratch && !ctx->scratch_external) HMFE_CHECK_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = d_workspace;
    ctx->scratch_cap = d_workspace ? bytes : 0;
    ctx->scratch_external = d_workspace != nullptr;
    return HMFE_OK;
}

int hmfe_ctx_reserve(hmfe_ctx* ctx, int64_t max_clips) {
    HMFE_REQUIRE(ctx && max_clips >= 0, "bad arguments");
    // the largest per-call descriptor block of any ctx stage: the IIR stages upload 4 int64 arrays of n_clips + 1
    // entries plus the (2S)^2 transition matrix of the exact scan (S <= 8); gather / crop records are 40 / 24 bytes
    const size_t bytes = (size_t)(max_clips + 1) * 48 + 16 * 16 * sizeof(double) + 4096;
    int rc = ctx->ring.reserve(bytes);
    if (rc != HMFE_OK) return rc;
    ctx->ring.forbid_growth();
    return HMFE_OK;
}

int64_t hmfe_trim_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int frame_length, int hop_length) {
    if (!h_offsets || n_clips < 0 || frame_length < 2 || hop_length < 1) return -1;
    const bool by_hop = frame_length == 2 * hop_length;
    int64_t units = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_offsets[i + 1] - h_offsets[i];
        units += by_hop ? (n + hop_length - 1) / hop_length : hmfe_trim_num_frames(n, frame_length, hop_length);
    }
    return std::max<int64_t>(1, units) * (int64_t)sizeof(float);
}

int64_t hmfe_iir_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int n_sections, int hop_length) {
    if (!h_offsets || n_clips < 0 || n_sections < 1 || n_sections > 8) return -1;
    // upper bound over the algorithms hmfe_iir_sos_batch / hmfe_iir_sos_trim_batch may choose: hop energies of the fused
    // trim (8 floats per hop group, one extra group per clip for the 16-byte alignment shift) or the carry vectors of the
    // exact scan (two 2S-vectors of doubles per chunk; chunks of 512 samples for batches of 32 Mi samples and more, of
    // 128 samples below that: the rule of iir_impl in iir.cu)
    int64_t groups = 0, chunks = 0;
    const int64_t total = n_clips > 0 ? h_offsets[n_clips] - h_offsets[0] : 0;
    const int64_t c_exact = total >= ((int64_t)32 << 20) ? 512 : 128;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_offsets[i + 1] - h_offsets[i];
        if (hop_length > 0) groups += (n + 3 + hop_length - 1) / hop_length + 1;
        chunks += (n + 3 + c_exact - 1) / c_exact + 1;
    }
    const int64_t a = std::max<int64_t>(1, groups) * 8 * (int64_t)sizeof(float);
    const int64_t b = 2 * chunks * 2 * n_sections * (int64_t)sizeof(double);
    return std::max(a, b);
}

void hmfe_ctx_destroy(hmfe_ctx* ctx) {
    if (!ctx) return;
    if (ctx->scratch && !ctx->scratch_external) cudaFree(ctx->scratch);
    for (auto& r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    delete ctx;
}

int hmfe_ctx_last_launches(const hmfe_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

int hmfe_ctx_set_profile(hmfe_ctx* ctx, int enable) {
    HMFE_REQUIRE(ctx, "NULL ctx");
    ctx->profile = enable != 0;
    return HMFE_OK;
}

int hmfe_ctx_profile_ms(hmfe_ctx* ctx, double* ms_by_kernel, int* launches_by_kernel) {
    HMFE_REQUIRE(ctx && ms_by_kernel && launches_by_kernel, "NULL argument");
    for (int i = 0; i < HMFE_K_COUNT; ++i) {
        ms_by_kernel[i] = 0.0;
        launches_by_kernel[i] = 0;
    }
    for (auto& r : ctx->prof) {
        float ms = 0.0f;
        HMFE_CHECK_CUDA(cudaEventSynchronize(r.b));
        HMFE_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_by_kernel[r.id] += ms;
        launches_by_kernel[r.id] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    return HMFE_OK;
}

int64_t hmfe_trim_num_frames(int64_t n_samples, int frame_length, int hop_length) {
    if (n_samples < 0 || frame_length < 1 || hop_length < 1) return -1;
    const int64_t padded = n_samples + 2 * (int64_t)(frame_length / 2);
    return padded < frame_length ? 0 : 1 + (padded - frame_length) / hop_length;
}

int hmfe_trim_batch(hmfe_ctx* ctx, const float* d_wav, const int64_t* h_offsets, int64_t n_clips, int frame_length,
                    int hop_length, float top_db, int64_t* d_start_end, void* stream) {
    HMFE_REQUIRE(ctx && h_offsets, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0 && frame_length >= 2 && hop_length >= 1, "bad trim arguments");
    ctx->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_wav && d_start_end, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t desc_bytes = 2 * (size_t)(n_clips + 1) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(desc_bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    int64_t* hc = static_cast<int64_t*>(hbuf);
    int64_t* hf = hc + (n_clips + 1);
    const bool by_hop = frame_length == 2 * hop_length;  // every sample read once (hop energies) instead of twice (frames)
    hf[0] = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t n = h_offsets[i + 1] - h_offsets[i];
        HMFE_REQUIRE(n >= 0 && n < (int64_t)1 << 30, "clip %lld has invalid length %lld", (long long)i, (long long)n);
        hc[i] = h_offsets[i];
        hf[i + 1] = hf[i] + (by_hop ? (n + hop_length - 1) / hop_length : hmfe_trim_num_frames(n, frame_length, hop_length));
    }
    hc[n_clips] = h_offsets[n_clips];
    int rc = ctx->ring.upload(slot, desc_bytes, st);
    if (rc != HMFE_OK) return rc;
    const int64_t* d_clip_off = static_cast<int64_t*>(dbuf);
    const int64_t* d_unit_off = d_clip_off + (n_clips + 1);
    const int64_t n_units = hf[n_clips];
    rc = ctx->reserve_scratch((size_t)std::max<int64_t>(1, n_units) * sizeof(float));
    if (rc != HMFE_OK) return rc;
    float* d_units = static_cast<float*>(ctx->scratch);
    const int index_grid = (int)std::min<int64_t>(n_clips, (int64_t)ctx->sm_count * 8);
    if (by_hop) {
        if (n_units > 0) {
            HopEnergyBatch hb{d_wav, d_clip_off, d_unit_off, d_units, n_clips, n_units, hop_length};
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_units + 7) / 8, (int64_t)ctx->sm_count * 8));
            ctx->prof_begin(HMFE_K_TRIM_POWER, st);
            hop_energy_kernel<<<grid, 256, 0, st>>>(hb);
            HMFE_CHECK_CUDA(cudaGetLastError());
            ctx->prof_end(st);
            ctx->last_launches++;
        }
        TrimHopBatch tb{d_units, d_clip_off, d_unit_off, d_start_end, n_clips, hop_length, top_db};
        ctx->prof_begin(HMFE_K_TRIM_INDEX, st);
        trim_index_hop_kernel<<<index_grid, 256, 0, st>>>(tb);
        HMFE_CHECK_CUDA(cudaGetLastError());
        ctx->prof_end(st);
        ctx->last_launches++;
        return ctx->ring.release(slot, st);
    }
    TrimBatch b{};
    b.wav = d_wav;
    b.clip_off = d_clip_off;
    b.frame_off = d_unit_off;
    b.n_clips = n_clips;
    b.n_frames_total = n_units;
    b.frame_length = frame_length;
    b.hop = hop_length;
    b.top_db = top_db;
    b.start_end = d_start_end;
    b.power = d_units;
    if (b.n_frames_total > 0) {
        const int64_t warps_wanted = b.n_frames_total;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps_wanted + 7) / 8, (int64_t)ctx->sm_count * 8));
        ctx->prof_begin(HMFE_K_TRIM_POWER, st);
        trim_frame_power_kernel<<<grid, 256, 0, st>>>(b);
        HMFE_CHECK_CUDA(cudaGetLastError());
        ctx->prof_end(st);
        ctx->last_launches++;
    }
    ctx->prof_begin(HMFE_K_TRIM_INDEX, st);
    trim_index_kernel<<<index_grid, 256, 0, st>>>(b);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches++;
    return ctx->ring.release(slot, st);
}

int hmfe_gather_batch(hmfe_ctx* ctx, const float* d_src, float* d_dst, const hmfe_gather_desc* h_descs,
                      int64_t n_chunks, void* stream) {
    HMFE_REQUIRE(ctx && (h_descs || n_chunks == 0), "NULL argument");
    HMFE_REQUIRE(n_chunks >= 0, "n_chunks < 0");
    ctx->last_launches = 0;
    if (n_chunks == 0) return HMFE_OK;
    HMFE_REQUIRE(d_src && d_dst, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int max_len = 0;
    for (int64_t i = 0; i < n_chunks; ++i) {
        const hmfe_gather_desc& d = h_descs[i];
        HMFE_REQUIRE(d.len >= 0 && d.a_end >= 0 && d.a_end <= d.b_end && d.b_end <= d.len && d.period >= 1 &&
                         d.a_phase >= 0 && d.b_start >= 0 && d.src_off >= 0 && d.dst_off >= 0,
                     "gather descriptor %lld is inconsistent", (long long)i);
        max_len = std::max(max_len, d.len);
    }
    if (max_len == 0) return HMFE_OK;
    const size_t bytes = (size_t)n_chunks * sizeof(hmfe_gather_desc);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_descs, bytes);
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    const int tiles = (max_len + kGatherTile - 1) / kGatherTile;
    HMFE_REQUIRE(n_chunks * tiles < (int64_t)INT32_MAX, "gather grid too large");
    ctx->prof_begin(HMFE_K_GATHER, st);
    gather_kernel<<<(unsigned)(n_chunks * tiles), 256, 0, st>>>(d_src, d_dst, static_cast<hmfe_gather_desc*>(dbuf), tiles);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

int hmfe_entire_plan_batch(hmfe_ctx* ctx, const int64_t* h_offsets, int64_t n_clips, const int64_t* d_start_end,
                           int sample_rate, double input_sec, int pad, int pad_zero, double max_sec, int hop, int item_frames,
                           int64_t dst_base, int alt, int64_t* d_desc, hmfe_gather_desc* d_gather, void* stream) {
    HMFE_REQUIRE(ctx && h_offsets, "NULL argument");
    HMFE_REQUIRE(n_clips >= 0 && sample_rate > 0 && input_sec > 0 && hop >= 1 && item_frames >= 1, "bad plan arguments");
    ctx->last_launches = 0;
    if (n_clips == 0) return HMFE_OK;
    HMFE_REQUIRE(d_start_end && d_desc && d_gather, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)(n_clips + 1) * sizeof(int64_t);
    void *hbuf = nullptr, *dbuf = nullptr;
    const int slot = ctx->ring.acquire(bytes, &hbuf, &dbuf);
    if (slot < 0) return slot;
    memcpy(hbuf, h_offsets, bytes);
    int rc = ctx->ring.upload(slot, bytes, st);
    if (rc != HMFE_OK) return rc;
    EntirePlanArgs a{};
    a.clip_off = static_cast<int64_t*>(dbuf);
    a.start_end = d_start_end;
    a.n_clips = n_clips;
    a.sample_rate = (double)sample_rate;
    a.input_sec = input_sec;
    a.max_sec = max_sec;
    a.pad = pad;
    a.pad_zero = pad_zero;
    a.hop = hop;
    a.item_frames = item_frames;
    a.dst_base = dst_base;
    a.alt = alt;
    a.out_start = d_desc;
    a.out_len = d_desc + n_clips;
    a.frame_off = d_desc + 2 * n_clips;
    a.item_prefix = d_desc + 3 * n_clips + 1;
    a.n_descs = d_desc + 4 * n_clips + 2;
    a.descs = d_gather;
    entire_plan_kernel<<<1, 1024, 0, st>>>(a);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->last_launches = 1;
    return ctx->ring.release(slot, st);
}

int hmfe_gather_device(hmfe_ctx* ctx, const float* d_src, float* d_dst, const hmfe_gather_desc* d_descs,
                       const int64_t* d_n_descs, int64_t max_chunks, int max_len, void* stream) {
    HMFE_REQUIRE(ctx, "NULL argument");
    HMFE_REQUIRE(max_chunks >= 0 && max_len >= 0, "bad arguments");
    ctx->last_launches = 0;
    if (max_chunks == 0 || max_len == 0) return HMFE_OK;
    HMFE_REQUIRE(d_src && d_dst && d_descs && d_n_descs, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = (max_len + kGatherTile - 1) / kGatherTile;
    const unsigned grid = (unsigned)std::min<int64_t>(max_chunks * tiles, (int64_t)ctx->sm_count * 16);
    ctx->prof_begin(HMFE_K_GATHER, st);
    gather_device_kernel<<<grid, 256, 0, st>>>(d_src, d_dst, d_descs, d_n_descs, tiles);
    HMFE_CHECK_CUDA(cudaGetLastError());
    ctx->prof_end(st);
    ctx->last_launches = 1;
    return HMFE_OK;
}

}  // extern "C"
