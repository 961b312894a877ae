// Lane-level phases of the fused STFT-power + mel kernel (n_fft = 1024).
//
// One warp transforms 2*NV consecutive frames per iteration (NV = 1 for V=float,
// 2 for V=f32x2): frames (a, b) are packed as real / imaginary part of one complex
// 1024-point FFT, split 1024 = 32 x 32:
//   pass 1  lane = n1, registers = n2 : 32-point FFT over n2 (in registers)
//   twiddle W_1024^(n1*k2), per-lane plane read from shared memory
//   exchange through a padded [32][33] shared-memory tile (conflict free)
//   pass 2  lane = k2, registers = n1 : 32-point FFT over n1 -> bin k = k2 + 32*k1
//   separation |X_a[k]|^2, |X_b[k]|^2 from Z[k] and Z[1024-k] (partner lane via shuffle)
//   sparse banded mel projection from the power tile in shared memory
// Every phase is a __host__ __device__ template taking the lane id explicitly, so
// csrc/host_check.cu can run the identical arithmetic on the CPU.
#pragma once
#include "fft_core.cuh"

namespace hmfe {

constexpr int kNfft = 1024;
constexpr int kBinsPad = 576;   // power tile rows (>= 513 + slack), in exchange-tile elements
constexpr int kXStride = 33;    // exchange tile row stride (elements)

// exchange / power tile element: (re, im) or (p_a, p_b)
template <typename V>
struct xelem;
template <>
struct __align__(8) xelem<float> {
    float a, b;
};
template <>
struct __align__(16) xelem<f32x2> {
    f32x2 a, b;
};

// Store the two halves of a tile element with separate instructions.  The halves live in
// unrelated registers; one wide store would need register moves to line them up first
// (4 MOV per STS.128 in the packed kernel), two half-width stores need none and cost the same
// number of shared-memory wavefronts.
#if defined(__CUDA_ARCH__)
HMFE_D void store_halves(xelem<float>* p, float a, float b) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
    asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(addr), "f"(b) : "memory");
}
HMFE_D void store_halves(xelem<f32x2>* p, f32x2 a, f32x2 b) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a.x), "f"(a.y) : "memory");
    asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(addr), "f"(b.x), "f"(b.y) : "memory");
}
#else
template <typename V>
inline void store_halves(xelem<V>* p, V a, V b) {
    p->a = a;
    p->b = b;
}
#endif

// ---- phase 1: windowed samples -> bit-reversed registers.  `fetch(t, n)` returns sample n
// (0..1023) of transform t's frame a (im=false) / frame b (im=true), already bounds-handled.
template <typename V, typename Fetch>
HMFE_HD void load_window(int lane, const float* __restrict__ win, Fetch fetch, V (&re)[32], V (&im)[32]) {
    constexpr int NV = lanes_of<V>::value;
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
        const int n = lane + 32 * n2;
        const float w = win[n];
        V xa, xb;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            vput(xa, t, fetch(t, false, n));
            vput(xb, t, fetch(t, true, n));
        }
        re[brev(n2, 5)] = vmuls(xa, w);
        im[brev(n2, 5)] = vmuls(xb, w);
    }
}

// ---- twiddle: Y[k2] *= W^(lane*k2), plane[k2*32 + lane] = (cos, -sin)
template <typename V>
HMFE_HD void apply_twiddle(int lane, const float2* __restrict__ plane, V (&re)[32], V (&im)[32]) {
#pragma unroll
    for (int k2 = 1; k2 < 32; ++k2) {
        const float2 w = plane[k2 * 32 + lane];
        const V nr = vfnmas(im[k2], w.y, vmuls(re[k2], w.x));  // re*wr - im*wi
        const V ni = vfmas(im[k2], w.x, vmuls(re[k2], w.y));   // re*wi + im*wr
        re[k2] = nr;
        im[k2] = ni;
    }
}

template <typename V>
HMFE_HD void exchange_store(int lane, xelem<V>* tile, const V (&re)[32], const V (&im)[32]) {
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) store_halves(&tile[k2 * kXStride + lane], re[k2], im[k2]);
}
// loads into bit-reversed positions, ready for the second DIT pass
template <typename V>
HMFE_HD void exchange_load(int lane, const xelem<V>* tile, V (&re)[32], V (&im)[32]) {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
        const xelem<V> e = tile[lane * kXStride + n1];
        re[brev(n1, 5)] = e.a;
        im[brev(n1, 5)] = e.b;
    }
}

// ---- separation
// power of both packed frames at bin lane + 32*k1 given own (zr, zi) and partner (pr, pi)
template <typename V>
HMFE_HD xelem<V> frame_powers(V zr, V zi, V pr, V pi) {
    const V u = vadd(zr, pr), v = vsub(zi, pi);
    const V s = vadd(zi, pi), d = vsub(pr, zr);
    xelem<V> o;
    o.a = vfma(v, v, vmul(u, u));
    o.b = vfma(d, d, vmul(s, s));
    return o;
}

// ---- mel: accumulate one slot for this lane from the power tile.  `trip` is a multiple of 8;
// two independent accumulators per output halve the dependent-FMA chain.
template <typename V>
HMFE_HD void mel_slot(int lane, const xelem<V>* __restrict__ ptile, const float* __restrict__ w, int start, int trip,
                      V& acc_a, V& acc_b) {
    V a0 = V{}, a1 = V{}, b0 = V{}, b1 = V{};
    const xelem<V>* p = ptile + start;
    const float* wl = w + lane;
    for (int i = 0; i < trip; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const float w0 = wl[(i + j) * 32], w1 = wl[(i + j + 1) * 32];
            const xelem<V> e0 = p[i + j], e1 = p[i + j + 1];
            a0 = vfmas(e0.a, w0, a0);
            b0 = vfmas(e0.b, w0, b0);
            a1 = vfmas(e1.a, w1, a1);
            b1 = vfmas(e1.b, w1, b1);
        }
    }
    acc_a = vadd(a0, a1);
    acc_b = vadd(b0, b1);
}

// =====================================================================================
// "pair" variant: ONE complex transform (two real frames) per warp iteration, with the packed
// FP32 instructions applied across ELEMENTS instead of across transforms: register q holds the
// elements at DIT positions q and q+16.  Stages M = 2..16 of the 32-point DIT act on the two
// halves alike (= fft_dit<16, f32x2>), only the last stage combines the halves of a register.
// Half the registers per warp of the "packed" variant (64 instead of 128 for the FFT data), so
// twice as many warps fit on an SM to hide the shared-memory and dependency latencies.
// =====================================================================================
constexpr int kPStride = 34;             // exchange-plane row stride in floats (even: 8-byte row pairs)
constexpr int kPlane = 32 * kPStride;    // floats per plane (re plane, im plane)

template <int P>
struct pair_last_stage {
    static HMFE_HD void run(f32x2 (&re)[16], f32x2 (&im)[16]) {
        butterfly<P, float>(re[P].x, im[P].x, re[P].y, im[P].y);  // positions P and P+16, twiddle exp(-2 pi i P / 32)
        if constexpr (P + 1 < 16) pair_last_stage<P + 1>::run(re, im);
    }
};

// in: register q = DIT positions (q, q+16) = input elements (2 brev4(q), 2 brev4(q) + 1); out: register p = bins (p, p+16)
HMFE_HD void fft32_paired(f32x2 (&re)[16], f32x2 (&im)[16]) {
    fft_dit<16, f32x2>(re, im);
    pair_last_stage<0>::run(re, im);
}

// windowed samples of frame a (re) / frame b (im): `fetch(second, n)` returns sample n of the frame
template <typename Fetch>
HMFE_HD void pair_load_window(int lane, const float* __restrict__ win, Fetch fetch, f32x2 (&re)[16], f32x2 (&im)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = lane + 32 * (2 * brev(q, 4));
        const float w0 = win[n], w1 = win[n + 32];
        re[q] = vmul(f32x2{fetch(false, n), fetch(false, n + 32)}, f32x2{w0, w1});
        im[q] = vmul(f32x2{fetch(true, n), fetch(true, n + 32)}, f32x2{w0, w1});
    }
}

// plane[p*32 + lane] = (cos_p, cos_{p+16}, -sin_p, -sin_{p+16}) of 2 pi lane k2 / 1024
HMFE_HD void pair_twiddle(int lane, const float4* __restrict__ plane, f32x2 (&re)[16], f32x2 (&im)[16]) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const float4 w = plane[p * 32 + lane];
        const f32x2 wr{w.x, w.y}, wi{w.z, w.w};
        const f32x2 nr = vsub(vmul(re[p], wr), vmul(im[p], wi));
        const f32x2 ni = vfma(re[p], wi, vmul(im[p], wr));
        re[p] = nr;
        im[p] = ni;
    }
}

HMFE_HD void pair_exchange_store(int lane, float* __restrict__ pre, float* __restrict__ pim, const f32x2 (&re)[16],
                                 const f32x2 (&im)[16]) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        pre[p * kPStride + lane] = re[p].x;
        pre[(p + 16) * kPStride + lane] = re[p].y;
        pim[p * kPStride + lane] = im[p].x;
        pim[(p + 16) * kPStride + lane] = im[p].y;
    }
}
// row `lane`, elements (2m, 2m+1) -> register q with m = brev4(q): one 8-byte load per plane
HMFE_HD void pair_exchange_load(int lane, const float* __restrict__ pre, const float* __restrict__ pim, f32x2 (&re)[16],
                                f32x2 (&im)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int c = 2 * brev(q, 4);
        re[q] = *reinterpret_cast<const f32x2*>(pre + lane * kPStride + c);
        im[q] = *reinterpret_cast<const f32x2*>(pim + lane * kPStride + c);
    }
}

// value of Z at second-pass output index k1 (0..31) from the paired registers
HMFE_HD float pair_get(const f32x2 (&v)[16], int k1) { return k1 < 16 ? v[k1].x : v[k1 - 16].y; }

// mel for the two frames of an item at once: the power tile element (p_a, p_b) is one f32x2
HMFE_HD void mel_slot_ab(int lane, const f32x2* __restrict__ ptile, const float* __restrict__ w, int start, int trip,
                         f32x2& acc) {
    f32x2 a0{}, a1{};
    const f32x2* p = ptile + start;
    const float* wl = w + lane;
    for (int i = 0; i < trip; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const float w0 = wl[(i + j) * 32], w1 = wl[(i + j + 1) * 32];
            a0 = vfmas(p[i + j], w0, a0);
            a1 = vfmas(p[i + j + 1], w1, a1);
        }
    }
    acc = vadd(a0, a1);
}

}  // namespace hmfe
