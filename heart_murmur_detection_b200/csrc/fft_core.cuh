// In-register radix-2 decimation-in-time FFT (N <= 32), fully unrolled, twiddles as
// immediates.  Butterflies use the FMA-fused form
//     a' = a + w*b   (2 dependent FMAs per component)
//     b' = 2a - a'   (1 FMA per component)
// i.e. 6 FP instructions per non-trivial radix-2 butterfly, 4 for w in {1, -i}.
// Input is expected in bit-reversed register order, output is in natural order.
// Works for V = float and V = f32x2 (two independent transforms, packed FP32).
#pragma once
#include "hmfe_common.cuh"

namespace hmfe {

// cos / sin of 2*pi*t/32, t = 0..15 (double-rounded-once literals)
__host__ __device__ constexpr float tw32_cos(int t) {
    switch (t) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654757f;
        case 5: return 0.55557023301960229f;
        case 6: return 0.38268343236508984f;
        case 7: return 0.19509032201612833f;
        case 8: return 0.0f;
        case 9: return -0.19509032201612833f;
        case 10: return -0.38268343236508984f;
        case 11: return -0.55557023301960229f;
        case 12: return -0.70710678118654757f;
        case 13: return -0.83146961230254524f;
        case 14: return -0.92387953251128674f;
        default: return -0.98078528040323043f;
    }
}
__host__ __device__ constexpr float tw32_sin(int t) { return t <= 8 ? tw32_cos(8 - t) : tw32_cos(t - 8); }

__host__ __device__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// One DIT butterfly with twiddle W = exp(-2*pi*i*T/32).
template <int T, typename V>
HMFE_HD void butterfly(V& ar, V& ai, V& br, V& bi) {
    if (T == 0) {
        const V xr = vadd(ar, br), xi = vadd(ai, bi);
        br = vsub(ar, br);
        bi = vsub(ai, bi);
        ar = xr;
        ai = xi;
    } else if (T == 8) {  // w = -i : w*b = (bi, -br)
        const V xr = vadd(ar, bi), xi = vsub(ai, br);
        const V yr = vsub(ar, bi), yi = vadd(ai, br);
        ar = xr;
        ai = xi;
        br = yr;
        bi = yi;
    } else {
        constexpr float c = tw32_cos(T), s = tw32_sin(T);  // w = (c, -s)
        const V xr = vfmas(bi, s, vfmas(br, c, ar));        // ar + c*br + s*bi
        const V xi = vfnmas(br, s, vfmas(bi, c, ai));       // ai + c*bi - s*br
        br = vfmsub2(ar, xr);
        bi = vfmsub2(ai, xi);
        ar = xr;
        ai = xi;
    }
}

template <int N, int M, int K, int J, typename V>
struct bf_loop {
    static HMFE_HD void run(V (&re)[N], V (&im)[N]) {
        constexpr int H = M / 2;
        butterfly<J*(32 / M), V>(re[K + J], im[K + J], re[K + J + H], im[K + J + H]);
        if constexpr (J + 1 < H)
            bf_loop<N, M, K, J + 1, V>::run(re, im);
        else if constexpr (K + M < N)
            bf_loop<N, M, K + M, 0, V>::run(re, im);
    }
};

template <int N, int M, typename V>
HMFE_HD void fft_stage(V (&re)[N], V (&im)[N]) {
    bf_loop<N, M, 0, 0, V>::run(re, im);
}

template <int N, typename V>
HMFE_HD void fft_dit(V (&re)[N], V (&im)[N]) {
    static_assert(N == 2 || N == 4 || N == 8 || N == 16 || N == 32, "N must be a power of two <= 32");
    if constexpr (N >= 2) fft_stage<N, 2, V>(re, im);
    if constexpr (N >= 4) fft_stage<N, 4, V>(re, im);
    if constexpr (N >= 8) fft_stage<N, 8, V>(re, im);
    if constexpr (N >= 16) fft_stage<N, 16, V>(re, im);
    if constexpr (N >= 32) fft_stage<N, 32, V>(re, im);
}

}  // namespace hmfe
