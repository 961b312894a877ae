// Lane-level phases of the HeAR mel-PCEN front-end kernel: a 400-point FFT (the reference calls
// torch.fft.rfft(frames, n=400), hear/python/data_processing/audio_utils.py:112-115 with fft_length =
// frame_length = 400 from :376-384), two real frames packed as one complex transform, split 400 = 25 x 16:
//   pass 1  lane = (transform, n2), registers = n1 : 25-point DFT over n1 (5 x 5, radix-5 butterflies)
//           sample index n = 16 n1 + n2
//   twiddle W_400^(n2 k1), exchange through shared memory
//   pass 2  lane = k1 (25 of 32 lanes), registers = n2 : 16-point FFT over n2 -> bin k = k1 + 25 k2
//   separation |X_a[k]|^2, |X_b[k]|^2 for k <= 200 from Z[k] and Z[400-k] (partner lane 25 - k1 via shuffle)
// Every phase is a __host__ __device__ template taking the lane id explicitly, so csrc/host_check.cu runs the
// identical arithmetic on the CPU.
#pragma once
#include "fft_core.cuh"
#include "logmel_core.cuh"

namespace hmfe {

constexpr int kHearN = 400;       // frame length = FFT length
constexpr int kHearShift = 160;   // frame step
constexpr int kHearBins = 201;
constexpr int kHearXStride = 17;  // exchange tile row stride (complex elements), rows = k1
constexpr int kHearXTile = 25 * kHearXStride;
constexpr int kHearPRows = 256;   // power tile rows (>= 201 + banded slack)

// cos / sin of 2 pi m / 25 for the products b*c of the 5 x 5 split (b, c in 1..4) and for the radix-5 butterfly
__host__ __device__ constexpr float tw25_cos(int m) {
    switch (m) {
        case 0: return 1.0f;
        case 1: return 0.96858316112863108f;
        case 2: return 0.87630668004386358f;
        case 3: return 0.72896862742141155f;
        case 4: return 0.53582679497899655f;
        case 5: return 0.30901699437494745f;
        case 6: return 0.062790519529313527f;
        case 8: return -0.42577929156507272f;
        case 9: return -0.63742398974868975f;
        case 10: return -0.80901699437494734f;
        case 12: return -0.99211470131447776f;
        default: return -0.63742398974868952f;  // 16
    }
}
__host__ __device__ constexpr float tw25_sin(int m) {
    switch (m) {
        case 0: return 0.0f;
        case 1: return 0.24868988716485479f;
        case 2: return 0.48175367410171532f;
        case 3: return 0.68454710592868862f;
        case 4: return 0.84432792550201508f;
        case 5: return 0.95105651629515353f;
        case 6: return 0.99802672842827156f;
        case 8: return 0.90482705246601947f;
        case 9: return 0.77051324277578925f;
        case 10: return 0.58778525229247325f;
        case 12: return 0.12533323356430454f;
        default: return -0.77051324277578936f;  // 16
    }
}

// 5-point DFT, forward (W = exp(-2 pi i / 5)), in place on five (re, im) pairs.
HMFE_HD void dft5(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3, float& r4,
                  float& i4) {
    constexpr float c1 = tw25_cos(5), c2 = tw25_cos(10), s1 = tw25_sin(5), s2 = tw25_sin(10);
    const float t1r = r1 + r4, t1i = i1 + i4, t2r = r2 + r3, t2i = i2 + i3;
    const float t3r = r1 - r4, t3i = i1 - i4, t4r = r2 - r3, t4i = i2 - i3;
    const float m1r = fmaf(c2, t2r, fmaf(c1, t1r, r0)), m1i = fmaf(c2, t2i, fmaf(c1, t1i, i0));
    const float m2r = fmaf(c1, t2r, fmaf(c2, t1r, r0)), m2i = fmaf(c1, t2i, fmaf(c2, t1i, i0));
    const float q1r = fmaf(s2, t4r, s1 * t3r), q1i = fmaf(s2, t4i, s1 * t3i);    // s1 t3 + s2 t4
    const float q2r = fmaf(-s1, t4r, s2 * t3r), q2i = fmaf(-s1, t4i, s2 * t3i);  // s2 t3 - s1 t4
    r0 = r0 + t1r + t2r;
    i0 = i0 + t1i + t2i;
    // X1 = m1 - i q1, X4 = m1 + i q1, X2 = m2 - i q2, X3 = m2 + i q2   (-i (a + ib) = b - ia)
    r1 = m1r + q1i;
    i1 = m1i - q1r;
    r4 = m1r - q1i;
    i4 = m1i + q1r;
    r2 = m2r + q2i;
    i2 = m2i - q2r;
    r3 = m2r - q2i;
    i3 = m2i + q2r;
}

template <int B, int C>
HMFE_HD void tw25_apply(float& r, float& i) {
    constexpr float c = tw25_cos(B * C), s = tw25_sin(B * C);  // W_25^(bc) = (c, -s)
    const float nr = fmaf(i, s, r * c);
    const float ni = fmaf(-r, s, i * c);
    r = nr;
    i = ni;
}

// register position p of the 25-point DFT output holds frequency index k1 = p / 5 + 5 (p % 5)
__host__ __device__ constexpr int hear_k1_of(int p) { return p / 5 + 5 * (p % 5); }

// 25-point DFT in registers: input in natural order (x[n1] at position n1), output digit-reversed (hear_k1_of).
HMFE_HD void dft25(float (&re)[25], float (&im)[25]) {
#define HMFE_D5(a, b, c, d, e) dft5(re[a], im[a], re[b], im[b], re[c], im[c], re[d], im[d], re[e], im[e])
    // step 1: for each b, DFT over a of x[5a + b] -> position 5c + b
    HMFE_D5(0, 5, 10, 15, 20);
    HMFE_D5(1, 6, 11, 16, 21);
    HMFE_D5(2, 7, 12, 17, 22);
    HMFE_D5(3, 8, 13, 18, 23);
    HMFE_D5(4, 9, 14, 19, 24);
    // step 2: position 5c + b *= W_25^(bc)
#define HMFE_T25(b, c) tw25_apply<b, c>(re[5 * c + b], im[5 * c + b])
    HMFE_T25(1, 1); HMFE_T25(2, 1); HMFE_T25(3, 1); HMFE_T25(4, 1);
    HMFE_T25(1, 2); HMFE_T25(2, 2); HMFE_T25(3, 2); HMFE_T25(4, 2);
    HMFE_T25(1, 3); HMFE_T25(2, 3); HMFE_T25(3, 3); HMFE_T25(4, 3);
    HMFE_T25(1, 4); HMFE_T25(2, 4); HMFE_T25(3, 4); HMFE_T25(4, 4);
#undef HMFE_T25
    // step 3: for each c, DFT over b of positions 5c + b -> position 5c + d holds k1 = c + 5d
    HMFE_D5(0, 1, 2, 3, 4);
    HMFE_D5(5, 6, 7, 8, 9);
    HMFE_D5(10, 11, 12, 13, 14);
    HMFE_D5(15, 16, 17, 18, 19);
    HMFE_D5(20, 21, 22, 23, 24);
#undef HMFE_D5
}

// ---- pass 1 for lane = 16 * tr + n2.  fetch(tr, second, n) = sample n (0..399) of the transform's frame a / b.
// plane[k1 * 16 + n2] = (cos, -sin)(2 pi n2 k1 / 400); tile = [2 transforms][25 rows k1][kHearXStride] complex.
template <typename Fetch>
HMFE_HD void hear_pass1(int lane, const float* __restrict__ win, const float2* __restrict__ plane, Fetch fetch,
                        float2* __restrict__ tile) {
    const int tr = lane >> 4, n2 = lane & 15;
    float re[25], im[25];
#pragma unroll
    for (int n1 = 0; n1 < 25; ++n1) {
        const int n = 16 * n1 + n2;
        const float w = win[n];
        re[n1] = fetch(tr, false, n) * w;
        im[n1] = fetch(tr, true, n) * w;
    }
    dft25(re, im);
    float2* t = tile + tr * kHearXTile;
#pragma unroll
    for (int p = 0; p < 25; ++p) {
        const int k1 = hear_k1_of(p);
        float r = re[p], i = im[p];
        if (k1 != 0) {
            const float2 w = plane[k1 * 16 + n2];
            const float nr = fmaf(-i, w.y, r * w.x);  // (r + i i)(w.x + i w.y)
            const float ni = fmaf(r, w.y, i * w.x);
            r = nr;
            i = ni;
        }
        t[k1 * kHearXStride + n2] = make_float2(r, i);
    }
}

// ---- pass 2 for lane = k1 < 25 on transform tr: returns Z[k1 + 25 k2] in z[k2]
HMFE_HD void hear_pass2(int lane, const float2* __restrict__ tile, int tr, float (&zr)[16], float (&zi)[16]) {
    const float2* t = tile + tr * kHearXTile + lane * kHearXStride;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
        const float2 e = t[n2];
        zr[brev(n2, 4)] = e.x;
        zi[brev(n2, 4)] = e.y;
    }
    fft_dit<16, float>(zr, zi);
}

// register of lane (25 - k1) % 25 that holds Z[400 - k] for k = k1 + 25 k2: what that lane hands to its partner
__host__ __device__ constexpr int hear_give_reg(bool lane_is_zero, int k2) { return lane_is_zero ? ((16 - k2) & 15) : 15 - k2; }

// float32 operations that must not be contracted into FMAs (the reference rounds every torch op separately)
#if defined(__CUDA_ARCH__)
HMFE_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
HMFE_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
HMFE_HD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
HMFE_HD float div_rn(float a, float b) { return __fdiv_rn(a, b); }
#else
HMFE_HD float mul_rn(float a, float b) { return a * b; }
HMFE_HD float add_rn(float a, float b) { return a + b; }
HMFE_HD float sub_rn(float a, float b) { return a - b; }
HMFE_HD float div_rn(float a, float b) { return a / b; }
#endif

// x^y for x > 0.  Device: exp2(y * log2 x) with the special-function log2 (relative error ~1e-6 for the PCEN operands,
// i.e. <= 1e-5 absolute on outputs bounded by ~6; budget 1e-4); host emulation: powf.
HMFE_HD float pow_pos(float x, float y) {
#if defined(__CUDA_ARCH__)
    return y == 0.5f ? sqrtf(x) : exp2f(y * __log2f(x));
#else
    return powf(x, y);
#endif
}

// scaling of audio_utils.py:361-365: x -= min; x /= (max + 1e-8) [max taken after the shift]; x = 2 x - 1
struct HearScale {
    float mn, den, inv;
};
HMFE_HD HearScale hear_make_scale(float mn, float mx) {
    const float den = add_rn(sub_rn(mx, mn), 1e-8f);
    return HearScale{mn, den, div_rn(1.0f, den)};
}
// The quotient (x - mn) / den is formed as q0 = t * (1 / den) plus one residual correction q0 + (t - q0 den) (1 / den):
// with a correctly rounded reciprocal this is the correctly rounded quotient (Markstein), three FMA-pipe instructions
// instead of the division sequence (0 <= t <= den, so no overflow; a constant batch, den = 1e-8, scales to -1 like the
// reference's 0 / 1e-8).
HMFE_HD float hear_scale(const HearScale& s, float x) {
    const float t = sub_rn(x, s.mn);
    const float q0 = mul_rn(t, s.inv);
    const float q = fmaf(fmaf(-q0, s.den, t), s.inv, q0);
    return sub_rn(mul_rn(q, 2.0f), 1.0f);
}

// ---- PCEN + row interpolation of one mel channel (audio_utils.py:121-246 and :386-445).
struct PcenParams {
    float alpha, c_in, c_state, delta, inv_root, floor, delta_root;
};

// load(t) = mel power of frame t, store(i, v) = output row i.  The smoother starts at the first frame; each step is
// the sum of two separately rounded products (the reference multiplies by two diagonal matrices and adds).  Output
// row i is torch's bilinear interpolation with align_corners=False: source position scale * (i + 0.5) - 0.5
// clamped at 0, rows floor and floor + 1 (clamped at T - 1).
template <typename Load, typename Store>
HMFE_HD void hear_pcen_column(const PcenParams& pp, int T, int out_rows, Load load, Store store) {
    const float scale = (float)T / (float)out_rows;
    float ema = 0.0f, p_prev = 0.0f, p_cur = 0.0f;
    int t_cur = -1;
    float x_next = load(0);
    for (int i = 0; i < out_rows; ++i) {
        float real = scale * ((float)i + 0.5f) - 0.5f;
        real = real < 0.0f ? 0.0f : real;
        int r0 = (int)floorf(real);
        r0 = r0 < T - 1 ? r0 : T - 1;
        const int r1 = r0 + 1 < T - 1 ? r0 + 1 : T - 1;
        float lam = real - (float)r0;
        lam = lam < 0.0f ? 0.0f : (lam > 1.0f ? 1.0f : lam);
        while (t_cur < r1) {
            ++t_cur;
            const float xv = x_next;
            if (t_cur + 1 < T) x_next = load(t_cur + 1);
            ema = t_cur == 0 ? xv : add_rn(mul_rn(xv, pp.c_in), mul_rn(ema, pp.c_state));
            p_prev = p_cur;
            const float g = pow_pos(add_rn(pp.floor, ema), pp.alpha);
            p_cur = sub_rn(pow_pos(add_rn(div_rn(xv, g), pp.delta), pp.inv_root), pp.delta_root);
        }
        const float a = r0 == t_cur ? p_cur : p_prev;
        store(i, add_rn(mul_rn(1.0f - lam, a), mul_rn(lam, p_cur)));
    }
}

}  // namespace hmfe
