// Shared helpers for the hmfe (heart-murmur front-end) sm_100a kernels.
//
// The arithmetic cores (FFT butterflies, frame separation, mel accumulation) are
// written once as templates over a value type V:
//   V = float  : one transform per lane
//   V = f32x2  : two independent transforms per lane, executed with Blackwell's
//                packed FP32 instructions (PTX add/mul/fma.rn.f32x2 -> SASS
//                FADD2/FMUL2/FFMA2), which halve the issue slots of the FP32 work.
// The same templates compile for the host (plain C++), which is how
// csrc/host_check.cu validates index math and precision without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define HMFE_HD __host__ __device__ __forceinline__
#define HMFE_D __device__ __forceinline__

namespace hmfe {

struct __align__(8) f32x2 {
    float x, y;
};

// ---------------------------------------------------------------- scalar ops
HMFE_HD float vset(float, float a) { return a; }
HMFE_HD float vadd(float a, float b) { return a + b; }
HMFE_HD float vsub(float a, float b) { return a - b; }
HMFE_HD float vmul(float a, float b) { return a * b; }
HMFE_HD float vmuls(float a, float s) { return a * s; }
HMFE_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }     // a*b + c
HMFE_HD float vfmas(float a, float s, float c) { return fmaf(a, s, c); }    // a*s + c   (s scalar)
HMFE_HD float vfnmas(float a, float s, float c) { return fmaf(-a, s, c); }  // c - a*s
HMFE_HD float vfmsub2(float a, float c) { return fmaf(2.0f, a, -c); }       // 2a - c

// ---------------------------------------------------------------- packed ops
#if defined(__CUDA_ARCH__)
HMFE_D unsigned long long& as_u64(f32x2& a) { return *reinterpret_cast<unsigned long long*>(&a); }
HMFE_D f32x2 vadd(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return d;
}
HMFE_D f32x2 vsub(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return d;
}
HMFE_D f32x2 vmul(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return d;
}
HMFE_D f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
    return d;
}
#else
HMFE_HD f32x2 vadd(f32x2 a, f32x2 b) { return f32x2{a.x + b.x, a.y + b.y}; }
HMFE_HD f32x2 vsub(f32x2 a, f32x2 b) { return f32x2{a.x - b.x, a.y - b.y}; }
HMFE_HD f32x2 vmul(f32x2 a, f32x2 b) { return f32x2{a.x * b.x, a.y * b.y}; }
HMFE_HD f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) { return f32x2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif
HMFE_HD f32x2 vneg(f32x2 a) { return f32x2{-a.x, -a.y}; }
HMFE_HD float vneg(float a) { return -a; }
HMFE_HD f32x2 vbcast(float s) { return f32x2{s, s}; }
HMFE_HD f32x2 vmuls(f32x2 a, float s) { return vmul(a, vbcast(s)); }
HMFE_HD f32x2 vfmas(f32x2 a, float s, f32x2 c) { return vfma(a, vbcast(s), c); }
HMFE_HD f32x2 vfnmas(f32x2 a, float s, f32x2 c) { return vfma(a, vbcast(-s), c); }
HMFE_HD f32x2 vfmsub2(f32x2 a, f32x2 c) { return vfma(a, vbcast(2.0f), vneg(c)); }

template <typename V>
struct lanes_of;
template <>
struct lanes_of<float> {
    static constexpr int value = 1;
};
template <>
struct lanes_of<f32x2> {
    static constexpr int value = 2;
};

HMFE_HD float vget(float a, int) { return a; }
HMFE_HD float vget(f32x2 a, int i) { return i ? a.y : a.x; }
HMFE_HD void vput(float& a, int, float v) { a = v; }
HMFE_HD void vput(f32x2& a, int i, float v) {
    if (i)
        a.y = v;
    else
        a.x = v;
}

}  // namespace hmfe
