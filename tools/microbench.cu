// Micro-benchmarks that size the FP32 roofline of the FFT kernels on B200:
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue/throughput, and warp shuffles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int ILP>
__global__ void k_ffma(float* out, int iters, float s) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], s, 0.5f);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float s) {
    unsigned long long acc[ILP];
    float2 sv = make_float2(s, s), cv = make_float2(0.5f, 0.25f);
    unsigned long long S = *reinterpret_cast<unsigned long long*>(&sv), Cc = *reinterpret_cast<unsigned long long*>(&cv);
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 v = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
        acc[i] = *reinterpret_cast<unsigned long long*>(&v);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma2(acc[i], S, Cc);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 v = *reinterpret_cast<float2*>(&acc[i]);
        r += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// packed FMA whose three operands are all different register pairs (the FFT butterflies look like
// this; the kernels above reuse two constant operands, which the operand-reuse cache serves)
template <int ILP>
__global__ void k_ffma2_distinct(float* out, int iters) {
    unsigned long long a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 va = make_float2(1.0f + threadIdx.x * 1e-6f + i * 1e-3f, 1.0f - i * 1e-3f);
        float2 vb = make_float2(0.999f + i * 1e-4f, 1.001f - i * 1e-4f);
        float2 vc = make_float2(i * 0.5f, threadIdx.x * 0.25f);
        a[i] = *reinterpret_cast<unsigned long long*>(&va);
        b[i] = *reinterpret_cast<unsigned long long*>(&vb);
        c[i] = *reinterpret_cast<unsigned long long*>(&vc);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma2(a[i], b[(i + 1) % ILP], c[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma2(c[i], b[i], a[(i + 3) % ILP]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 v = *reinterpret_cast<float2*>(&a[i]);
        float2 w = *reinterpret_cast<float2*>(&c[i]);
        r += v.x + v.y + w.x + w.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_ffma_distinct(float* out, int iters) {
    float a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = 1.0f + threadIdx.x * 1e-6f + i * 1e-3f;
        b[i] = 0.999f + i * 1e-4f;
        c[i] = i * 0.5f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fmaf(a[i], b[(i + 1) % ILP], c[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fmaf(c[i], b[i], a[(i + 3) % ILP]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// mixed: FFMA2 interleaved with shuffles (issue-slot sharing)
template <int ILP>
__global__ void k_shfl(float* out, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __shfl_xor_sync(0xffffffffu, acc[i], 1 + (i & 15));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], s, 0.5);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    constexpr int ILP = 8;
    const double n_thread_ops = (double)blocks * threads * iters * ILP;
    float ms = time_ms([&] { k_ffma<ILP><<<blocks, threads>>>(out, iters, 0.999f); });
    printf(", \"ffma_tflops\": %.2f, \"ffma_warp_instr_per_ns\": %.2f", 2 * n_thread_ops / ms * 1e-9, n_thread_ops / 32 / ms * 1e-6);
    ms = time_ms([&] { k_ffma2<ILP><<<blocks, threads>>>(out, iters, 0.999f); });
    printf(", \"ffma2_tflops\": %.2f, \"ffma2_warp_instr_per_ns\": %.2f", 4 * n_thread_ops / ms * 1e-9, n_thread_ops / 32 / ms * 1e-6);
    for (int w : {1, 2, 3, 4}) {  // warps per scheduler: 128 * w threads per SM (one CTA per SM)
        const double ops = (double)sms * 128 * w * iters * 16 * 2;  // 2 x ILP(16) instructions per iteration
        ms = time_ms([&] { k_ffma2_distinct<16><<<sms, 128 * w>>>(out, iters); });
        printf(", \"ffma2_distinct_w%d_tflops\": %.2f", w, 4 * ops / ms * 1e-9);
        ms = time_ms([&] { k_ffma_distinct<16><<<sms, 128 * w>>>(out, iters); });
        printf(", \"ffma_distinct_w%d_tflops\": %.2f", w, 2 * ops / ms * 1e-9);
    }
    ms = time_ms([&] { k_shfl<ILP><<<blocks, threads>>>(out, iters / 4); });
    printf(", \"shfl_warp_instr_per_ns\": %.2f", n_thread_ops / 4 / 32 / ms * 1e-6);
    double* dout;
    cudaMalloc(&dout, sizeof(double) * blocks * threads);
    ms = time_ms([&] { k_dfma<ILP><<<blocks, threads>>>(dout, iters / 8, 0.999); });
    printf(", \"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm_at_1965\": %.1f", 2 * n_thread_ops / 8 / ms * 1e-9,
           n_thread_ops / 8 / (ms * 1e-3) / sms / 1.965e9);
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
