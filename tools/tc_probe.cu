// Stand-alone probe of the sm_100a units behind the tensor-core mel projection (csrc/tc_ptx.cuh): checks, against
// a CPU computation, every layout assumption the product kernel makes BEFORE that kernel depends on them.
//   1. tensor-memory allocation, tcgen05.st / tcgen05.ld lane and column mapping (A operand written by 4 warps)
//   2. tcgen05.mma kind::f16 (bf16 x bf16 -> f32), A from tensor memory (M = 128), B from shared memory through a
//      K-major SWIZZLE_128B descriptor, N = 16, K = 512 as 32 instructions
//   3. the "8 stored rows" trick: stride-byte-offset such that rows 8..15 of B alias the next K atom (their D
//      columns are discarded) so that a B tile costs 8 KB instead of 16 KB
//   4. cp.async.bulk global -> shared with mbarrier complete_tx
//   5. setmaxnreg.dec / .inc across warpgroups
// Every wait is bounded; a failed wait is reported, the program never hangs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o build/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../heart_murmur_detection_b200/csrc/tc_ptx.cuh"

using namespace hmfe::tc;

struct ProbeParams {
    const uint32_t* a_words;  // [128][256] bf16 pairs (k even in the low half)
    const uint8_t* b_image;   // bytes copied verbatim to the (1024-aligned) B tile
    uint32_t b_bytes;
    uint32_t sbo, lbo, atom_stride, k_per_atom;  // descriptor parameters; k_per_atom MMAs (K = 16 each) per K atom
    uint32_t layout;                             // kSwizzle*
    uint32_t n_mma;                              // number of K = 16 instructions
    uint32_t use_bulk;                           // 1: B image through cp.async.bulk
    float* d_out;                                // [128][16]
    uint32_t* a_back;                            // [128][256]
    uint32_t* status;                            // [8]
};

__global__ void __launch_bounds__(160, 1) probe_mma(const ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[2];
    const uint32_t tile = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t* tile_ptr = smem + (tile - smem_u32(smem));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_mma = smem_u32(&s_bar[0]), bar_cp = smem_u32(&s_bar[1]);

    if (threadIdx.x == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_cp, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) p.status[0] = tmem;
    const uint32_t tmem_a = tmem, tmem_d = tmem + 256;

    // ---- 1. A rows -> tensor memory, and straight back
    if (warp < 4) {
        const int row = 32 * warp + lane;
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
        for (int c = 0; c < 256; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = p.a_words[row * 256 + c + j];
            tmem_st8(tmem_a + lane_base + c, v);
        }
        tmem_wait_st();
        for (int c = 0; c < 256; c += 8) {
            uint32_t v[8];
            tmem_ld8(tmem_a + lane_base + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) p.a_back[row * 256 + c + j] = v[j];
        }
    }
    // ---- B image -> shared memory
    if (p.use_bulk) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar_cp, p.b_bytes);
            bulk_g2s(tile, p.b_image, p.b_bytes, bar_cp);
        }
        const bool ok = mbar_wait(bar_cp, 0);
        if (!ok && lane == 0) atomicOr(&p.status[1], 1u << warp);
    } else {
        for (uint32_t i = threadIdx.x; i < p.b_bytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(tile_ptr)[i] = reinterpret_cast<const uint32_t*>(p.b_image)[i];
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- 2. the MMAs, one thread
    if (warp == 4 && lane == 0) {
        const uint32_t idesc = idesc_bf16_f32(128, 16);
        for (uint32_t k = 0; k < p.n_mma; ++k) {
            const uint32_t addr = tile + (k / p.k_per_atom) * p.atom_stride + (k % p.k_per_atom) * 32u;
            mma_ts_f16(tmem_d, tmem_a + 8 * k, smem_desc(addr, p.lbo, p.sbo, p.layout), idesc, k > 0);
        }
        mma_commit(bar_mma);
    }
    const bool ok = mbar_wait(bar_mma, 0);
    if (!ok && lane == 0) atomicOr(&p.status[2], 1u << warp);
    tc_fence_after();
    if (warp < 4) {
        const int row = 32 * warp + lane;
        uint32_t v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * warp) << 16), v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) p.d_out[row * 16 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    if (threadIdx.x == 0) p.status[3] = 0xd0e5u;
}

// ---- 6. tensor-pipe time of small-N instructions: REPS back-to-back tcgen05.mma (M = 128, K = 16, A from tensor memory,
// B from shared memory) issued by one thread; cycles from the first issue to the commit's arrival
template <int N>
__global__ void __launch_bounds__(128, 1) probe_mma_time(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t tile = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const uint32_t bar = smem_u32(&s_bar);
    for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (tile - smem_u32(smem)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    long long t_issue = 0, t_done = 0;
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16_f32(128, N);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    mma_ts_f16(tmem + 256, tmem + 8 * k, smem_desc(tile + (k >> 2) * 2048 + (k & 3) * 32, 0, 1024, kSwizzle128B), idesc, k > 0);
            }
            mma_commit(bar);
        }
        __syncwarp();
        t_issue = clock64() - t0;
        mbar_wait(bar, 0);
        t_done = clock64() - t0;
        if ((threadIdx.x & 31) == 0) {
            out[0] = t_issue;
            out[1] = t_done;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- 5. setmaxnreg across warpgroups: 4 warps give registers away, 12 warps take them
__global__ void __launch_bounds__(512, 1) probe_setmaxnreg(float* out) {
    const int warp = threadIdx.x >> 5;
    float acc[96];
    if (warp < 4) {
        setmaxnreg_dec<32>();
        out[threadIdx.x] = 1.0f;
    } else {
        setmaxnreg_inc<160>();
#pragma unroll
        for (int i = 0; i < 96; ++i) acc[i] = out[512 + ((threadIdx.x * 7 + i * 13) & 1023)];
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
#pragma unroll
            for (int i = 0; i < 96; ++i) acc[i] = fmaf(acc[i], acc[(i + 1) % 96], 0.5f);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 96; ++i) s += acc[i];
        out[threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------ host
static uint16_t f2bf(float f) {  // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

struct Layout {
    const char* name;
    uint32_t rows_stored, sbo, atom_stride;  // SWIZZLE_128B K-major, 64 bf16 per atom row
};

// byte offset of element (n, kappa) in a SWIZZLE_128B K-major tile
static uint32_t elem_off(const Layout& L, int n, int kappa) {
    const int atom = kappa / 64, chunk = (kappa % 64) / 8, e = kappa % 8;
    const uint32_t row_off = (uint32_t)(n / 8) * L.sbo + (uint32_t)(n % 8) * 128u;
    return (uint32_t)atom * L.atom_stride + row_off + (uint32_t)((chunk ^ (n % 8)) * 16 + e * 2);
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s sm_%d%d\n", prop.name, prop.major, prop.minor);

    uint32_t *d_a, *d_aback, *d_status;
    uint8_t* d_b;
    float* d_d;
    const size_t img_cap = 32768;
    CK(cudaMalloc(&d_a, 128 * 256 * 4));
    CK(cudaMalloc(&d_aback, 128 * 256 * 4));
    CK(cudaMalloc(&d_status, 32));
    CK(cudaMalloc(&d_b, img_cap));
    CK(cudaMalloc(&d_d, 128 * 16 * 4));
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)img_cap + 2048));

    const Layout layouts[3] = {
        {"full16 (16 rows stored, SBO 1024, atom 2048)", 16, 1024, 2048},
        {"alias8 (8 rows stored, SBO 1024 = next atom, atom 1024)", 8, 1024, 1024},
        {"dup8 (8 rows stored, SBO 0, atom 1024)", 8, 0, 1024},
    };
    int failures = 0;
    srand(1234);
    for (int li = 0; li < 3; ++li) {
        const Layout& L = layouts[li];
        for (int exp = 0; exp < 6; ++exp) {
            // exp 0..3: A = identity shifted to K range [128 exp, 128 exp + 128), small-integer B (exact)
            // exp 4: random data, generic copy; exp 5: random data, bulk copy
            std::vector<float> A(128 * 512, 0.0f), P(16 * 512, 0.0f);
            if (exp < 4) {
                for (int r = 0; r < 128; ++r) A[r * 512 + 128 * exp + r] = 1.0f;
                for (int n = 0; n < 16; ++n)
                    for (int k = 0; k < 512; ++k) P[n * 512 + k] = (float)((n * 37 + k * 11) % 251 - 125);
            } else {
                for (auto& v : A) v = bf2f(f2bf((float)rand() / RAND_MAX * 0.02f));
                for (auto& v : P) v = bf2f(f2bf(expf(((float)rand() / RAND_MAX - 0.5f) * 20.0f)));
            }
            std::vector<uint32_t> aw(128 * 256);
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < 256; ++c)
                    aw[r * 256 + c] = (uint32_t)f2bf(A[r * 512 + 2 * c]) | ((uint32_t)f2bf(A[r * 512 + 2 * c + 1]) << 16);
            // image: 8 K atoms (+1 atom of slack that the alias layout reads for rows 8..15 of the last atom)
            const uint32_t img_bytes = 8 * L.atom_stride + 1024;
            std::vector<uint8_t> img(img_bytes, 0);
            for (int n = 0; n < (int)L.rows_stored; ++n)
                for (int k = 0; k < 512; ++k) {
                    const uint16_t h = f2bf(P[n * 512 + k]);
                    memcpy(&img[elem_off(L, n, k)], &h, 2);
                }
            CK(cudaMemcpy(d_a, aw.data(), aw.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_b, img.data(), img_bytes, cudaMemcpyHostToDevice));
            CK(cudaMemset(d_status, 0, 32));
            CK(cudaMemset(d_d, 0xff, 128 * 16 * 4));
            ProbeParams pp{d_a, d_b, img_bytes, L.sbo, 0, L.atom_stride, 4, (uint32_t)kSwizzle128B, 32, exp == 5 ? 1u : 0u,
                           d_d, d_aback, d_status};
            probe_mma<<<1, 160, img_cap + 2048>>>(pp);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("[%s exp %d] kernel error: %s\n", L.name, exp, cudaGetErrorString(e));
                return 3;
            }
            uint32_t st[8];
            std::vector<float> D(128 * 16);
            std::vector<uint32_t> ab(128 * 256);
            CK(cudaMemcpy(st, d_status, 32, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ab.data(), d_aback, ab.size() * 4, cudaMemcpyDeviceToHost));
            int a_bad = 0;
            for (size_t i = 0; i < ab.size(); ++i) a_bad += ab[i] != aw[i];
            // expected D: rows of B that exist (stored), aliased rows per layout
            double max_err = 0, max_ref = 0;
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 16; ++n) {
                    double ref = 0;
                    bool defined = true;
                    for (int k = 0; k < 512; ++k) {
                        float b;
                        if (n < (int)L.rows_stored)
                            b = P[n * 512 + k];
                        else if (L.sbo == 0)
                            b = P[(n - 8) * 512 + k];
                        else  // alias: row n - 8 of the NEXT atom
                            b = k + 64 < 512 ? P[(n - 8) * 512 + k + 64] : 0.0f;
                        ref += (double)A[r * 512 + k] * b;
                    }
                    if (!defined) continue;
                    const double err = fabs((double)D[r * 16 + n] - ref);
                    max_err = fmax(max_err, err);
                    max_ref = fmax(max_ref, fabs(ref));
                    if (n < 8 && err > 1e-3 * fmax(1.0, fabs(ref))) ++bad;
                }
            const bool pass = a_bad == 0 && bad == 0 && st[1] == 0 && st[2] == 0 && st[3] == 0xd0e5u;
            failures += !pass;
            printf("[%s] exp %d: %s  tmem_base=0x%08x a_readback_mismatch=%d d_bad(cols<8)=%d max_err(all cols)=%.3g "
                   "max_ref=%.3g wait_fail(cp,mma)=%x,%x\n",
                   L.name, exp, pass ? "PASS" : "FAIL", st[0], a_bad, bad, max_err, max_ref, st[1], st[2]);
            if (!pass && exp < 4) {  // decode: which (n', k') did the hardware read for output (r, n)?
                int shown = 0;
                for (int r = 0; r < 128 && shown < 12; r += 9)
                    for (int n = 0; n < 16 && shown < 12; n += 5) {
                        printf("    D[%d][%d] = %g (expected P[%d][%d] = %g)\n", r, n, D[r * 16 + n], n, 128 * exp + r,
                               n < 8 ? P[n * 512 + 128 * exp + r] : NAN);
                        ++shown;
                    }
            }
        }
    }
    {
        float* d_o;
        CK(cudaMalloc(&d_o, (512 + 1024) * 4));
        CK(cudaMemset(d_o, 0, (512 + 1024) * 4));
        probe_setmaxnreg<<<2, 512>>>(d_o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("[setmaxnreg dec 32 / inc 160 over 4 warpgroups] %s\n", e == cudaSuccess ? "PASS" : cudaGetErrorString(e));
        failures += e != cudaSuccess;
    }
    {
        long long* d_t;
        CK(cudaMalloc(&d_t, 16));
        const int reps = 32;
        auto run = [&](auto kern, int n) -> int {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
            for (int it = 0; it < 2; ++it) kern<<<1, 128, 40960>>>(d_t, reps);
            CK(cudaDeviceSynchronize());
            long long t[2];
            CK(cudaMemcpy(t, d_t, 16, cudaMemcpyDeviceToHost));
            printf("[mma time] M=128 N=%3d K=16 TS: %d instructions: issue %lld cycles, complete %lld cycles = %.1f cycles / instruction\n",
                   n, reps * 32, t[0], t[1], (double)t[1] / (reps * 32));
            return 0;
        };
        if (run(probe_mma_time<16>, 16) || run(probe_mma_time<32>, 32) || run(probe_mma_time<64>, 64) ||
            run(probe_mma_time<128>, 128) || run(probe_mma_time<256>, 256))
            return 2;
    }
    printf("tc_probe: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
