// hmfe_ctx: per-caller scratch (descriptor ring + growable device scratch) for the
// plan-less stages (trim, gather, IIR, spectrogram ops).
#pragma once
#include <limits.h>
#include <string.h>

#include <vector>

#include "api_common.h"

enum HmfeKernelId {
    HMFE_K_IIR_ZERO_STATE = 0,
    HMFE_K_IIR_CARRY = 1,
    HMFE_K_IIR_FINAL = 2,
    HMFE_K_TRIM_POWER = 3,
    HMFE_K_TRIM_INDEX = 4,
    HMFE_K_GATHER = 5,
    HMFE_K_SPEC_MEAN = 6,
    HMFE_K_SPEC_CROP = 7,
    HMFE_K_IIR_OVERLAP = 8,
    HMFE_K_COUNT = 9
};

struct hmfe_ctx {
    hmfe::DescRing ring;
    void* scratch = nullptr;
    size_t scratch_cap = 0;
    bool scratch_external = false;  // caller-provided workspace (hmfe_ctx_set_workspace): never (re)allocated here
    int sm_count = 148;
    int last_launches = 0;
    // IIR algorithm choice (hmfe_ctx_set_iir_algo) and what the last call used
    int iir_algo = HMFE_IIR_ALGO_AUTO;
    int iir_rows = HMFE_IIR_ROWS_AUTO;
    int iir_last_algo = 0, iir_last_C = 0, iir_last_W = 0, iir_last_rows = 0;
    // decay length of the last filter seen (depends on the coefficients only)
    int iir_cache_S = 0, iir_cache_W = 0;
    double iir_cache_sos[6 * 8] = {};
    // measurement hook (bench.py roofline): events around every kernel, on the launch stream
    bool profile = false;
    struct ProfRec {
        int id;
        cudaEvent_t a, b;
    };
    std::vector<ProfRec> prof;

    int prof_begin(int id, cudaStream_t st) {
        if (!profile) return HMFE_OK;
        ProfRec r{id, nullptr, nullptr};
        HMFE_CHECK_CUDA(cudaEventCreate(&r.a));
        HMFE_CHECK_CUDA(cudaEventCreate(&r.b));
        HMFE_CHECK_CUDA(cudaEventRecord(r.a, st));
        prof.push_back(r);
        return HMFE_OK;
    }
    int prof_end(cudaStream_t st) {
        if (!profile || prof.empty()) return HMFE_OK;
        HMFE_CHECK_CUDA(cudaEventRecord(prof.back().b, st));
        return HMFE_OK;
    }

    // grows the device scratch; growing frees the old block (cudaFree synchronises the device,
    // so kernels still using it have finished)
    int reserve_scratch(size_t bytes) {
        if (bytes <= scratch_cap) return HMFE_OK;
        if (scratch_external) {
            hmfe::set_error("caller-provided workspace of %zu bytes is too small: this call needs %zu (query hmfe_*_workspace_bytes)",
                            scratch_cap, bytes);
            return HMFE_ERR_INVALID;
        }
        size_t cap = scratch_cap ? scratch_cap : (size_t)1 << 20;
        while (cap < bytes) cap *= 2;
        if (scratch) HMFE_CHECK_CUDA(cudaFree(scratch));
        scratch = nullptr;
        scratch_cap = 0;
        HMFE_CHECK_CUDA(cudaMalloc(&scratch, cap));
        scratch_cap = cap;
        return HMFE_OK;
    }
};
