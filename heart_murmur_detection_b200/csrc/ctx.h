// hmfe_ctx: per-caller scratch (descriptor ring + growable device scratch) for the
// plan-less stages (trim, gather, IIR, spectrogram ops).
#pragma once
#include <limits.h>
#include <string.h>

#include "api_common.h"

struct hmfe_ctx {
    hmfe::DescRing ring;
    void* scratch = nullptr;
    size_t scratch_cap = 0;
    int sm_count = 148;
    int last_launches = 0;

    // grows the device scratch; growing frees the old block (cudaFree synchronises the device,
    // so kernels still using it have finished)
    int reserve_scratch(size_t bytes) {
        if (bytes <= scratch_cap) return HMFE_OK;
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
