// Host-side plumbing shared by the C-ABI translation units: error reporting and a
// small ring of pinned/device descriptor buffers for per-call ragged-batch metadata.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/hmfe.h"

namespace hmfe {

void set_error(const char* fmt, ...);

#define HMFE_CHECK_CUDA(expr)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::hmfe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                              __LINE__);                                                        \
            return HMFE_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define HMFE_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::hmfe::set_error(__VA_ARGS__);  \
            return HMFE_ERR_INVALID;         \
        }                                    \
    } while (0)

// Per-call metadata (clip offsets, work prefixes, gather descriptors) is built on the host
// in pinned memory and copied to the device on the caller's stream.  A ring of slots keeps
// the call asynchronous: a slot is reused only after the event recorded behind its last
// consumer has completed.
class DescRing {
   public:
    static constexpr int kSlots = 4;
    ~DescRing();
    // returns slot index (>= 0) or a negative error code
    int acquire(size_t bytes, void** host, void** dev);
    int upload(int slot, size_t bytes, cudaStream_t s);  // H2D async of the first `bytes`
    int release(int slot, cudaStream_t s);               // record the reuse event
    int reserve(size_t bytes);  // allocate every slot up front: later calls needing <= bytes allocate nothing
    void forbid_growth() { fixed_ = true; }

   private:
    void* h_[kSlots] = {};
    void* d_[kSlots] = {};
    size_t cap_[kSlots] = {};
    cudaEvent_t ev_[kSlots] = {};
    bool pending_[kSlots] = {};
    int next_ = 0;
    bool fixed_ = false;
};

int device_sm_count();

}  // namespace hmfe
