#include <stdarg.h>
#include <stdio.h>

#include "api_common.h"

namespace hmfe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

DescRing::~DescRing() {
    for (int i = 0; i < kSlots; ++i) {
        if (ev_[i]) cudaEventDestroy(ev_[i]);
        if (h_[i]) cudaFreeHost(h_[i]);
        if (d_[i]) cudaFree(d_[i]);
    }
}

int DescRing::acquire(size_t bytes, void** host, void** dev) {
    const int s = next_;
    next_ = (next_ + 1) % kSlots;
    if (!ev_[s]) HMFE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_[s], cudaEventDisableTiming));
    if (pending_[s]) {
        HMFE_CHECK_CUDA(cudaEventSynchronize(ev_[s]));
        pending_[s] = false;
    }
    if (bytes > cap_[s]) {
        if (fixed_) {
            set_error("descriptor staging of %zu bytes was reserved, this call needs %zu (hmfe_ctx_reserve)", cap_[s], bytes);
            return HMFE_ERR_INVALID;
        }
        size_t cap = cap_[s] ? cap_[s] : 4096;
        while (cap < bytes) cap *= 2;
        if (h_[s]) HMFE_CHECK_CUDA(cudaFreeHost(h_[s]));
        if (d_[s]) HMFE_CHECK_CUDA(cudaFree(d_[s]));
        h_[s] = d_[s] = nullptr;
        cap_[s] = 0;
        HMFE_CHECK_CUDA(cudaMallocHost(&h_[s], cap));
        HMFE_CHECK_CUDA(cudaMalloc(&d_[s], cap));
        cap_[s] = cap;
    }
    *host = h_[s];
    *dev = d_[s];
    return s;
}

int DescRing::reserve(size_t bytes) {
    for (int s = 0; s < kSlots; ++s) {
        if (!ev_[s]) HMFE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_[s], cudaEventDisableTiming));
        if (bytes <= cap_[s]) continue;
        if (pending_[s]) {
            HMFE_CHECK_CUDA(cudaEventSynchronize(ev_[s]));
            pending_[s] = false;
        }
        if (h_[s]) HMFE_CHECK_CUDA(cudaFreeHost(h_[s]));
        if (d_[s]) HMFE_CHECK_CUDA(cudaFree(d_[s]));
        h_[s] = d_[s] = nullptr;
        cap_[s] = 0;
        HMFE_CHECK_CUDA(cudaMallocHost(&h_[s], bytes));
        HMFE_CHECK_CUDA(cudaMalloc(&d_[s], bytes));
        cap_[s] = bytes;
    }
    return HMFE_OK;
}

int DescRing::upload(int slot, size_t bytes, cudaStream_t s) {
    if (bytes) HMFE_CHECK_CUDA(cudaMemcpyAsync(d_[slot], h_[slot], bytes, cudaMemcpyHostToDevice, s));
    return HMFE_OK;
}

int DescRing::release(int slot, cudaStream_t s) {
    HMFE_CHECK_CUDA(cudaEventRecord(ev_[slot], s));
    pending_[slot] = true;
    return HMFE_OK;
}

int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace hmfe

extern "C" {

int hmfe_version(void) { return 100; }

const char* hmfe_last_error(void) { return hmfe::g_err; }

int hmfe_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    HMFE_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    HMFE_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return HMFE_OK;
}

}  // extern "C"
