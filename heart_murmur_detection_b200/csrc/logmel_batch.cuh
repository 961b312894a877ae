// Batch descriptors, work-item lookup and the plan object shared by the log-mel kernels
// (logmel.cu: FP32 mel projection from shared memory; logmel_tc.cu: mel projection on the tensor cores).
#pragma once
#include <stdint.h>

#include <vector>

#include "api_common.h"
#include "logmel_core.cuh"

namespace hmfe {

constexpr int kMaxSlots = 8;
constexpr int kTileElems = 32 * kXStride;  // 1056 >= kBinsPad

struct MelMeta {
    int n_slots, total_trip, n_mels;
    int trip[kMaxSlots], wbase[kMaxSlots];
};

struct LogmelBatch {
    const float* wav;
    const float* wav_alt;        // second sample buffer: clips with a negative start s live at wav_alt[-s - 1]
    float* out;
    const int64_t* clip_start;   // [n_clips] ragged only: first sample of each clip in wav
    const int64_t* clip_len;     // [n_clips] ragged only
    const int64_t* frame_off;    // [n_clips+1] ragged only
    const int64_t* item_prefix;  // [n_clips+1] ragged only
    unsigned* stats;             // [n_clips][2] : max bits, min bits of the clip's mel power
    unsigned long long* queue;   // next unclaimed work item (dynamic distribution over the warps)
    int64_t n_clips, n_items;
    int64_t n_frames_total;      // host-planned batches: frames of the whole batch (logmel_generic.cu)
    const int64_t* n_items_dev;  // != NULL: the item count is item_prefix[n_clips] in device memory (device-planned batches)
    int uniform_n, uniform_T, uniform_items;  // > 0 when every clip has the same length
    int hop;
    int stagger_ns;  // start-up delay per warp index (HMFE_LOGMEL_STAGGER_NS overrides the default)
    uint32_t* status;  // tensor-core variant: protocol-error word (0 = fine)
};

struct LogmelTables {
    const float* win;   // 1024, 0.5 * Hann
    const float2* tw;   // [32][32]
    const float* melw;  // [total_trip][32]
    const int* start;   // [n_slots][32]
    const int* row;     // [n_slots][32]
};

struct GenericTables {  // logmel_generic.cu (n_fft != 1024)
    const float* win;   // [n_fft] periodic Hann
    const float2* tw;   // [n_fft / 2] exp(-2 pi i k / n_fft)
    const float* mel;   // [n_mels][n_fft / 2 + 1]
    const int* lo;      // [n_mels] first / one past the last non-zero bin of the row
    const int* hi;
};

HMFE_D float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
HMFE_D f32x2 shfl(f32x2 v, int src) {
    return f32x2{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

HMFE_D int64_t item_count(const LogmelBatch& b) { return b.n_items_dev ? *b.n_items_dev : b.n_items; }

// Work item descriptor: 2*NV consecutive frames of one clip.
struct ItemCtx {
    const float* x;  // clip samples
    float* o;        // clip output rows
    int64_t clip;
    int nsamp, T, f0;
    bool valid;
};

template <int FR>
HMFE_D ItemCtx locate_item(const LogmelBatch& b, int n_mels, int64_t item, int64_t it_end, int64_t& clip) {
    ItemCtx c;
    c.valid = item < it_end;
    if (!c.valid) {
        c.x = b.wav;
        c.o = b.out;
        c.clip = 0;
        c.nsamp = c.T = c.f0 = 0;
        return c;
    }
    int64_t q;
    if (b.uniform_items > 0) {
        clip = item / b.uniform_items;
        q = item - clip * b.uniform_items;
        c.nsamp = b.uniform_n;
        c.T = b.uniform_T;
        c.x = b.wav + clip * (int64_t)c.nsamp;
        c.o = b.out + clip * (int64_t)c.T * n_mels;
    } else {
        if (clip < 0 || item >= b.item_prefix[min(clip + 4, b.n_clips)]) {
            // first item of this warp, or a jump to a far block: binary search, largest c with prefix[c] <= item
            int64_t lo = max(clip, (int64_t)0), hi = b.n_clips;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (b.item_prefix[mid] <= item)
                    lo = mid;
                else
                    hi = mid;
            }
            clip = lo;
        }
        while (item >= b.item_prefix[clip + 1]) ++clip;
        q = item - b.item_prefix[clip];
        c.nsamp = (int)b.clip_len[clip];
        const int64_t f0g = b.frame_off[clip];
        c.T = (int)(b.frame_off[clip + 1] - f0g);
        const int64_t s0 = b.clip_start[clip];
        c.x = s0 >= 0 ? b.wav + s0 : b.wav_alt + (-s0 - 1);
        c.o = b.out + f0g * n_mels;
    }
    c.clip = clip;
    c.f0 = (int)q * FR;
    return c;
}

}  // namespace hmfe

struct hmfe_logmel_plan {
    int sample_rate, n_fft, hop, n_mels, n_bins, variant;
    int pad_mode = HMFE_PAD_CONSTANT;
    double f_min, f_max;
    std::vector<float> mel_dense;
    hmfe::MelMeta meta;
    float *d_win = nullptr, *d_melw = nullptr;
    float2* d_tw = nullptr;
    int *d_start = nullptr, *d_row = nullptr;
    // n_fft != 1024 (logmel_generic.cu): dense mel basis and the non-zero range of its rows; d_win / d_tw hold the
    // full Hann window and the n_fft / 2 twiddles in that case
    float* d_gen_mel = nullptr;
    int *d_gen_lo = nullptr, *d_gen_hi = nullptr;
    size_t table_smem = 0;
    // tensor-core variant (logmel_tc.cu): mel weights as bf16 (hi, lo) pairs in the tensor-memory A layout, and the
    // status word its bounded waits report protocol errors through
    uint32_t* d_tc_a = nullptr;   // [128][256] words
    uint32_t* d_tc_status = nullptr;
    bool tc_ok = false;           // the plan's shape fits the tensor-core kernel (n_mels <= 64, hop <= 512, ...)
    int tc_fft_warps = 8;  // measured on B200: 8 FFT warps at 224 registers beat 11 at 160 (c1 0.327 vs 0.334 ms, c2 3.58 vs 3.62 ms)
    hmfe::DescRing ring;
    int last_launches = 0;
    int sm_count = 148;
    // optional per-kernel timing (bench.py roofline): events recorded on the launch stream
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;  // triples: before power, after power, after finalize
};

