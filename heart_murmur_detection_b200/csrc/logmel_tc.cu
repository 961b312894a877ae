// Log-mel power kernel with the mel projection on the 5th-generation tensor cores (sm_100a): replaces
// librosa.feature.melspectrogram inside pre_process_audio_mel_t (/root/reference/src/util.py:481-492) like
// logmel.cu, with the `einsum('ft,mf->mt', S, mel_basis)` of that call as tcgen05.mma instead of shared-memory FMAs.
//
// One CTA per SM, 16 warps in 4 warpgroups with re-allocated registers (setmaxnreg):
//   warps 0-3    "epilogue": tensor-memory lane quarter q = warp; read the accumulators (tcgen05.ld), add the four
//                partial products, store the mel rows, clip max / min
//   warps 4..    NF "FFT" warps, each autonomous as in logmel.cu: claim 4-frame items, window + 1024-point FFT of two
//                packed transforms, power -> bf16 (hi, lo) pairs -> the warp's B tile in shared memory, then one
//                elected lane issues the 32 tcgen05.mma (K = 16 each) of the tile and commits them.  (A dedicated
//                MMA warp was measured first: ~15 dependent instructions per MMA at single-warp latency = ~2 600
//                cycles per tile, the whole kernel ran at the pace of that one warp, 0.60 ms on c1; issued from the
//                FFT warps the same instructions hide behind ten other warps.)
// Frames reach shared memory by 1-D bulk asynchronous copies (cp.async.bulk, SASS UBLKCP) issued one item ahead and
// completed on an mbarrier; no thread waits for global memory.
//
// The contraction D[128 x 16] = A[128 x 512] * B[16 x 512]^T per item:
//   A (tensor memory, loaded once per CTA): row 32q + i = mel 16q + i as bf16 "hi" (i < 16) or "lo" (i >= 16) part of
//     the float32 Slaney weight, w = hi + lo to 2^-17; K is the FFT bin PERMUTED so that a lane's 16 bins
//     (lane + 32 k1) are two 16-byte chunks: kappa = 8 lane + 256 (k1 / 8) + k1 % 8
//   B (shared memory, K-major, 128-byte swizzle, 8 rows stored): rows 0-3 = bf16 hi part of the power of frames 0-3,
//     rows 4-7 = lo part; rows 8-15 of the N = 16 instruction alias the next K atom, their D columns are never read
//   mel[m][frame j] = D[hi row m][j] + D[hi row m][4 + j] + D[lo row m][j] + D[lo row m][4 + j]
// i.e. all four products of (w_hi + w_lo)(p_hi + p_lo): relative error ~1e-5, inside the 1e-4 max|S| budget.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "api_common.h"
#include "logmel_batch.cuh"
#include "tables.h"
#include "tc_ptx.cuh"

#ifndef HMFE_TC_EPI_SPIN
#define HMFE_TC_EPI_SPIN 0
#endif

namespace hmfe {

using namespace tc;

constexpr int kRegionBytes = 19456;                        // per FFT warp: [B tile 8192 | frame staging 11264]
constexpr int kPowerBytes = 8192;                          //   the 16 896-byte FFT exchange tile overlays both
constexpr int kTmemCols = 512, kTmemD = 256, kDCols = 16;  // A: columns 0-255, accumulator of FFT warp w: 256 + 16 w
static_assert(32 * kXStride * 16 <= kRegionBytes, "exchange tile must fit its region");

enum : uint32_t {
    kErrRawWait = 1u, kErrFullWait = 2u, kErrFreeWait = 4u, kErrEpiSpin = 8u, kErrSmemAlign = 16u, kErrTmemBase = 32u,
};

// Item record, written to shared memory by lane 0 of the FFT warp when the item is located (while the warp's registers
// are free) and read back by whoever needs a field: all lanes (`packed`, before the frame fetch), lane 0 (the bulk copy),
// the epilogue warps (row0, clip, nvalid).  Keeping it in registers across the transform cost ~40 local-memory spill
// instructions per item at the 160 registers the FFT warps get (ncu: 18 % of their stall samples on those loads).
// Three records per warp, item i in record i % 3: the record of item i - 2 is free when item i + 1 is located, because
// the epilogue's release of item i - 2 was awaited before item i - 1 was handed over.
struct __align__(16) TcMeta {
    uint32_t row0;       // first output row of the item (frame index in the whole batch)
    int clip;
    uint32_t packed;     // nvalid | a << 3 | zlo << 5 | zhi << 17: frames that exist (0 = no item), staging index of span
                         // position 0 (0..3: keeps the bulk copy 16-byte aligned on both sides), span positions [0, zlo)
                         // and [zhi, span length) lie outside the clip (zero)
    uint32_t dst_bytes;  // bulk copy: staging index of the first copied float | bytes << 12 (0 bytes: nothing to copy)
    const float* src;    // bulk copy: global source, 16-byte aligned
    uint32_t item;
    uint32_t pad_;
};
static_assert(sizeof(TcMeta) == 32, "TcMeta is two 16-byte words");

struct TcTables {
    const float* win;
    const float2* tw;
    const uint32_t* a_words;  // [128][256]
};

constexpr int kRing = 16;          // > NF: at most one tile per FFT warp is in flight
constexpr uint32_t kSentinel = 0xffu;

// Shared-memory layout, as compile-time offsets from the (1024-byte aligned) start of the dynamic shared memory: every
// address in the kernel is then "symbol + constant (+ warp region)", no pointer lives in a register.
//   regions   per FFT warp, see kRegionBytes
//   tw, win   twiddle plane [32][32] float2, window 0.5 * Hann [1024]
//   full      per FFT warp: its MMAs are complete (the B tile may be overwritten)
//   dfree     per FFT warp: all four epilogue warps have read its accumulator and meta record
//   raw       per FFT warp: its frame copy has landed
//   done_seq / done_who   an in-order ring instead of polling per-warp barriers (an mbarrier test costs ~150 cycles:
//             sweeping 11 of them took longer than an item): an FFT warp appends its index when it issues the MMAs of a
//             tile and commits them to the slot's barrier; the epilogue warps consume the slots in ring order
template <int NF>
struct Lay {
    static constexpr uint32_t tw = NF * kRegionBytes;
    static constexpr uint32_t win = tw + 1024 * 8;
    static constexpr uint32_t full = win + 1024 * 4;
    static constexpr uint32_t dfree = full + NF * 8;
    static constexpr uint32_t raw = dfree + NF * 8;
    static constexpr uint32_t done_seq = raw + NF * 8;
    static constexpr uint32_t done_who = done_seq + kRing * 8;
    static constexpr uint32_t meta = (done_who + kRing * 4 + 15) & ~15u;
    static constexpr uint32_t tmem = meta + 3 * NF * sizeof(TcMeta);  // tensor-memory base, finished-warp count, ring tail
    static constexpr uint32_t done = tmem + 4;
    static constexpr uint32_t tail = tmem + 8;
    static constexpr uint32_t bytes = tmem + 16;
};
template <int NF>
constexpr size_t tc_smem_bytes() {
    return Lay<NF>::bytes;
}

extern __shared__ __align__(1024) uint8_t tc_smem[];

template <typename T>
HMFE_TC_D T* sptr(uint32_t off) { return reinterpret_cast<T*>(tc_smem + off); }
HMFE_TC_D uint32_t saddr(uint32_t off) { return smem_u32(tc_smem) + off; }

// The staging buffer of an FFT warp holds the span of an item's four frames, clip positions [P0, P0 + 3 hop + 1024) with
// P0 = f0 hop - 512, from staging index `a`; positions outside the clip (centre padding, frames beyond the last one) are
// ZERO there: one fetch path for interior and edge items, and no NaN bit pattern of stale shared memory can reach a
// transform that packs an existing frame.
constexpr int kStageFloats = 3 + (3 * 512 + kNfft) + 3 + 2;  // 2568: offset a, span at hop 512, round-up of the copy
static_assert(kStageFloats * 4 <= kRegionBytes - kPowerBytes, "staging area too small");

// bf16 (hi, lo) split of eight powers -> one 16-byte chunk each
HMFE_TC_D void split8(const float (&p)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float p0 = p[2 * e], p1 = p[2 * e + 1];
        h[e] = pack_bf16x2(p0, p1);
        const float h0 = __uint_as_float(h[e] << 16), h1 = __uint_as_float(h[e] & 0xffff0000u);
        l[e] = pack_bf16x2(p0 - h0, p1 - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// Exchange store as ONE 16-byte store per element: the two 8-byte halves of store_halves() (stride 16 bytes between
// lanes) are a 2-way bank conflict each unless ptxas merges them, which it does in logmel.cu but not here
// (ncu: 257 wavefronts per item for 64 STS.64 instead of 128).
HMFE_TC_D void exchange_store_v4(uint32_t base, const f32x2 (&re)[32], const f32x2 (&im)[32]) {
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(k2 * kXStride * 16)), "f"(re[k2].x),
                     "f"(re[k2].y), "f"(im[k2].x), "f"(im[k2].y)
                     : "memory");
}
HMFE_TC_D void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------ FFT warps
template <int NF, bool HOP512>
HMFE_TC_D void fft_role(const LogmelBatch& b, int w, int lane, int n_mels) {
    using V = f32x2;
    using L = Lay<NF>;
    constexpr int FR = 4;
    const uint32_t region = (uint32_t)w * kRegionBytes;
    const uint32_t tile_addr = saddr(region), stage_addr = tile_addr + kPowerBytes;
    const uint32_t bar_full = saddr(L::full + 8 * w), bar_free = saddr(L::dfree + 8 * w), bar_raw = saddr(L::raw + 8 * w);
    const int hop = HOP512 ? 512 : b.hop;
    const int span_len = 3 * hop + kNfft;
    constexpr uint64_t desc_hi = smem_desc(0, 0, 1024, kSwizzle128B);
    constexpr uint32_t idesc = idesc_bf16_f32(128, kDCols);

    // item indices fit 32 bits (the host checks n_items < 2^31 - 2^20 for this variant; the shared queue counter
    // is 64 bits wide): half the registers of the 64-bit bookkeeping of logmel.cu
    constexpr uint32_t kItemBlock = 8;
    const uint32_t it_end = (uint32_t)min(item_count(b), (int64_t)0x7ff00000);
    auto claim = [&]() -> uint32_t {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(b.queue, (unsigned long long)kItemBlock);
        v = __shfl_sync(0xffffffffu, v, 0);
        return (uint32_t)min(v, (unsigned long long)0x7ff00000u);
    };
    auto next_item = [&](uint32_t item) -> uint32_t {  // blocks are aligned to kItemBlock: no block-end register
        if ((item + 1) % kItemBlock != 0) return item + 1;
        if (item >= it_end) return item;
        return claim();
    };
    // a wait that gave up (protocol error) is reported once; the warp then stops waiting so that the kernel ends
    bool alive = true;
    auto wait = [&](uint32_t bar, uint32_t parity, uint32_t err) {
        if (alive && !mbar_wait(bar, parity)) {
            alive = false;
            if (lane == 0) atomicOr(b.status, err);
        }
    };
    int64_t clip_cursor = -1;
    // Everything about an item is worked out while the warp's registers are free (between two transforms) and
    // parked in the warp's record `rec` in shared memory; the copy itself is started later, when the previous item
    // has left the staging buffer.
    TcMeta* recs = sptr<TcMeta>(L::meta) + 3 * w;
    auto locate = [&](uint32_t item, uint32_t rec) {
        const ItemCtx c = locate_item<FR>(b, n_mels, (int64_t)item, (int64_t)it_end, clip_cursor);
        TcMeta m;
        m.row0 = 0;
        m.clip = (int)c.clip;
        m.packed = 0;
        m.dst_bytes = 0;
        m.src = b.wav;
        m.item = item;
        m.pad_ = 0;
        if (c.valid) {
            m.row0 = (uint32_t)((c.o - b.out) / n_mels) + (uint32_t)c.f0;
            const int P0 = c.f0 * hop - kNfft / 2;
            const int first = max(0, P0), end = min(c.nsamp, P0 + span_len);
            uint32_t a = 0, zlo = (uint32_t)span_len, zhi = (uint32_t)span_len;
            if (end > first) {
                const float* g = c.x + first;
                const int skip = (int)((reinterpret_cast<uintptr_t>(g) & 15) >> 2);
                const int d0 = first - P0 - skip;  // span position of the first copied sample (>= -3)
                a = (uint32_t)((-d0) & 3);
                zlo = (uint32_t)(first - P0);
                zhi = (uint32_t)(end - P0);
                m.src = g - skip;
                m.dst_bytes = (uint32_t)((int)a + d0) | ((uint32_t)(((end - first + skip) * 4 + 15) & ~15) << 12);
            }
            m.packed = (uint32_t)min(4, c.T - c.f0) | (a << 3) | (zlo << 5) | (zhi << 17);
        }
        if (lane == 0) recs[rec] = m;
    };
    auto start_copy = [&](uint32_t rec) {  // lane 0
        const TcMeta m = recs[rec];
        if ((m.packed & 7u) == 0) return;
        const uint32_t bytes = m.dst_bytes >> 12;
        if (bytes) {
            mbar_expect_tx(bar_raw, bytes);
            bulk_g2s(stage_addr + 4u * (m.dst_bytes & 0xfffu), m.src, bytes, bar_raw);
        } else {
            mbar_arrive(bar_raw);
        }
    };
    auto rec_after = [](uint32_t rec) { return rec == 2 ? 0u : rec + 1; };

    if (b.stagger_ns > 0) __nanosleep((unsigned)(w * b.stagger_ns));
    uint32_t item = claim(), rec = 0, n_done = 0;
    locate(item, 0);
    if (lane == 0) start_copy(0);
    uint32_t nitem = next_item(item);
    locate(nitem, 1);

    while (item < it_end) {
        __syncwarp();  // lane 0's record
        const uint32_t packed = recs[rec].packed;
        const uint32_t a = (packed >> 3) & 3u, zlo = (packed >> 5) & 0xfffu, zhi = packed >> 17;
        float* stage = sptr<float>(region + kPowerBytes) + a;  // span position 0
        wait(bar_raw, n_done & 1, kErrRawWait);
        if (zlo > 0 || zhi < (uint32_t)span_len) {  // edge item: zeros outside the clip (after the copy, which rounds outwards)
            for (uint32_t j = lane; j < zlo; j += 32) stage[j] = 0.0f;
            for (uint32_t j = zhi + lane; j < (uint32_t)span_len; j += 32) stage[j] = 0.0f;
            __syncwarp();
        }
        V re[32], im[32];
        {
            const float* sp = stage + lane;  // sample `lane` of the item's first frame
            const float* win = sptr<float>(L::win) + lane;
            if (HOP512) {
                // 50 % overlap: frame j = half-frames (j, j + 1) of the span, every staged sample is read once
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    float h5[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) h5[q] = sp[512 * q + 32 * m];
                    // (win[n + 512] = 0.5 - win[n] would save this second read, but loses the relative accuracy of the
                    // window's small tail values: mel bins 80 dB below the clip maximum moved by 2e-3 relative)
                    const float w0 = win[32 * m], w1 = win[32 * (m + 16)];
                    re[brev(m, 5)] = vmuls(V{h5[0], h5[2]}, w0);
                    im[brev(m, 5)] = vmuls(V{h5[1], h5[3]}, w0);
                    re[brev(m + 16, 5)] = vmuls(V{h5[1], h5[3]}, w1);
                    im[brev(m + 16, 5)] = vmuls(V{h5[2], h5[4]}, w1);
                }
            } else {
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) {
                    const float wv = win[32 * n2];
                    const float* q = sp + 32 * n2;
                    re[brev(n2, 5)] = vmuls(V{q[0], q[2 * hop]}, wv);
                    im[brev(n2, 5)] = vmuls(V{q[hop], q[3 * hop]}, wv);
                }
            }
        }
        __syncwarp();
        // both 32-point passes run the same unrolled butterfly code (one copy in the instruction cache)
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            fft_dit<32, V>(re, im);
            if (pass == 0) {
                apply_twiddle<V>(lane, sptr<const float2>(L::tw), re, im);
                // the tensor core must be done with the previous B tile before the exchange overwrites it
                wait(bar_full, n_done & 1, kErrFullWait);  // completion n_done (completion 0 is the arrival at start-up)
                exchange_store_v4(tile_addr + 16u * (uint32_t)lane, re, im);
                __syncwarp();
                exchange_load<V>(lane, sptr<const xelem<V>>(region), re, im);
                __syncwarp();
                // the staging area is free again: start the copy of the next item's frames
                fence_proxy_async();
                if (lane == 0) start_copy(rec_after(rec));
            }
        }
        {   // separation of the packed frames, power, bf16 (hi, lo) split, B tile rows
            const int src = (32 - lane) & 31;
            const uint32_t row_chunk = (uint32_t)(lane >> 3) * 1024u;
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
                float f0p[8], f1p[8], f2p[8], f3p[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k1 = 8 * j8 + e;
                    const V give_r = lane == 0 ? re[(32 - k1) & 31] : re[31 - k1];
                    const V give_i = lane == 0 ? im[(32 - k1) & 31] : im[31 - k1];
                    const V pr = shfl(give_r, src), pi = shfl(give_i, src);
                    const xelem<V> pw = frame_powers<V>(re[k1], im[k1], pr, pi);
                    f0p[e] = pw.a.x;
                    f2p[e] = pw.a.y;
                    f1p[e] = pw.b.x;
                    f3p[e] = pw.b.y;
                }
                const uint32_t atom = tile_addr + row_chunk + (uint32_t)j8 * 4096u;
                uint4 hi, lo;
#define HMFE_TC_STORE_ROW(J, P)                                                        \
    split8(P, hi, lo);                                                                 \
    sts128(atom + (J) * 128u + (uint32_t)(((lane & 7) ^ (J)) << 4), hi);               \
    sts128(atom + ((J) + 4) * 128u + (uint32_t)(((lane & 7) ^ ((J) + 4)) << 4), lo);
                HMFE_TC_STORE_ROW(0, f0p)
                HMFE_TC_STORE_ROW(1, f1p)
                HMFE_TC_STORE_ROW(2, f2p)
                HMFE_TC_STORE_ROW(3, f3p)
#undef HMFE_TC_STORE_ROW
            }
        }
        // the epilogue must have consumed the previous item's accumulator and meta record
        wait(bar_free, n_done & 1, kErrFreeWait);
        fence_proxy_async();  // the B tile (generic-proxy stores of all lanes) becomes visible to the tensor core
        __syncwarp();
        tc_fence_after();     // after the epilogue's tcgen05.ld of this accumulator (its mbarrier arrive was observed)
        if (elect_one()) {
            const uint32_t slot = atomicAdd(sptr<uint32_t>(L::tail), 1u) % kRing;
            sptr<uint32_t>(L::done_who)[slot] = (uint32_t)w | (rec << 8);
            __threadfence_block();  // meta record and ring entry before the barrier the commits complete
            const uint32_t desc_lo = tile_addr >> 4;
            const uint32_t d_tmem = kTmemD + kDCols * w;  // tensor-memory base is 0 (all 512 columns are ours; checked at start)
#pragma unroll
            for (int k = 0; k < 32; ++k)
                mma_ts_f16(d_tmem, 8 * k, desc_hi | (uint64_t)(desc_lo + (k >> 2) * 64 + (k & 3) * 2), idesc, k > 0);
            mma_commit(saddr(L::done_seq) + 8u * slot);  // -> epilogue warps, in ring order
            mma_commit(bar_full);                        // -> this warp: its B tile may be overwritten
        }
        __syncwarp();
        ++n_done;
        // registers are free here: locate the item after the next one
        rec = rec_after(rec);
        item = nitem;
        nitem = next_item(item);
        locate(nitem, rec_after(rec));
    }
    // all of this warp's accumulators have been produced before it reports completion
    wait(bar_full, n_done & 1, kErrFullWait);
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(sptr<uint32_t>(L::done), 1u) == (uint32_t)NF - 1) {  // the last FFT warp closes the ring
            const uint32_t slot = atomicAdd(sptr<uint32_t>(L::tail), 1u) % kRing;
            sptr<uint32_t>(L::done_who)[slot] = kSentinel;
            mbar_arrive(saddr(L::done_seq) + 8u * slot);
        }
    }
}

// ------------------------------------------------------------------------------------------------ epilogue warps
template <int NF>
HMFE_TC_D void epilogue_role(const LogmelBatch& b, uint32_t tmem, int q, int lane, int n_mels) {
    using L = Lay<NF>;
    const int col = 16 * q + (lane & 15);
    const int fsel = lane >> 4;  // lanes 0-15 store frames 0 and 1, lanes 16-31 frames 2 and 3
    const uint32_t done_seq = saddr(L::done_seq);
    for (uint32_t seq = 0; seq < (1u << 30); ++seq) {
        const uint32_t slot = seq % kRing;
        if (!(HMFE_TC_EPI_SPIN ? mbar_spin(done_seq + 8u * slot, (seq / kRing) & 1u) : mbar_wait(done_seq + 8u * slot, (seq / kRing) & 1u))) break;
        const uint32_t who = *reinterpret_cast<volatile uint32_t*>(sptr<uint32_t>(L::done_who) + slot);
        if (who == kSentinel) return;
        const uint32_t w = who & 0xffu;
        tc_fence_after();
        const uint4 m4 = *reinterpret_cast<const uint4*>(sptr<TcMeta>(L::meta) + 3 * w + (who >> 8));
        const uint32_t row0 = m4.x, clip = m4.y;
        const int nvalid = (int)(m4.z & 7u);
        uint32_t v[8];
        tmem_ld8(tmem + kTmemD + kDCols * w + ((uint32_t)(32 * q) << 16), v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(saddr(L::dfree) + 8u * w);
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d[j] = __uint_as_float(v[j]) + __uint_as_float(v[4 + j]);
            d[j] += __shfl_xor_sync(0xffffffffu, d[j], 16);
            d[j] = fmaxf(d[j], 0.0f);
        }
        const float va = fsel ? d[2] : d[0], vb = fsel ? d[3] : d[1];
        const int fa = 2 * fsel, fb = fa + 1;
        const bool oka = col < n_mels && fa < nvalid, okb = col < n_mels && fb < nvalid;
        float* out = b.out + (int64_t)row0 * n_mels;
        if (oka) out[(int64_t)fa * n_mels + col] = va;
        if (okb) out[(int64_t)fb * n_mels + col] = vb;
        uint32_t hi = 0u, lo = 0x7f800000u;
        if (oka) {
            hi = __float_as_uint(va) & 0x7fffffffu;
            lo = hi;
        }
        if (okb) {
            const uint32_t bb = __float_as_uint(vb) & 0x7fffffffu;
            hi = max(hi, bb);
            lo = min(lo, bb);
        }
        hi = __reduce_max_sync(0xffffffffu, hi);  // non-negative floats order like their bit patterns
        lo = __reduce_min_sync(0xffffffffu, lo);
        if (lane == 0) {
            atomicMax(b.stats + 2 * (int64_t)clip, hi);
            atomicMin(b.stats + 2 * (int64_t)clip + 1, lo);
        }
    }
    if (lane == 0) atomicOr(b.status, kErrEpiSpin);
}

// 4 epilogue warps + NF FFT warps, in whole warpgroups: setmaxnreg acts on warpgroups, and a block with a partial
// warpgroup is charged registers for the whole of it (480 threads at 136 registers: "too many resources requested").
//   NF = 11: 512 threads launched at 128 registers; epilogue warpgroup -> 32, the other three warpgroups -> 160
//            (128 * 32 + 384 * 160 = 65 536); warp 15 has no region in shared memory and idles
//   NF = 8 : 384 threads launched at 168 registers; epilogue -> 32, the two FFT warpgroups -> 224 (61 440)
template <int NF>
struct TcCfg {
    static constexpr int threads = NF > 8 ? 512 : 384;
    static constexpr int fft_regs = NF > 8 ? 160 : 224;
    static constexpr int epi_regs = 32;
};

template <int NF, bool HOP512>
__global__ void __launch_bounds__(TcCfg<NF>::threads, 1) logmel_tc_kernel(const LogmelBatch b, const TcTables tb, int n_mels) {
    constexpr int kTcThreads = TcCfg<NF>::threads;
    using L = Lay<NF>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((smem_u32(tc_smem) & 1023u) != 0) {  // the swizzled B tiles need the 1024-byte alignment the declaration asks for
        if (threadIdx.x == 0) atomicOr(b.status, kErrSmemAlign);
        return;
    }

    for (int i = threadIdx.x; i < 1024; i += kTcThreads) {
        sptr<float2>(L::tw)[i] = tb.tw[i];
        sptr<float>(L::win)[i] = tb.win[i];
    }
    if (threadIdx.x == 0) {
        for (int w = 0; w < NF; ++w) {
            mbar_init(saddr(L::full + 8 * w), 1);
            mbar_init(saddr(L::dfree + 8 * w), 4);
            mbar_init(saddr(L::raw + 8 * w), 1);
            // phase 0 of "B tile free" and "accumulator free" completes here: the first item of a warp waits like any other
            mbar_arrive(saddr(L::full + 8 * w));
            for (int q = 0; q < 4; ++q) mbar_arrive(saddr(L::dfree + 8 * w));
        }
        for (int i = 0; i < kRing; ++i) mbar_init(saddr(L::done_seq + 8 * i), 1);
        *sptr<uint32_t>(L::done) = 0;
        *sptr<uint32_t>(L::tail) = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(saddr(L::tmem), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sptr<uint32_t>(L::tmem);
    if (warp < 4) {  // the mel weights: row 32 warp + lane of A, 256 words
        const uint32_t* row = tb.a_words + (32 * warp + lane) * 256;
        for (int c = 0; c < 256; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(row + c + j);
            tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + c, v);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4) {
        setmaxnreg_dec<TcCfg<NF>::epi_regs>();
        epilogue_role<NF>(b, tmem, warp, lane, n_mels);
    } else {
        setmaxnreg_inc<TcCfg<NF>::fft_regs>();
        if (tmem != 0) {  // cannot happen while the CTA owns all 512 columns; the roles assume base 0
            if (threadIdx.x == 128) {  // report, and close the ring so that the epilogue warps return
                atomicOr(b.status, kErrTmemBase);
                sptr<uint32_t>(L::done_who)[0] = kSentinel;
                mbar_arrive(saddr(L::done_seq));
            }
        } else if (warp - 4 < NF) {
            fft_role<NF, HOP512>(b, warp - 4, lane, n_mels);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
static uint16_t f32_to_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf16_to_f32(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// A operand: [128 rows][256 words]; row 32q + i = mel 16q + (i % 16), hi part for i < 16, lo part for i >= 16;
// word c of a row = K positions (2c, 2c + 1), position kappa <-> FFT bin (kappa % 256) / 8 + 32 (8 (kappa / 256) + kappa % 8)
std::vector<uint32_t> tc_build_a_words(const std::vector<float>& mel_dense, int n_mels, int n_bins) {
    std::vector<uint32_t> a(128 * 256, 0u);
    for (int r = 0; r < 128; ++r) {
        const int q = r / 32, i = r % 32, mel = 16 * q + (i % 16);
        if (mel >= n_mels) continue;
        for (int kappa = 0; kappa < 512; ++kappa) {
            const int lane = (kappa % 256) / 8, k1 = 8 * (kappa / 256) + kappa % 8;
            const int bin = lane + 32 * k1;
            const float wv = mel_dense[(size_t)mel * n_bins + bin];
            const uint16_t hi = f32_to_bf16_rn(wv);
            const uint16_t part = i < 16 ? hi : f32_to_bf16_rn(wv - bf16_to_f32(hi));
            a[r * 256 + kappa / 2] |= (uint32_t)part << (16 * (kappa & 1));
        }
    }
    return a;
}

bool tc_shape_ok(const hmfe_logmel_plan* p) {
    if (p->n_fft != kNfft || p->n_mels > 64 || p->hop > 512 || p->hop < 1) return false;
    for (int m = 0; m < p->n_mels; ++m)  // the Nyquist bin is outside K = 512: its weight must be zero
        if (p->mel_dense[(size_t)m * p->n_bins + 512] != 0.0f) return false;
    return true;
}

template <int NF, bool HOP512>
static int launch_tc_n(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    const size_t smem = tc_smem_bytes<NF>();
    auto kern = logmel_tc_kernel<NF, HOP512>;
    HMFE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (b.n_items + NF - 1) / NF;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count));
    TcTables tb{p->d_win, p->d_tw, p->d_tc_a};
    kern<<<grid, TcCfg<NF>::threads, smem, st>>>(b, tb, p->n_mels);
    HMFE_CHECK_CUDA(cudaGetLastError());
    return HMFE_OK;
}

int launch_logmel_tc(hmfe_logmel_plan* p, const LogmelBatch& b, cudaStream_t st) {
    HMFE_REQUIRE(p->tc_ok && p->d_tc_a && p->pad_mode == HMFE_PAD_CONSTANT,
                 "the tensor-core log-mel variant needs n_mels <= 64, hop <= 512 and constant padding");
    HMFE_REQUIRE(b.n_items_dev != nullptr || b.n_items < (int64_t)0x7ff00000,
                 "the tensor-core log-mel variant indexes work items with 32 bits: %lld items in one call", (long long)b.n_items);
    const bool h512 = p->hop == 512;
    if (p->tc_fft_warps <= 8) return h512 ? launch_tc_n<8, true>(p, b, st) : launch_tc_n<8, false>(p, b, st);
    return h512 ? launch_tc_n<11, true>(p, b, st) : launch_tc_n<11, false>(p, b, st);
}

}  // namespace hmfe
