// Last step of a train-head call, by ONE warp after the train kernel has completed: the six 64-bit
// fixed-point totals -> the five terms and the loss (reference models/yolov2.py:1132-1138), and the
// workspace back to zero for the next call.
//
// Sharded batches (world > 1): the only data-path exchange of the whole path -- the six partial
// sums of every rank -- happens HERE, inside the kernel, over NVLink peer memory, with a low-latency protocol:
// every 8-byte word a rank stores into a peer's buffer carries 32 bits of payload and the 32-bit sequence number
// of the call (8-byte stores are single-copy atomic), so a word validates itself: no fence, no separate flag, one
// NVLink store latency end to end.  The warp stores its rank's 7 values (six sums + flags) as 14 such words into
// its slot of every rank's buffer, polls the 14 words of every rank's slot in its OWN buffer until they carry this
// call's sequence number, and adds the slots up in rank order.  Integer sums: every rank obtains the same bits, and
// the same bits as one GPU over the whole batch.  No NCCL call, no extra launch, no host round trip.
// Exchange buffer layout (uint64 words): [parity 0/1][rank < YH_MAX_RANKS][16]; word 2*YH_MAX_RANKS*16: the local
// call counter.  Two parities: a rank can be at most one call ahead of a peer that has not read its previous slot.
#pragma once

#include "yh_common.cuh"

constexpr int kYhXchSlotWords = 16;
constexpr int kYhXchSeqWord = 2 * YH_MAX_RANKS * kYhXchSlotWords;
constexpr size_t kYhXchBytes = (size_t)(kYhXchSeqWord + 8) * 8;

__device__ __forceinline__ void yh_st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long yh_ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long yh_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All 32 lanes of one warp; the train kernel of this call has completed and its sums are visible.
__device__ __forceinline__ void yh_finalize_warp(const YhFinalParams& f, int lane) {
    constexpr double kFix = 4294967296.0;
    unsigned long long a = lane < 7 ? __ldcg(f.acc + lane) : 0ull;  // lanes 0..5 sums, lane 6 flags
    bool lost = false;
    if (f.world > 1) {
        unsigned long long* mine = f.peer[f.rank];
        const unsigned long long seq = yh_ld_sys(mine + kYhXchSeqWord) + 1ull;
        const unsigned seq32 = (unsigned)seq | 0x80000000u;  // (never zero: a zero-filled buffer matches no call)
        const int par = (int)(seq & 1ull);
        // lane l < 14: half (l & 1) of value (l >> 1), tagged with the sequence number
        const unsigned long long val = __shfl_sync(0xffffffffu, a, lane >> 1);
        const unsigned half = (lane & 1) ? (unsigned)(val >> 32) : (unsigned)val;
        const unsigned long long word = ((unsigned long long)seq32 << 32) | half;
        const int slot = (par * YH_MAX_RANKS + f.rank) * kYhXchSlotWords;
        if (lane < 14)
            for (int q = 0; q < f.world; ++q) yh_st_sys(f.peer[q] + slot + lane, word);
        // every rank's slot in this rank's buffer: the 14 * world words are polled by all lanes AT ONCE (word w on lane
        // w mod 32) -- one local round trip once the last rank's words are in, not one per rank -- and handed over
        // through shared memory; bounded: a rank that never arrives turns the loss into NaN after 4 s instead of
        // hanging the device
        __shared__ unsigned s_pay[YH_MAX_RANKS * 14];
        constexpr int kMaxW = (14 * YH_MAX_RANKS + 31) / 32;
        const int nw = 14 * f.world;
        const unsigned long long* base = mine + (size_t)par * YH_MAX_RANKS * kYhXchSlotWords;
        unsigned long long wv[kMaxW];
#pragma unroll
        for (int i = 0; i < kMaxW; ++i) {
            const int w = lane + 32 * i;
            wv[i] = w < nw ? yh_ld_sys(base + (w / 14) * kYhXchSlotWords + (w % 14)) : ((unsigned long long)seq32 << 32);
        }
        bool ok = true;
        const unsigned long long t0 = yh_globaltimer();
        for (;;) {
            bool all = true;
#pragma unroll
            for (int i = 0; i < kMaxW; ++i) {
                const int w = lane + 32 * i;
                if (w < nw && (unsigned)(wv[i] >> 32) != seq32) {
                    wv[i] = yh_ld_sys(base + (w / 14) * kYhXchSlotWords + (w % 14));
                    all = all && (unsigned)(wv[i] >> 32) == seq32;
                }
            }
            if (all) break;
            if (yh_globaltimer() - t0 > 4000000000ull) { ok = false; break; }
        }
#pragma unroll
        for (int i = 0; i < kMaxW; ++i) {
            const int w = lane + 32 * i;
            if (w < nw) s_pay[w] = (unsigned)wv[i];
        }
        __syncwarp();
        // value v (lanes 0..6) = sum over the ranks, in rank order, of hi:lo (the flags: or)
        unsigned long long tot = 0ull;
        if (lane < 7) {
            for (int r = 0; r < f.world; ++r) {
                const unsigned long long v = ((unsigned long long)s_pay[r * 14 + 2 * lane + 1] << 32) | s_pay[r * 14 + 2 * lane];
                tot = lane == 6 ? (tot | v) : tot + v;
            }
        }
        lost = !__all_sync(0xffffffffu, ok);
        a = tot;
        if (lane == 0) yh_st_sys(mine + kYhXchSeqWord, seq);
    }
    unsigned long long v[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) v[q] = __shfl_sync(0xffffffffu, a, q);
    if (lane == 0) {
        double tot[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) tot[q] = (double)v[q] / kFix;
        const unsigned bad = lost ? 63u : (unsigned)v[6];
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const double t0 = (bad & 1u) ? nan : tot[0] * f.inv_den[0];
        const double t1 = (bad & 2u) ? nan : tot[1] * f.inv_den[1];
        const double t2 = (bad & 4u) ? nan : tot[2] * f.inv_den[2];
        const double t3 = (bad & 24u) ? nan : (tot[3] - tot[4]) * f.inv_den[3];
        const double t4 = (bad & 32u) ? nan : tot[5] * f.inv_den[4];
        f.terms[0] = (float)t0; f.terms[1] = (float)t1; f.terms[2] = (float)t2;
        f.terms[3] = (float)t3; f.terms[4] = (float)t4;
        f.loss[0] = (float)(f.lam[0] * t0 + f.lam[1] * t1 + f.lam[2] * t2 + f.lam[3] * t3 + f.lam[4] * t4);
#pragma unroll
        for (int q = 0; q < 8; ++q) f.acc[q] = 0ull;  // ready for the next launch
    }
}
