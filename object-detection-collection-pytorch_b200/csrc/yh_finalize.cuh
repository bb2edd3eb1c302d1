// Last step of a train-head call, by ONE warp after the train kernel has completed: the six 64-bit
// fixed-point totals -> the five terms and the loss (reference models/yolov2.py:1132-1138), and the
// workspace back to zero for the next call.
//
// Sharded batches (world > 1): the only data-path exchange of the whole path -- the six partial
// sums of every rank -- happens HERE, inside the kernel, over NVLink peer memory: the warp stores its rank's
// sums into its slot of every rank's exchange buffer (plain peer stores, 56 bytes per peer), publishes them
// with a sequence word, waits for the sequence words of all ranks in its own buffer and adds the slots up in
// rank order.  Integer sums: every rank obtains the same bits, and the same bits as one GPU over the whole
// batch.  No NCCL call, no extra launch, no host round trip; the wait is hidden behind the next kernel of the
// launch chain.  Exchange buffer layout (uint64 words): [parity 0/1][rank < YH_MAX_RANKS][8] = six sums, flags,
// sequence; word 2*YH_MAX_RANKS*8: the local call counter.  Two parities: a rank can be at most one call ahead
// of a peer that has not read the previous slot yet.
#pragma once

#include "yh_common.cuh"

constexpr int kYhXchSeqWord = 2 * YH_MAX_RANKS * 8;
constexpr size_t kYhXchBytes = (size_t)(kYhXchSeqWord + 8) * 8;

__device__ __forceinline__ void yh_st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long yh_ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long yh_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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
        const int par = (int)(seq & 1ull);
        const int slot = (par * YH_MAX_RANKS + f.rank) * 8;
        for (int q = 0; q < f.world; ++q)
            if (lane < 7) yh_st_sys(f.peer[q] + slot + lane, a);
        __threadfence_system();
        __syncwarp();
        if (lane < f.world) {  // publish: the slot of rank q is complete
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.peer[lane] + slot + 7), "l"(seq) : "memory");
        }
        // wait for every rank's slot in this rank's buffer (bounded: a rank that never arrives turns the
        // loss into NaN after 4 s instead of hanging the device)
        bool ok = true;
        if (lane < f.world) {
            const unsigned long long* flag = mine + (par * YH_MAX_RANKS + lane) * 8 + 7;
            const unsigned long long t0 = yh_globaltimer();
            while (yh_ld_acquire_sys(flag) != seq) {
                if (yh_globaltimer() - t0 > 4000000000ull) { ok = false; break; }
                __nanosleep(64);
            }
        }
        lost = !__all_sync(0xffffffffu, ok);
        __threadfence_system();
        unsigned long long tot = 0ull;
        if (lane < 7)
            for (int r = 0; r < f.world; ++r) {
                const unsigned long long v = yh_ld_sys(mine + (par * YH_MAX_RANKS + r) * 8 + lane);
                tot = lane == 6 ? (tot | v) : tot + v;
            }
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
