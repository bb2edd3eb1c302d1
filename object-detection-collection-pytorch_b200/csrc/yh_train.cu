// Fused YOLO train head: decode + responsible-predictor assignment + five loss terms + dL/dy
// in one pass over the head tensor, ONE launch per batch.
//
// Replaces YOLOv2.get_loss / YOLOv1.get_loss and their autograd backward
// (reference models/yolov2.py:747-1140, models/yolov1.py:556-931, models/utils.py:5-65).
//
// The kernel is HBM-bound (~60 flop per 200 bytes) and, at the batch sizes the reference trains
// with, only a few microseconds long.  Under a saturated memory system every dependent global
// load costs microseconds (measured: 2-4 us per round trip once ~80 KB per SM are in flight), so
// the kernel is organised around ONE round trip for everything the sparse part needs:
//   * the flattened batch of grid cells is cut into tiles of a few dozen cells (a multiple of 4
//     cells, so every tile starts 16-byte aligned; one tile per CTA at the headline size, a
//     grid-stride loop over tiles for larger tensors; 3 CTAs of 8 streaming warps + 1 record warp
//     per SM: 72 registers, no spills -- a local-memory reload after the dense pass queues behind the
//     draining gradient stores like any other memory access);
//   * the record warp's first lane is the TMA producer: one bulk copy brings a speculative window
//     of ground-truth records (placed where the tile's records sit if boxes are spread evenly over
//     the images) and goes out FIRST, then the tile itself follows in chunks, each on its own
//     mbarrier, a bounded number of chunks in flight;
//   * dense pass (streaming warps): as chunks land, every thread reads float4s from shared memory
//     and writes the matching float4 of dL/dy straight from registers with coalesced 16-byte
//     stores.  dL/dy is zero everywhere except the objectness channel, whose no-object gradient
//     depends only on the thread's own float4 (at most one objectness logit falls into any 4
//     consecutive floats): every byte of y is read once and every byte of dy written once;
//   * sparse pass (concurrently): the record warp picks the tile's records from the window in CSR
//     order, asks for a private early copy of their cells (so a record of the tile's last chunk
//     does not wait for the end of the stream), and publishes the list; the records are dealt to
//     the warps, one warp per record: lanes decode the A boxes of the record's cell from shared
//     memory, IoU against the record, redux-argmax picks the responsible predictor, the five
//     lanes of that predictor finish one channel each, lanes then cover the C classes for the
//     softmax/class term; the 5+C gradient values wait in a shared-memory patch;
//   * after a CTA barrier the patches overwrite their rows (dense value + record gradient; if two
//     records share a cell, in CSR order with read-modify-write, so collisions on one predictor
//     accumulate deterministically);
//   * the six sums of a CTA are added onto 64-bit fixed-point accumulators with integer atomics
//     (exact, order-independent: deterministic); yh_train_finalize_kernel, a one-warp programmatic
//     dependent launched right behind this kernel, turns the totals into terms and loss and
//     re-zeroes the workspace (a last-CTA ticket inside this kernel costs three dependent L2 round
//     trips at the very end of its critical path; -DYH_X_ONE_KERNEL keeps that variant);
//   * launched as a programmatic dependent: the prologue overlaps the previous kernel's tail.
#include <limits.h>
#include <string.h>

#include "yh_common.cuh"
#include "yh_finalize.cuh"
#include "yh_record.cuh"

namespace {

#ifndef YH_X_DENSE
#define YH_X_DENSE 256
#endif
constexpr int kDense = YH_X_DENSE;        // streaming threads per CTA
constexpr int kDenseWarps = kDense / 32;
constexpr int kThreads = kDense + 32;     // + one record warp
constexpr int kWarps = kThreads / 32;
#ifndef YH_X_CTAS
#define YH_X_CTAS 3
#endif
constexpr int kCtasPerSm = YH_X_CTAS;
#ifndef YH_X_CHUNKS
#define YH_X_CHUNKS 4
#endif
#ifndef YH_X_AHEAD
#define YH_X_AHEAD 2
#endif
constexpr int kChunks = YH_X_CHUNKS;      // TMA chunks (mbarriers) per tile
constexpr int kAhead = YH_X_AHEAD;        // chunks in flight per CTA
constexpr int kTileBytesMax = (kCtasPerSm >= 4 ? 40 : (kCtasPerSm == 3 ? 52 : 80)) * 1024;  // shared-memory stage of one tile
constexpr int kMaxGrid = 2048;
constexpr int kWindow = 128;              // speculative record window (records) per tile
constexpr int kWindowMax = 256;           // record window buffer: an exact window after a miss may be this long
#ifndef YH_X_SLOTS
#define YH_X_SLOTS 24
#endif
#ifndef YH_X_SHADOW
#define YH_X_SHADOW 12
#endif
constexpr int kSlots = YH_X_SLOTS;        // patches (records processed during the dense pass) per tile, <= 32
constexpr int kPatchFloats = 32;          // floats per patch row: 5 + C must fit (else the record waits for the end)
constexpr int kShadowMax = YH_X_SHADOW;   // records per tile up to which they are processed during the dense pass
constexpr int kCellSlots = 12;            // records whose cell gets a private early copy (the others wait for their chunk)

#ifdef YH_X_TRACE
__device__ unsigned long long g_xtrace[4096 * 24];
__device__ unsigned int g_rcycles[4096 * 16];
__device__ __forceinline__ unsigned long long xt_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ unsigned long long g_tiletrace[4096 * 32];
#define XTT(k, w) do { if (threadIdx.x == 0 && (k) < 10) g_tiletrace[blockIdx.x * 32 + 1 + 3 * (k) + (w)] = xt_now(); } while (0)
#define XT_DECL unsigned long long xt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define XT(slot) do { if ((threadIdx.x & 31) == 0 || threadIdx.x == kThreads - 1) xt_[slot] = xt_now(); } while (0)
#define XT_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 8; ++i_) g_xtrace[blockIdx.x * 24 + i_] = xt_[i_]; \
                      if (threadIdx.x == kDense) for (int i_ = 0; i_ < 8; ++i_) g_xtrace[blockIdx.x * 24 + 8 + i_] = xt_[i_]; \
                      if (threadIdx.x == kThreads - 1) for (int i_ = 0; i_ < 8; ++i_) g_xtrace[blockIdx.x * 24 + 16 + i_] = xt_[i_]; } while (0)
#else
#define XTT(k, w) do { } while (0)
#define XT_DECL
#define XT(slot) do { } while (0)
#define XT_FLUSH
#endif

struct TrainParams {
    YhGeom g;
    const float* y;
    float* dy;
    const YhGt* gt;
    const int32_t* gt_off;
    float* terms;
    float* loss;
    int32_t* resp;
    float* iou_resp;
    unsigned long long* acc;  // [8]: six fixed-point sums, [6] non-finite flags, [7] ticket
    int total_cells;
    int tile_cells;       // cells per tile (multiple of 4)
    int num_tiles;
    int m_local;          // records in gt
    int late_wait;        // *_overlapped entry points: wait for the previous kernels before publishing, not at the start
    int cell_slots;       // private cell copies per tile that fit in shared memory (<= kCellSlots)
    int cell_slot_floats; // floats per private cell copy (cell + alignment slack, multiple of 4)
    long long total_floats;
    float rec_per_cell;   // m_local / total_cells: where a tile's records sit if boxes are spread evenly
    float lam[5];
    double inv_den[5];    // 1/(2M), 1/(2M), 1/M, 1/(M(P-1)), 1/M
    float cxy, cwh, cconf, cno, ccls;  // gradient coefficients (see train_impl)
};

template <bool VEC>
__device__ __forceinline__ float4 load4(const float* base, int idx4) {
#ifdef YH_X_CS
    if (VEC) return __ldcs(reinterpret_cast<const float4*>(base) + idx4);
#endif
    if (VEC) return __ldg(reinterpret_cast<const float4*>(base) + idx4);
    const float* s = base + 4 * idx4;
    return make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
}
template <bool VEC>
__device__ __forceinline__ void store4(float* base, int idx4, const float4& v) {
    if (VEC) {
#ifdef YH_X_CS
        __stcs(reinterpret_cast<float4*>(base) + idx4, v);
#else
        reinterpret_cast<float4*>(base)[idx4] = v;
#endif
    } else {
        float* d = base + 4 * idx4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}
__device__ __forceinline__ float pick4(const float4& v, int j) {
    return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}
__device__ __forceinline__ void put4(float4& v, int j, float x) {
    if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x;
}

// TV/TA/TC != 0 fix version / boxes per cell / classes at compile time (index arithmetic folds,
// divisions become multiplies); 0 keeps them as run-time values from the geometry.
// VEC: y and dy are 16-byte aligned (float4 accesses); otherwise the same code with scalar accesses.
template <bool WRITE_DY, bool VEC, int TV, int TA, int TC>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) yh_train_kernel(const TrainParams p) {
    extern __shared__ __align__(128) float s_tile[];  // the tile's slice of y
    __shared__ __align__(16) int4 s_win[3 * kWindowMax];  // window of ground-truth records (speculative, or exact after a miss)
    __shared__ __align__(8) uint64_t s_bar[kChunks + 3];  // one mbarrier per chunk, + the window's, + the private cell copies', + the exact window's
    __shared__ float red[kWarps * 6];
    __shared__ int s_rjj[kWindowMax];   // the tile's records in CSR order: index into gt ...
    __shared__ int s_rlc[kWindowMax];   // ... and tile-local cell
    __shared__ int s_w0;                // first record of the window in s_win
    __shared__ int s_covered;           // the window holds all records of the tile's images
    __shared__ __align__(16) float s_patch[kSlots * kPatchFloats];  // record gradients waiting for the dense pass
    __shared__ float s_pdense[kSlots];                            // dense objectness value of a patched row
    __shared__ int s_pr[kSlots];                                  // responsible anchor of a patch
    __shared__ int s_nrec, s_npatch;  // records listed (-1: list incomplete, scan gt instead) / patched
    __shared__ int s_ncell;           // records of the list with a private cell copy
    __shared__ int s_collide;         // bit i: patch i shares its cell with another patch (applied in order)
    __shared__ int s_ready;           // tile index + 1 once the record list of that tile is published

    const YhGeom& g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int version = TV ? TV : g.version;
    const int A = TA ? TA : g.a, C = TC ? TC : g.c;
    const int bs = version == 2 ? 5 + C : 5;
    const int cf = version == 2 ? A * (5 + C) : 5 * A + C;
    const int cells = g.cells;
    const int R = p.tile_cells;
    const int nimg = g.n;
    const bool record_warp = warp == kDenseWarps;

    XT_DECL;
    XT(0);
    if (tid == kDense) {
#pragma unroll
        for (int c = 0; c <= kChunks + 2; ++c) yh_mbar_init(&s_bar[c], 1);
        yh_mbar_fence_init();
        s_ready = 0;
    }
    // everything above overlaps the tail of the previous kernel of the stream (programmatic dependent
    // launch); nothing below may run before that kernel has completed: it may have produced y, and it
    // may be the previous launch of this kernel, which shares the workspace
    if (!p.late_wait) yh_grid_dependency_wait();
    yh_grid_launch_dependents();
    __syncthreads();

    WarpSums sums = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float my_pw = lane < 5 * A ? g.pw[lane / 5] : 0.f;  // anchor multipliers of lane / 5
    const float my_ph = lane < 5 * A ? g.ph[lane / 5] : 0.f;
    const bool multi_tile = p.num_tiles > (int)gridDim.x;
    // records can be processed during the dense pass when their 5+C gradients fit a patch row
    const bool patchable = WRITE_DY && 5 + C <= kPatchFloats;
    uint32_t phase = 0;  // parity of the mbarriers: every barrier completes once per tile
    int pre0 = -1, pre1 = -1;  // record warp: exact record range of this CTA's next tile, once known

    // ---- the six partial sums of the CTA -> global totals -> (last CTA) terms and loss.
    // Every lane carries partial sums (dense rows; record channels by lane role): the warps fold
    // theirs into shared memory, ONE thread folds the warps in a fixed order and adds the CTA's sums
    // onto six 64-bit fixed-point accumulators with integer atomics (exact and order-independent, so
    // the totals are deterministic), then takes a ticket; the last ticket holder reads the totals
    // back and writes terms and loss.  The publishing thread is one that never stores to dy, so its
    // fence only waits for its own six atomics, not for the CTA's gradient stores to drain.
    constexpr int kPublisher = kThreads - 1;
    auto write_warp_sums = [&]() {
        const float s_no = yh_warp_sum(sums.no);
        const float s_xy = yh_warp_sum(sums.xy), s_wh = yh_warp_sum(sums.wh), s_conf = yh_warp_sum(sums.conf);
        const float s_nr = yh_warp_sum(sums.nr), s_cls = yh_warp_sum(sums.cls);
        if (lane == 0) {
            float* r = red + warp * 6;
            r[0] = s_xy; r[1] = s_wh; r[2] = s_conf; r[3] = s_no; r[4] = s_nr; r[5] = s_cls;
        }
    };
    auto publish = [&]() {
        constexpr double kFix = 4294967296.0;  // 2^32: 2.3e-10 resolution, totals up to 2^31
        XT(4);
        unsigned flags = 0u;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            float a = 0.f;
            for (int w = 0; w < kWarps; ++w) a += red[w * 6 + q];
            if (a >= 0.f && a < 1073741824.f) atomicAdd(p.acc + q, (unsigned long long)((double)a * kFix + 0.5));
            else flags |= 1u << q;  // NaN / inf / out of range: the term becomes NaN, like the reference's
        }
        if (flags) atomicOr(p.acc + 6, (unsigned long long)flags);
        XT(5);
#ifndef YH_X_ONE_KERNEL
        // The totals are turned into terms and loss by yh_train_finalize_kernel, a one-warp programmatic
        // dependent of this kernel: it is resident long before this kernel ends and runs the moment it has
        // completed.  A ticket here (release, round trip, read-back by the last CTA: three dependent L2
        // round trips under the draining gradient stores, ~2.3 us) would sit on the kernel's critical path.
        return;
#endif
        // release (this thread's atomics above are performed before the ticket is visible) and acquire
        // (the last ticket holder sees every CTA's sums) in ONE acq_rel atomic: no sequentially
        // consistent fence, which costs about a microsecond while the gradient stores drain
        unsigned long long tk;
        asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(tk) : "l"(p.acc + 7), "l"(1ull) : "memory");
        XT(6);
        if (tk == gridDim.x - 1) {
            double tot[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) tot[q] = (double)__ldcg(p.acc + q) / kFix;
            const unsigned bad = (unsigned)__ldcg(p.acc + 6);
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            const double t0 = (bad & 1u) ? nan : tot[0] * p.inv_den[0];
            const double t1 = (bad & 2u) ? nan : tot[1] * p.inv_den[1];
            const double t2 = (bad & 4u) ? nan : tot[2] * p.inv_den[2];
            const double t3 = (bad & 24u) ? nan : (tot[3] - tot[4]) * p.inv_den[3];
            const double t4 = (bad & 32u) ? nan : tot[5] * p.inv_den[4];
            p.terms[0] = (float)t0; p.terms[1] = (float)t1; p.terms[2] = (float)t2;
            p.terms[3] = (float)t3; p.terms[4] = (float)t4;
            p.loss[0] = (float)(p.lam[0] * t0 + p.lam[1] * t1 + p.lam[2] * t2 + p.lam[3] * t3 + p.lam[4] * t4);
#pragma unroll
            for (int q = 0; q < 8; ++q) p.acc[q] = 0ull;  // ready for the next launch
            XT(7);
        }
    };
    bool sums_final = false;

    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, phase ^= 1u) {
        // everything is 32-bit: train_impl checks that the tensor has < 2^31 floats
#ifdef YH_X_TRACE
        const int tk_ = (t - (int)blockIdx.x) / (int)gridDim.x;
        if (threadIdx.x == 0 && tk_ == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); g_tiletrace[blockIdx.x * 32] = sm_ + 1; }
        XTT(tk_, 0);
#endif
        const int c0 = t * R;
        const int nc = min(R, p.total_cells - c0);
        const int nfl = nc * cf;
        const int nf4 = nfl >> 2;  // whole float4s; a tail (< 4 floats) exists only at the very end of y
        const float* yt = p.y + (size_t)c0 * cf;
        float* dt = p.dy + (size_t)c0 * cf;
        const int n0 = c0 / cells;
        const int cells_left = cells - (c0 - n0 * cells);  // cells of image n0 from c0 on
        const int n_hi = (c0 + nc - 1) / cells;
        const int ch4 = (nf4 + kChunks - 1) / kChunks;     // float4s per chunk

        // record window of this tile: speculative for a CTA's first tile; for its later tiles the exact
        // record range was read while the previous tile was being processed (multi-tile launches)
        int w0 = (int)(p.rec_per_cell * (float)(c0 + (nc >> 1))) - (kWindow >> 1);
        w0 = max(0, min(w0, p.m_local - kWindow));
        int wn = min(kWindow, p.m_local - w0);
        if (pre0 >= 0 && pre1 - pre0 <= kWindowMax) {
            w0 = pre0;
            wn = pre1 - pre0;
        }
        pre0 = -1;
        if (record_warp && t + (int)gridDim.x < p.num_tiles) {  // (loads that nobody waits for during this tile)
            const int c0n = (t + (int)gridDim.x) * R;
            const int ncn = min(R, p.total_cells - c0n);
            pre0 = __ldg(p.gt_off + c0n / cells);
            pre1 = min(__ldg(p.gt_off + (c0n + ncn - 1) / cells + 1), p.m_local);
        }

        auto issue_chunk = [&](int c) {  // one thread: chunk c of the tile -> shared memory
            const int lo = c * ch4, n4 = min(ch4, nf4 - lo);
            if (VEC && n4 > 0) {
                yh_mbar_expect_tx(&s_bar[c], (uint32_t)n4 * 16u);
                yh_bulk_load(s_tile + 4 * lo, yt + 4 * lo, (uint32_t)n4 * 16u, &s_bar[c]);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&s_bar[c])) : "memory");
            }
        };
        if (tid == kDense) {
            // TMA producer: the window first (everything the record warp does hangs on it), then the
            // first chunks of the tile
            yh_fence_proxy_async();  // (multi-tile: the stage was read through the generic proxy)
            if (wn > 0) {
                yh_mbar_expect_tx(&s_bar[kChunks], (uint32_t)wn * 48u);
                yh_bulk_load(s_win, p.gt + w0, (uint32_t)wn * 48u, &s_bar[kChunks]);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&s_bar[kChunks])) : "memory");
            }
#pragma unroll
            for (int c = 0; c < kChunks && c < kAhead; ++c) issue_chunk(c);
            // multi-tile launches: the stage is single-buffered, so the next tile's loads can only
            // start once this tile is done; pulling the next tile into L2 now keeps DRAM busy through
            // this tile's tail and turns that later round trip into an L2 hit
            const int tn_ = t + (int)gridDim.x;
            if (VEC && tn_ < p.num_tiles) {
                const int c0n = tn_ * R;
                const int nfn = (min(R, p.total_cells - c0n) * cf) & ~3;
                if (nfn > 0) yh_bulk_prefetch_l2(p.y + (size_t)c0n * cf, (uint32_t)nfn * 4u);
            }
        }

        // CSR offsets of the tile's images (every warp: the dense pass needs the box counts)
#ifdef YH_X_NOMETA
        const int o0 = 0, o1 = 3, o2 = 6, o3 = 0;
#else
        const int o0 = __ldg(p.gt_off + n0), o1 = __ldg(p.gt_off + n0 + 1);
        const int o2 = __ldg(p.gt_off + min(n0 + 2, nimg));
        const int o3 = (record_warp && n_hi > n0 + 1) ? __ldg(p.gt_off + n_hi + 1) : 0;
#endif
        const int thr0 = cells_left * cf;           // tile-local float index where image n0 + 1 starts
        const bool two_img = R <= cells;            // a tile then touches at most two images
        const float kn0 = (float)(o1 - o0), kn1 = (float)(o2 - o1);
        // boxes in the image of tile-local float index f (the no-object term counts once per box)
        auto kn_of = [&](int f) -> float {
            if (f < thr0) return kn0;
            if (two_img) return kn1;
            const int n = min(n0 + 1 + (f - thr0) / (cells * cf), nimg - 1);  // (clamped: idle lanes pass garbage)
            return (float)(__ldg(p.gt_off + n + 1) - __ldg(p.gt_off + n));
        };
        // record header -> tile-local cell if the record lies in this tile, else -1
        auto local_cell = [&](const int4& h) -> int {
            const bool ok = h.y >= 0 && h.y < g.s_h && h.z >= 0 && h.z < g.s_w && h.x >= 0 && h.x < nimg;
            const int lc = ok ? h.x * cells + h.y * g.s_w + h.z - c0 : -1;
            return (lc >= 0 && lc < nc) ? lc : -1;
        };
        if (!VEC) {  // unaligned y: the streaming warps copy the tile themselves
            for (int i = tid; i < nfl; i += kThreads) s_tile[i] = __ldg(yt + i);
            __syncthreads();
        } else if (tid < (nfl & 3)) {
            s_tile[4 * nf4 + tid] = __ldg(yt + 4 * nf4 + tid);  // floats past the last whole float4 of the tensor
        }

        // Patchable records are dealt round-robin to the warps (a record is ~2700 cycles of one warp's
        // dependent arithmetic); warp 0 is left out, it keeps the tile's chunks flowing.  Each record
        // goes into its own patch slot, so the order in which they are processed does not matter; the
        // patches are applied in list (CSR) order afterwards.  The first cell_slots records of the list
        // read their cell from a private early copy (see the record warp), the others from the tile
        // once their chunk has landed.  try_records(c, block) processes what is ready: records with a
        // private copy once those copies have landed (block: wait for them), the others if their cell
        // lies in chunks <= c.  Returns false if the record warp has not published the list yet.
        float* s_cell = s_tile + (((size_t)R * cf + 3) & ~(size_t)3) + 4;  // private cell copies, after the tile
        unsigned rec_done = 0u;
        int my_npatch = -1, my_ncell = 0, my_w0 = 0;
        bool cells_in = false;
        auto try_records = [&](int upto_chunk, bool block) -> bool {
            if (warp == 0) return true;
            if (my_npatch < 0) {
                if (*reinterpret_cast<volatile int*>(&s_ready) != t + 1) return false;
                __threadfence_block();
                yh_mbar_wait(&s_bar[kChunks], phase);      // (both complete by now: they make the window,
                yh_mbar_wait(&s_bar[kChunks + 2], phase);  //  speculative or exact, visible to this warp)
                my_npatch = s_npatch;
                my_ncell = s_ncell;
                my_w0 = s_w0;
            }
            if (!cells_in && my_ncell > 0) {
                if (block) { yh_mbar_wait(&s_bar[kChunks + 1], phase); cells_in = true; }
                else cells_in = yh_mbar_test(&s_bar[kChunks + 1], phase);
            }
            for (int i = warp - 1, k = 0; i < my_npatch; i += kWarps - 1, ++k) {
                if ((rec_done >> k) & 1u) continue;
                const int jj = s_rjj[i], lcell = s_rlc[i];
                const float* ycell;
                if (i < my_ncell) {
                    if (!cells_in) continue;
                    ycell = s_cell + (size_t)i * p.cell_slot_floats + (int)(((long long)(c0 + lcell) * cf) & 3);
                } else {
                    const int cb = min(((lcell + 1) * cf - 1) / (4 * ch4), kChunks - 1);  // chunk of the cell's last float
                    if (cb > upto_chunk) continue;
                    ycell = s_tile + lcell * cf;
                }
                rec_done |= 1u << k;
                const int4* rp = s_win + 3 * (jj - my_w0);
                RecordRegs rr;
                rr.hd = rp[0];
                rr.tt = *reinterpret_cast<const float4*>(rp + 1);
                rr.bb = *reinterpret_cast<const float4*>(rp + 2);
#ifdef YH_X_TRACE
                const long long rc0 = clock64();
#endif
                const int r = process_record<2>(p, version, A, C, rr, jj, ycell, nullptr, false,
                                                s_patch + i * kPatchFloats, s_pdense + i, kn_of(lcell * cf), lane,
                                                my_pw, my_ph, sums);
                if (lane == 0) s_pr[i] = r;
#ifdef YH_X_TRACE
                if (lane == 0) g_rcycles[blockIdx.x * 16 + warp] = (unsigned int)(clock64() - rc0);
#endif
            }
            return true;
        };

        if (!record_warp) {
            // ================= streaming warps: the dense pass =================
            const float4* tile4 = reinterpret_cast<const float4*>(s_tile);
            for (int c = 0; c < kChunks; ++c) {
                const int lo = c * ch4, hi = min(lo + ch4, nf4);
                yh_mbar_wait(&s_bar[c], phase);  // (empty chunks complete at once: every barrier flips once per tile)
                if (tid == 0 && c + kAhead < kChunks) issue_chunk(c + kAhead);  // keep kAhead chunks in flight
                if (version == 2 && two_img) {
                    // The common case, kept lean (the kernel's issue slots are shared with the tail of the kernel
                    // next to it): floats lf..lf+3 sit at positions m..m+3 of a predictor row, the objectness logit
                    // (position 4) is among them iff 1 <= m <= 4; m advances by (4 * kDense) mod (5 + C) per trip
                    // instead of being divided out; a tile touches at most two images.  Branch-free.
                    const int mstep = (4 * kDense) % bs;
                    int m = (4 * (lo + tid)) % bs;
                    for (int i4 = lo + tid; i4 < hi; i4 += kDense) {
                        const float4 v = tile4[i4];
                        const int lf = 4 * i4;
                        const bool has = (unsigned)(m - 1) < 4u;
                        float tl = v.w;
                        tl = m == 2 ? v.z : tl;
                        tl = m == 3 ? v.y : tl;
                        tl = m == 4 ? v.x : tl;
                        float conf;
                        float w = noobj_term(tl, (lf + 4 - m) < thr0 ? kn0 : kn1, &conf);
                        w = has ? w : 0.f;
                        sums.no += w;
                        const float val = noobj_grad(w, conf, p.cno);
                        float4 o;
                        o.x = m == 4 ? val : 0.f;
                        o.y = m == 3 ? val : 0.f;
                        o.z = m == 2 ? val : 0.f;
                        o.w = m == 1 ? val : 0.f;
                        if (WRITE_DY) store4<VEC>(dt, i4, o);
                        m += mstep;
                        m = m >= bs ? m - bs : m;
                    }
                } else
                for (int i4 = lo + tid; i4 < hi; i4 += kDense) {
                    const float4 v = tile4[i4];
                    const int lf = 4 * i4;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (version == 2) {
                        // (tiles that span more than two images: the box count is looked up per float4)
                        const int m = lf % bs;
                        const bool has = (unsigned)(m - 1) < 4u;
                        float tl = v.w;
                        tl = m == 2 ? v.z : tl;
                        tl = m == 3 ? v.y : tl;
                        tl = m == 4 ? v.x : tl;
                        float conf;
                        float w = noobj_term(tl, kn_of(lf + 4 - m), &conf);
                        w = has ? w : 0.f;
                        sums.no += w;
                        const float val = noobj_grad(w, conf, p.cno);
                        o.x = m == 4 ? val : 0.f;
                        o.y = m == 3 ? val : 0.f;
                        o.z = m == 2 ? val : 0.f;
                        o.w = m == 1 ? val : 0.f;
                    } else {
                        // v1 cell: B blocks of (tx,ty,tw,th,to), then C class logits
                        const int pc = lf % cf;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            int qq = pc + j;
                            if (qq >= cf) qq -= cf;
                            if (qq < 5 * A && qq % 5 == 4) {
                                float conf;
                                const float w = noobj_term(pick4(v, j), kn_of(lf + j), &conf);
                                sums.no += w;
                                put4(o, j, noobj_grad(w, conf, p.cno));
                            }
                        }
                    }
                    if (WRITE_DY) store4<VEC>(dt, i4, o);
                }
                try_records(c, false);
            }
            while (!try_records(kChunks - 1, true)) { }  // (the record warp publishes after one round trip)
            if (tid < (nfl & 3)) {  // floats past the last whole float4 of the tensor
                const int lf = 4 * nf4 + tid;
                const int pc = lf % cf;
                const bool is_to = version == 2 ? (pc % bs == 4) : (pc < 5 * A && pc % 5 == 4);
                float o = 0.f;
                if (is_to) {
                    float conf;
                    const float w = noobj_term(s_tile[lf], kn_of(lf), &conf);
                    sums.no += w;
                    o = noobj_grad(w, conf, p.cno);
                }
                if (WRITE_DY) dt[lf] = o;
            }
        } else {
            // ================= record warp: the tile's ground-truth records =================
            const int rec0 = o0;
            int rec1 = min(n_hi == n0 ? o1 : (n_hi == n0 + 1 ? o2 : o3), p.m_local);
#if defined(YH_X_NOSPARSE) || defined(YH_X_NOMETA)
            rec1 = rec0;
#endif
            XT(4);  // record warp: offsets arrived
            // the tile's records, in CSR order, into the shared list
            int n = 0;
            auto append = [&](int jj, int lc) {
                const unsigned bal = __ballot_sync(0xffffffffu, lc >= 0);
                if (lc >= 0) {
                    const int idx = n + __popc(bal & ((1u << lane) - 1u));
                    if (idx < kWindowMax) { s_rjj[idx] = jj; s_rlc[idx] = lc; }
                }
                n += __popc(bal);
            };
            bool covered = rec0 >= w0 && rec1 <= w0 + wn;  // the window has them all
            yh_mbar_wait(&s_bar[kChunks], phase);  // (always: the barrier's phase must be consumed)
            int wbase = w0;
            // The guess missed (box counts far from uniform): one more round trip fetches the exact
            // window -- the offsets are known now -- instead of sending the tile's records down the
            // slow path after the dense pass.  (The barrier flips once per tile either way.)
            if (!covered && rec1 - rec0 <= kWindowMax && rec1 > rec0) {
                if (lane == 0) {
                    yh_fence_proxy_async();
                    yh_mbar_expect_tx(&s_bar[kChunks + 2], (uint32_t)(rec1 - rec0) * 48u);
                    yh_bulk_load(s_win, p.gt + rec0, (uint32_t)(rec1 - rec0) * 48u, &s_bar[kChunks + 2]);
                }
                yh_mbar_wait(&s_bar[kChunks + 2], phase);
                wbase = rec0;
                covered = true;
            } else if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&s_bar[kChunks + 2])) : "memory");
            }
            if (lane == 0) { s_w0 = wbase; s_covered = covered ? 1 : 0; }
            if (covered) {
                for (int base = rec0; base < rec1; base += 32) {
                    const int jj = base + lane;
                    append(jj, jj < rec1 ? local_cell(s_win[3 * (jj - wbase)]) : -1);
                }
            } else {
                for (int base = rec0; base < rec1; base += 32) {
                    const int jj = base + lane;
                    append(jj, jj < rec1 ? local_cell(__ldg(reinterpret_cast<const int4*>(p.gt + jj))) : -1);
                }
            }
            __syncwarp();
            XT(5);  // record warp: list built
            const bool complete = n <= kWindowMax;
            // Tiles dense with records (BASELINE config 5: ~17 per tile) gain nothing from hiding them behind
            // the stream -- the streaming warps would spend more time on records than on the stream -- so
            // beyond kShadowMax records they are all processed after the dense pass, by all warps at once.
            const int n_patch = (patchable && complete && covered && n <= kShadowMax) ? min(n, kSlots) : 0;
            // Private early copies of the first records' cells: the tile's last chunk lands when the
            // stream ends, and a record found in it would add its ~1.4 us to the CTA's critical path;
            // a 16-byte aligned bulk copy of just the cell, asked for now, is back long before that.
            const int n_cell = VEC ? min(n_patch, p.cell_slots) : 0;
            uint32_t cbytes = 0;
            if (lane < n_cell) {
                const long long f = (long long)(c0 + s_rlc[lane]) * cf;  // first float of the cell in y
                const int shift = (int)(f & 3);
                const int win = (shift + cf + 3) & ~3;
                if (f - shift + win <= (p.total_floats & ~3ll)) {
                    cbytes = (uint32_t)win * 4u;
                    yh_bulk_load(s_cell + (size_t)lane * p.cell_slot_floats, p.y + (f - shift), cbytes, &s_bar[kChunks + 1]);
                } else {  // the very last cells of the tensor: plain loads
                    for (int q = 0; q < cf; ++q) s_cell[(size_t)lane * p.cell_slot_floats + shift + q] = __ldg(p.y + f + q);
                }
            }
            for (int o = 16; o > 0; o >>= 1) cbytes += __shfl_xor_sync(0xffffffffu, cbytes, o);
            if (lane == 0) {
                if (cbytes) yh_mbar_expect_tx(&s_bar[kChunks + 1], cbytes);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&s_bar[kChunks + 1])) : "memory");
                s_nrec = complete ? n : -1;
                s_npatch = n_patch;
                s_ncell = n_cell;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *reinterpret_cast<volatile int*>(&s_ready) = t + 1;  // publish the list
            __syncwarp();
            {   // which patched records share their cell with another one?  Those patches are applied one by
                // one, in order, by this warp; all others by the warps that produced them
                bool shared_cell = false;
                if (lane < n_patch)
                    for (int j = 0; j < n_patch; ++j) shared_cell = shared_cell || (j != lane && s_rlc[j] == s_rlc[lane]);
                const unsigned bal = __ballot_sync(0xffffffffu, shared_cell);
                if (lane == 0) s_collide = (int)bal;  // bit i: patch i shares its cell  (kSlots <= 32)
            }
            for (int c = 0; c < kChunks; ++c) {  // this warp's share, as copies and chunks land
                try_records(c - 1, false);
                yh_mbar_wait(&s_bar[c], phase);
            }
            try_records(kChunks - 1, true);
            XT(6);  // record warp: records processed
        }

        // ---------------- the records' gradients go onto the dense rows ----------------
        const bool last_tile = t + (int)gridDim.x >= p.num_tiles;
        if (last_tile) write_warp_sums();  // (final unless records are left over, see below)
        XT(1);
        __syncthreads();  // the tile's dense dL/dy is written (and visible to the whole CTA); patches are ready
        XT(2);
        XTT(tk_, 1);
        const int nrec = s_nrec, npatch = s_npatch;
        // (no records left over: the sums folded before the barrier are final and get published
        //  right after the patch stores, without another CTA barrier)
        if (last_tile && !(nrec < 0 || nrec > npatch)) sums_final = true;
        {
            const unsigned shared_mask = (unsigned)s_collide;
            // every warp applies the patches it produced: a plain overwrite of the 5+C floats of the row
            // (the dense pass' values there are zero but the objectness channel, kept in s_pdense)
            for (int i = warp - 1; i < npatch && warp > 0; i += kWarps - 1) {
                if ((shared_mask >> i) & 1u) continue;
                const int lcell = s_rlc[i], r = s_pr[i];
                if (lane < 5 + C) {
                    float* addr = dt + lcell * cf + (version == 2 ? r * bs + lane : (lane < 5 ? r * 5 + lane : 5 * A + lane - 5));
                    float val = s_patch[i * kPatchFloats + lane];
                    if (lane == 4) val = __fadd_rn(s_pdense[i], val);
                    *addr = val;
                }
            }
            // patches of cells that carry several records: in CSR order by one warp; the first one of a
            // cell overwrites, the later ones read the row back
            if (record_warp) {
                for (unsigned todo = shared_mask; todo; todo &= todo - 1u) {
                    const int i = __ffs(todo) - 1;
                    const int lcell = s_rlc[i], r = s_pr[i];
                    const bool again = __any_sync(0xffffffffu, lane < i && s_rlc[lane] == lcell);
                    if (lane < 5 + C) {
                        float* addr = dt + lcell * cf + (version == 2 ? r * bs + lane : (lane < 5 ? r * 5 + lane : 5 * A + lane - 5));
                        float val = s_patch[i * kPatchFloats + lane];
                        if (again) val = __fadd_rn(*addr, val);
                        else if (lane == 4) val = __fadd_rn(s_pdense[i], val);
                        *addr = val;
                    }
                    __syncwarp();
                }
            }
        }
        if (nrec < 0 || nrec > npatch) {
            // records that were not patched during the dense pass (window miss, more than kSlots in
            // the tile, wide class rows, loss-only call): now, by all warps, dealt by cell so that
            // records of one cell stay on one warp in CSR order; read-modify-write on top of dense
            // values and patches.  The tile is still in shared memory.
            if (npatch > 0) __syncthreads();
            const bool track = npatch == 0 && R <= 32 * kWarps;  // cells this warp has updated fit one mask
            unsigned seen = 0u;
            const bool in_window = s_covered != 0;  // the listed records sit in s_win, from s_w0 on
            const int win0 = s_w0;
            if (in_window) {
                yh_mbar_wait(&s_bar[kChunks], phase);      // (both complete: make the window visible
                yh_mbar_wait(&s_bar[kChunks + 2], phase);  //  to every warp)
            }
            const int rec0 = o0;
            const int lim = nrec >= 0 ? nrec : min(__ldg(p.gt_off + n_hi + 1), p.m_local);
            for (int base = nrec >= 0 ? npatch : rec0; base < lim; base += 32) {
                int lc = -1, jj = 0;
                if (base + lane < lim) {
                    if (nrec >= 0) { jj = s_rjj[base + lane]; lc = s_rlc[base + lane]; }
                    else { jj = base + lane; lc = local_cell(__ldg(reinterpret_cast<const int4*>(p.gt + jj))); }
                    if (lc % kWarps != warp) lc = -1;
                }
                unsigned bal = __ballot_sync(0xffffffffu, lc >= 0);
                while (bal) {
                    const int b = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int j = __shfl_sync(0xffffffffu, jj, b);
                    const int lcell = __shfl_sync(0xffffffffu, lc, b);
                    RecordRegs rr;
                    if (nrec >= 0 && in_window) {  // the record window in shared memory has it
                        const int4* rp = s_win + 3 * (j - win0);
                        rr.hd = rp[0];
                        rr.tt = *reinterpret_cast<const float4*>(rp + 1);
                        rr.bb = *reinterpret_cast<const float4*>(rp + 2);
                    } else {
                        const int4* rp = reinterpret_cast<const int4*>(p.gt + j);
                        rr.hd = __ldg(rp);
                        rr.tt = __ldg(reinterpret_cast<const float4*>(rp + 1));
                        rr.bb = __ldg(reinterpret_cast<const float4*>(rp + 2));
                    }
                    // first record of this cell on this warp (records are dealt by cell): its row still
                    // holds the dense pass' values -- unless patches were applied before
                    const unsigned bit = 1u << ((lcell / kWarps) & 31);
                    const bool again = !track || (seen & bit);
                    seen |= bit;
                    process_record<WRITE_DY ? 1 : 0>(p, version, A, C, rr, j, s_tile + lcell * cf, dt + lcell * cf, again,
                                                     nullptr, nullptr, kn_of(lcell * cf), lane, my_pw, my_ph, sums);
                }
            }
        }
        if (multi_tile) __syncthreads();  // the stage and the lists are rewritten by the next tile
        XTT(tk_, 2);
    }
    XT(3);

#ifdef YH_X_NOREDUCE
    if (sums.no == 123.456f) p.terms[0] = sums.no + sums.xy + sums.cls;
    return;
#endif
    if (!sums_final) {  // (records were left over in the last tile)
        write_warp_sums();
        __syncthreads();
    }
    if (tid == kPublisher) {
        publish();
        // (*_overlapped) this kernel ran next to the tails of the kernels in front of it and must not complete
        // before they have (work launched after it could overtake them): ONE thread of the grid waits for them --
        // the grid is not complete until it exits, every other CTA leaves and frees its SM.  The workspace is this
        // call's own (THE OVERLAP CONTRACT, include/yolohead.h), so the sums need not wait.
        if (p.late_wait && blockIdx.x == 0) yh_grid_dependency_wait();
    }
    XT_FLUSH;
}
#ifdef YH_X_TRACE
extern "C" YH_API int yh_x_trace_copy(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_xtrace, (size_t)n * 8);
}
extern "C" YH_API int yh_x_tiletrace_copy(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_tiletrace, (size_t)n * 8);
}
extern "C" YH_API int yh_x_rcycles_copy(unsigned int* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_rcycles, (size_t)n * 4);
}
#endif

// Totals -> terms and loss (reference models/yolov2.py:1132-1138; sharded batches: summed over all ranks through
// peer memory first, yh_finalize.cuh), and the workspace back to zero for the next launch.  Launched as a
// programmatic dependent right behind yh_train_kernel (and behind the fused step's kernel, yh_nms.cu).
__global__ void __launch_bounds__(32) yh_train_finalize_kernel(const YhFinalParams f) {
    yh_grid_launch_dependents();
    yh_grid_dependency_wait();  // the train kernel has completed and its sums are visible
    yh_finalize_warp(f, (int)threadIdx.x);
}


template <bool WDY, bool VEC, int TV, int TA, int TC>
int launch_variant(const TrainParams& p, int grid, cudaStream_t stream) {
    const size_t smem = (((size_t)p.tile_cells * p.g.cell_floats + 3) & ~(size_t)3) * 4 + 16 +
                        (size_t)p.cell_slots * p.cell_slot_floats * 4;
    static size_t configured[64] = {0};  // per device: opt in to > 48 KB of dynamic shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_train_kernel<WDY, VEC, TV, TA, TC>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(train)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    return yh_check_cuda(yh_launch_pdl(yh_train_kernel<WDY, VEC, TV, TA, TC>, dim3(grid), dim3(kThreads), smem, stream, p),
                         "yh_train launch");
}

template <bool WDY, bool VEC>
int launch_geometry(const TrainParams& p, int grid, cudaStream_t stream) {
    const YhGeom& g = p.g;
    // compile-time geometries for the shapes the reference trains: YOLOv2 5 anchors x 20 classes
    // (VOC, models/yolov2.py:49-70) and YOLOv1 B=2, C=20 (config.py:7-11); anything else runs the
    // same kernel with run-time geometry
    if (g.version == 2 && g.a == 5 && g.c == 20) return launch_variant<WDY, VEC, 2, 5, 20>(p, grid, stream);
    if (g.version == 1 && g.a == 2 && g.c == 20) return launch_variant<WDY, VEC, 1, 2, 20>(p, grid, stream);
    return launch_variant<WDY, VEC, 0, 0, 0>(p, grid, stream);
}

void tiling(long long total_cells, int cf, int* tile_cells, int* num_tiles, int* grid) {
    // tiles: whole quads of cells (16-byte aligned starts), at most kTileBytesMax bytes, and a tile
    // count that fills the resident CTA slots evenly (one tile per CTA when the tensor is small)
    int slots = yh_sm_count() * kCtasPerSm;
    if (slots > kMaxGrid) slots = kMaxGrid;
    const long long quads_total = (total_cells + 3) / 4;
    long long max_q = kTileBytesMax / (16ll * cf);
    if (max_q < 1) max_q = 1;
    long long rounds = (quads_total + slots * max_q - 1) / (slots * max_q);  // tiles per CTA
    if (rounds < 1) rounds = 1;
    long long qpt = (quads_total + slots * rounds - 1) / (slots * rounds);    // quads per tile
    if (qpt < 1) qpt = 1;
    const long long tiles = (quads_total + qpt - 1) / qpt;
    *tile_cells = (int)(qpt * 4);
    *num_tiles = (int)tiles;
    *grid = (int)(tiles < slots ? tiles : slots);
}

}  // namespace

int yh_train_tiling(long long total_cells, int cell_floats, int* tile_cells, int* num_tiles, int* grid) {
    if (total_cells <= 0 || cell_floats <= 0 || kTileBytesMax / (16ll * cell_floats) < 1) return YH_ERR_UNSUPPORTED;
    tiling(total_cells, cell_floats, tile_cells, num_tiles, grid);
    return YH_OK;
}

int yh_launch_finalize(const YhFinalParams& f, cudaStream_t stream) {
#ifdef YH_X_ONE_KERNEL
    return 0;
#else
    return yh_check_cuda(yh_launch_pdl(yh_train_finalize_kernel, dim3(1), dim3(32), 0, stream, f), "yh_train_finalize launch");
#endif
}

// Loss normalisation shared by the train head and the fused step: denominators of the five means and the
// gradient coefficients.  d(mean)/d(activation):  xy: 2(s-t)/(2M);  wh: 2(q-T)/(2M) * dq/dt with dq/dt = q/2;
// conf: 2(conf-iou)/M;  noobj: 2 conf /(M(P-1));  cls: 2/M
void yh_loss_coefs(const float* lambdas_host, int m_global, int preds, YhLossCoef* k, YhFinalParams* f) {
    const double M = (double)m_global;
    const double P1 = (double)preds - 1.0;
    k->cxy = (float)(lambdas_host[0] / M);
    k->cwh = (float)(lambdas_host[1] / (2.0 * M));
    k->cconf = (float)(lambdas_host[2] * 2.0 / M);
    k->cno = (float)(lambdas_host[3] * 2.0 / (M * P1));
    k->ccls = (float)(lambdas_host[4] * 2.0 / M);
    for (int i = 0; i < 5; ++i) f->lam[i] = lambdas_host[i];
    f->inv_den[0] = 1.0 / (2.0 * M);
    f->inv_den[1] = 1.0 / (2.0 * M);
    f->inv_den[2] = 1.0 / M;
    f->inv_den[3] = 1.0 / (M * P1);
    f->inv_den[4] = 1.0 / M;
}

int yh_fill_exchange(YhFinalParams* f, const YhExchange* xch_host) {
    f->rank = 0;
    f->world = 1;
    for (int q = 0; q < YH_MAX_RANKS; ++q) f->peer[q] = nullptr;
    if (!xch_host || xch_host->world <= 1) return YH_OK;
    YH_REQUIRE(xch_host->world <= YH_MAX_RANKS && xch_host->rank >= 0 && xch_host->rank < xch_host->world, YH_ERR_INVALID,
               "bad exchange: rank %d of %d (at most %d ranks)", xch_host->rank, xch_host->world, YH_MAX_RANKS);
    for (int q = 0; q < xch_host->world; ++q) {
        YH_REQUIRE(xch_host->slots[q] && ((uintptr_t)xch_host->slots[q] & 7) == 0, YH_ERR_INVALID,
                   "exchange buffer of rank %d is NULL or misaligned", q);
        f->peer[q] = reinterpret_cast<unsigned long long*>(xch_host->slots[q]);
    }
    f->rank = xch_host->rank;
    f->world = xch_host->world;
    return YH_OK;
}

// The train head of every train entry point: the train kernel and, right behind it, the one-warp finalize kernel.
int yh_train_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
                  const float* anchors_wh_host, float img_h, float img_w, const YhGt* gt,
                  const int32_t* gt_off, int m_local, int m_global, const float* lambdas_host,
                  float* dy, float* terms, float* loss, int32_t* resp, float* iou_resp, void* ws,
                  size_t ws_bytes, void* stream, int late_wait, const YhExchange* xch_host) {
    TrainParams p;
    int rc = yh_make_geom(&p.g, version, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w);
    if (rc) return rc;
    YH_REQUIRE(y && gt_off && terms && loss && lambdas_host, YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(m_local >= 0 && (m_local == 0 || gt), YH_ERR_INVALID, "bad ground-truth arguments");
    // torch.stack([]) raises in collate_fn (models/yolov2.py:1538) and the means are over M
    YH_REQUIRE(m_global > 0, YH_ERR_EMPTY, "no ground-truth boxes in the batch (m_global=%d)", m_global);
    YH_REQUIRE(m_local <= m_global, YH_ERR_INVALID, "m_local > m_global");
    YH_REQUIRE(5 * a <= 32, YH_ERR_UNSUPPORTED, "more than 6 anchors/boxes per cell (got %d)", a);
    YH_REQUIRE(ws && ws_bytes >= yh_train_workspace_bytes(), YH_ERR_WORKSPACE,
               "workspace too small: %zu < %zu", ws_bytes, yh_train_workspace_bytes());
    YH_REQUIRE(((uintptr_t)y & 3) == 0 && ((uintptr_t)dy & 3) == 0 && ((uintptr_t)gt & 15) == 0,
               YH_ERR_INVALID, "misaligned pointer (y/dy need 4-byte, gt 16-byte alignment)");

    p.y = y; p.dy = dy; p.gt = gt; p.gt_off = gt_off;
    p.terms = terms; p.loss = loss; p.resp = resp; p.iou_resp = iou_resp;
    YH_REQUIRE(((uintptr_t)ws & 7) == 0, YH_ERR_INVALID, "workspace must be 8-byte aligned");
    p.acc = reinterpret_cast<unsigned long long*>(ws);
    const long long total_cells = (long long)n * p.g.cells;
    const int cf = p.g.cell_floats;
    YH_REQUIRE(total_cells * cf < (1ll << 31), YH_ERR_UNSUPPORTED, "head tensor has 2^31 or more floats");
    p.total_cells = (int)total_cells;
    p.m_local = m_local;
    p.late_wait = late_wait;
    p.rec_per_cell = (float)((double)m_local / (double)total_cells);

    YhLossCoef kc;
    YhFinalParams f;
    yh_loss_coefs(lambdas_host, m_global, p.g.preds, &kc, &f);
    p.cxy = kc.cxy; p.cwh = kc.cwh; p.cconf = kc.cconf; p.cno = kc.cno; p.ccls = kc.ccls;
    for (int i = 0; i < 5; ++i) { p.lam[i] = f.lam[i]; p.inv_den[i] = f.inv_den[i]; }

    YH_REQUIRE(kTileBytesMax / (16ll * cf) >= 1, YH_ERR_UNSUPPORTED, "cell too wide for the shared-memory stage (%d floats per cell)", cf);
    int grid = 0;
    tiling(total_cells, cf, &p.tile_cells, &p.num_tiles, &grid);

    // private cell copies: as many as fit next to kCtasPerSm CTAs' tiles in the SM's shared memory
    p.total_floats = total_cells * cf;
    p.cell_slot_floats = (cf + 6 + 3) & ~3;
    {
        const long long per_cta = (227ll * 1024) / kCtasPerSm - 1024 - 18 * 1024;  // minus reserve and static arrays
        const long long left = per_cta - ((long long)p.tile_cells * cf * 4 + 32);
        long long slots = left > 0 ? left / (p.cell_slot_floats * 4ll) : 0;
        p.cell_slots = (int)(slots < kCellSlots ? slots : kCellSlots);
    }
    f.acc = p.acc; f.terms = p.terms; f.loss = p.loss;
    rc = yh_fill_exchange(&f, xch_host);
    if (rc) return rc;

    const bool vec = ((uintptr_t)y & 15) == 0 && ((uintptr_t)dy & 15) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dy) {
        rc = vec ? launch_geometry<true, true>(p, grid, st) : launch_geometry<true, false>(p, grid, st);
    } else {
        rc = vec ? launch_geometry<false, true>(p, grid, st) : launch_geometry<false, false>(p, grid, st);
    }
    if (rc) return rc;
    return yh_launch_finalize(f, st);
}

extern "C" {

size_t yh_train_workspace_bytes(void) { return 256; }
size_t yh_exchange_bytes(void) { return kYhXchBytes; }

int yh_v2_train(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                float img_h, float img_w, const YhGt* gt, const int32_t* gt_off, int m_local,
                int m_global, const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return yh_train_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, gt, gt_off, m_local,
                         m_global, lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream, 0, nullptr);
}

int yh_v1_train(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss, int32_t* resp,
                float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return yh_train_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, gt, gt_off, m_local, m_global,
                         lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream, 0, nullptr);
}

int yh_v2_train_overlapped(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                           float img_h, float img_w, const YhGt* gt, const int32_t* gt_off, int m_local,
                           int m_global, const float* lambdas_host, float* dy, float* terms, float* loss,
                           int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return yh_train_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, gt, gt_off, m_local,
                         m_global, lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream, 1, nullptr);
}

int yh_v1_train_overlapped(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                           const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                           const float* lambdas_host, float* dy, float* terms, float* loss, int32_t* resp,
                           float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return yh_train_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, gt, gt_off, m_local, m_global,
                         lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream, 1, nullptr);
}

int yh_v2_train_sharded(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                        float img_h, float img_w, const YhGt* gt, const int32_t* gt_off, int m_local,
                        int m_global, const float* lambdas_host, float* dy, float* terms, float* loss,
                        int32_t* resp, float* iou_resp, int flags, const YhExchange* xch_host, void* ws,
                        size_t ws_bytes, void* stream) {
    return yh_train_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, gt, gt_off, m_local,
                         m_global, lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream,
                         (flags & YH_STEP_OVERLAPPED) ? 1 : 0, xch_host);
}

int yh_v1_train_sharded(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                        const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                        const float* lambdas_host, float* dy, float* terms, float* loss, int32_t* resp,
                        float* iou_resp, int flags, const YhExchange* xch_host, void* ws, size_t ws_bytes,
                        void* stream) {
    return yh_train_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, gt, gt_off, m_local, m_global,
                         lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream,
                         (flags & YH_STEP_OVERLAPPED) ? 1 : 0, xch_host);
}

}  // extern "C"
