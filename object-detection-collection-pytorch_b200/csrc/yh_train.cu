// Fused YOLO train head: decode + responsible-predictor assignment + five loss terms + dL/dy
// in one pass over the head tensor, ONE launch per batch.
//
// Replaces YOLOv2.get_loss / YOLOv1.get_loss and their autograd backward
// (reference models/yolov2.py:747-1140, models/yolov1.py:556-931, models/utils.py:5-65).
//
// The kernel is HBM-bound (~60 flop per 200 bytes), so it is organised around data movement:
//   * one persistent CTA per SM owns a contiguous range of grid cells of the flattened batch;
//   * the range is cut into mini-chunks of a few cells (<= ~4 KB, a multiple of 4 cells so that
//     every chunk starts 16-byte aligned) that are dealt round-robin to the CTA's warps;
//   * every warp runs its OWN pipeline with no block-wide barrier: 1-D TMA bulk loads
//     (cp.async.bulk + mbarrier) fill a private ring of input stages three chunks ahead, the
//     warp assembles dL/dy for the chunk in a private output stage and pushes it back with a
//     TMA bulk store.  Every byte of y is read once and every byte of dy written once, fully
//     coalesced, with no zero-fill pass over dy;
//   * the CTA's slice of the CSR offsets and its ground-truth records are staged in shared
//     memory once, so the per-chunk work never waits on global memory;
//   * the partial sums go through a last-block-done reduction in a fixed order (deterministic).
//
// Per mini-chunk a warp does
//   dense pass : one lane per predictor row (v2) / cell (v1): conf = sigmoid(to), the no-object
//                term and its gradient, zeros elsewhere;
//   sparse pass: ground-truth records whose cell lies in the chunk, in CSR order (so collisions
//                on one predictor accumulate deterministically): lanes decode the A boxes of the
//                cell, IoU against the record, shuffle-argmax picks the responsible predictor,
//                lanes then cover the C classes for the softmax/class term.
#include <limits.h>

#include "yh_common.cuh"

namespace {

constexpr int kMaxWarps = 8;
constexpr int kInStages = 3;
constexpr int kOutStages = 2;
constexpr int kStageBytesTarget = 4096;
constexpr int kMaxGrid = 1024;
constexpr int kPartials = 8;     // floats per CTA in the workspace (6 used)
constexpr int kOffCap = 128;     // CSR offsets cached per CTA (images + 1)
constexpr int kGtCap = 384;      // ground-truth records cached per CTA (18 KB)
constexpr int kWinRegs = 5;      // int4 loads per thread of the speculative record window
constexpr int kMaskWords = 16;   // chunk-has-records bitmask (<= 512 mini-chunks per CTA)
constexpr int kClsRegs = 4;      // class logits per lane kept in registers (C <= 128)

// Optional per-warp timeline (scratch/trace_train.cu builds this file with -DYH_TRACE).
#ifdef YH_TRACE
__device__ long long g_trace[1024 * kMaxWarps * 64];
#define YH_TR(slot)                                                                              \
    do {                                                                                         \
        if (lane == 0 && (slot) < 64) g_trace[((size_t)blockIdx.x * kMaxWarps + warp) * 64 + (slot)] = clock64(); \
    } while (0)
#else
#define YH_TR(slot) do { } while (0)
#endif

// cells per mini-chunk for a cell of `cf` floats: whole quads of cells, <= kStageBytesTarget bytes,
// <= 32 cells (the kernel tracks touched cells of a stage in one 32-bit word)
__host__ __device__ constexpr int mini_cells(int cf) {
    return 4 * ((kStageBytesTarget / (16 * cf)) < 1 ? 1 : ((kStageBytesTarget / (16 * cf)) > 8 ? 8 : kStageBytesTarget / (16 * cf)));
}

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
    float* partials;      // [gridDim][kPartials]
    unsigned int* ticket;
    long long total_cells;
    long long quads_total;  // ceil(total_cells / 4)
    int mc;                 // cells per mini-chunk (multiple of 4)
    int m_local;            // records in gt
    int warps_log2;         // log2(warps per CTA)
    int q_base, q_extra;    // quads of cells per CTA: q_base (+1 for the first q_extra CTAs)
    float rec_per_cell;     // m_local / total_cells: where a CTA's records sit if spread evenly
    int tma_in, tma_out;    // base pointers 16-byte aligned
    float lam[5];
    double inv_den[5];      // 1/(2M), 1/(2M), 1/M, 1/(M(P-1)), 1/M
    float cxy, cwh, cconf, cno, ccls;  // gradient coefficients (see train_impl)
};

struct WarpSums {
    float no, xy, wh, conf, nr, cls;
};

// order-preserving float <-> int map (for redux.sync max on floats; NaNs are not ordered)
__device__ __forceinline__ int yh_ordered(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float yh_unordered(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// One ground-truth record against its cell; all 32 lanes cooperate and the work is laid out for
// LATENCY (a record sits on the critical path of the warp that owns its chunk):
//   * lane l < 5A owns ONE activation (anchor l/5, channel l%5), so the 5A exp/sigmoid chains run
//     side by side; four shuffles hand every lane its anchor's box, the IoU is formed, and two
//     redux.sync steps pick the responsible anchor (max IoU, then lowest index: torch's first max);
//   * the five lanes of the responsible anchor then each finish THEIR channel (x, y, w, h, conf:
//     target transform, squared error, gradient, read-modify-write of the output row) in parallel;
//     their squared errors accumulate in per-lane registers by channel role (ChannelSums);
//   * class softmax: lanes stride the classes, max through redux.sync on an order-preserving
//     integer image, then ONE butterfly for (sum e, sum e^2): with p = e / sum e,
//       sum_c (p_c - 1[c=t])^2 = S2 - 2 p_t + 1   and   sum_c (p_c - 1[c=t]) p_c = S2 - p_t,
//     S2 = sum p^2, so no third reduction is needed.
// `cellp` / `ocell` point at the cell's floats in the input / output stage; `my_pw/my_ph` are the
// anchor multipliers of this lane's anchor (lane / 5).  Requires 5A <= 32.
struct RecordRegs {
    int4 hd;    // img, cy, cx, cls
    float4 tt;  // stx, sty, tw, th
    float4 bb;  // x1, y1, x2, y2
};

template <bool WRITE_DY>
__device__ __forceinline__ void process_record(const TrainParams& p, const int version, const int A, const int C,
                                               const RecordRegs& rr, int jj, const float* cellp, float* ocell,
                                               int lane, float my_pw, float my_ph, WarpSums& s) {
    const YhGeom& g = p.g;
    const int bs = version == 2 ? 5 + C : 5;
#ifdef YH_TRACE
    const int warp = threadIdx.x >> 5;
#endif
    YH_TR(50);
    const int4 hd = rr.hd;
    const float4 tt = rr.tt;
    const float4 bb = rr.bb;

    const int a = lane / 5, q = lane - 5 * a;
    const bool mine = lane < 5 * A;
    float act = 0.f;
    if (mine) {
        const float t = cellp[a * bs + q];
        const bool is_exp = version == 2 && (q == 2 || q == 3);
        const float e = expf(is_exp ? t : -t);
        act = is_exp ? e : __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
    }
    YH_TR(51);
    const int l0 = mine ? 5 * a : 0;
    const float bx_s = __shfl_sync(0xffffffffu, act, l0);
    const float by_s = __shfl_sync(0xffffffffu, act, l0 + 1);
    const float bw_a = __shfl_sync(0xffffffffu, act, l0 + 2);
    const float bh_a = __shfl_sync(0xffffffffu, act, l0 + 3);
    int key = INT_MIN;  // order-preserving image of this lane's IoU; NaN (as torch) ranks highest
    float iou = 0.f;
    if (mine) {
        const YhBox pb = yh_decode_box(bx_s, by_s, bw_a, bh_a, my_pw, my_ph, hd.z, hd.y, g.gw, g.gh);
        const YhBox gb{bb.x, bb.y, bb.z, bb.w};
        iou = yh_iou_xyxy(pb, gb);
        key = iou != iou ? INT_MAX : yh_ordered(iou);
    }
    YH_TR(52);
    const int best = __reduce_max_sync(0xffffffffu, key);
    const int r = __reduce_min_sync(0xffffffffu, (mine && key == best) ? a : 1 << 20);  // first max
    const float iou_r = __shfl_sync(0xffffffffu, iou, 5 * r);

    // the five lanes of the responsible anchor finish one channel each
    if (mine && a == r) {
        float d, grad;
        if (q < 2) {            // x, y: (sigmoid(t) - target)^2, models/yolov2.py:1046-1050
            d = act - (q == 0 ? tt.x : tt.y);
            grad = p.cxy * d * act * (1.f - act);
            s.xy += d * d;
        } else if (q < 4) {     // w, h: (sqrt(act) - sqrt(target))^2, models/yolov2.py:946-947, 1063-1067
            const float t = q == 2 ? tt.z : tt.w;
            const float tgt = version == 2 ? __fsqrt_rn(__fdiv_rn(t, q == 2 ? my_pw : my_ph)) : __fsqrt_rn(t);
            const float qv = __fsqrt_rn(act);
            d = qv - tgt;
            grad = p.cwh * d * qv;
            if (version != 2) grad *= 1.f - act;  // v1: d sqrt(sigmoid(t)) / dt, models/yolov1.py:745-761
            s.wh += d * d;
        } else {                // objectness: (iou - conf)^2 and the no-object correction
            d = act - iou_r;
            grad = (p.cconf * d - p.cno * act) * act * (1.f - act);
            s.conf += d * d;
            s.nr += act * act;
            if (p.resp) p.resp[jj] = r;
            if (p.iou_resp) p.iou_resp[jj] = iou_r;
        }
        if (WRITE_DY) ocell[r * bs + q] += grad;
    }
    YH_TR(53);

    // class term
    const int coff = version == 2 ? r * bs + 5 : 5 * A;
    const float* cl = cellp + coff;
    float lg[kClsRegs];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        const int c = lane + 32 * k;
        lg[k] = c < C ? cl[c] : -INFINITY;
        mx = fmaxf(mx, lg[k]);
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) mx = fmaxf(mx, cl[c]);
    mx = yh_unordered(__reduce_max_sync(0xffffffffu, yh_ordered(mx)));
    YH_TR(54);
    float s1 = 0.f, s2 = 0.f, et = 0.f;  // sum e, sum e^2, e of the target class (owning lane only)
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        const int c = lane + 32 * k;
        lg[k] = c < C ? expf(lg[k] - mx) : 0.f;
        s1 += lg[k];
        s2 += lg[k] * lg[k];
        if (c == hd.w) et = lg[k];
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
        const float e = expf(cl[c] - mx);
        s1 += e;
        s2 += e * e;
        if (c == hd.w) et = e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    YH_TR(55);
    const bool has_t = hd.w >= 0 && hd.w < C;
    et = __shfl_sync(0xffffffffu, et, has_t ? (hd.w & 31) : 0);
    const float inv = __fdiv_rn(1.0f, s1);
    const float S2 = s2 * inv * inv;
    const float pt = has_t ? et * inv : 0.f;
    const float dot = S2 - pt;
    if (lane == 0) s.cls += S2 - 2.f * pt + (has_t ? 1.f : 0.f);
    YH_TR(56);
    if (WRITE_DY) {
        float* ocl = ocell + coff;
#pragma unroll
        for (int k = 0; k < kClsRegs; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float pc = lg[k] * inv;
                ocl[c] += p.ccls * pc * (pc - (c == hd.w ? 1.f : 0.f) - dot);
            }
        }
        for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
            const float pc = expf(cl[c] - mx) * inv;
            ocl[c] += p.ccls * pc * (pc - (c == hd.w ? 1.f : 0.f) - dot);
        }
    }
    __syncwarp();
    YH_TR(57);
}

// TV/TA/TC != 0 fix version / boxes per cell / classes at compile time (index arithmetic folds,
// divisions become shifts and multiplies); 0 keeps them as run-time values from the geometry.
template <bool WRITE_DY, int TV, int TA, int TC>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) yh_train_kernel(const TrainParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_off[kOffCap];
    __shared__ __align__(16) YhGt s_gt[kGtCap];
    __shared__ unsigned int s_chunkmask[kMaskWords];  // mini-chunks of this CTA that contain records
    __shared__ int s_cell[kGtCap];  // flat cell index (image * cells + cy * s_w + cx) of each cached record
    __shared__ float red[kMaxWarps * 6];
    __shared__ bool is_last;

    const YhGeom& g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    YH_TR(0);
    constexpr bool kFixed = TV != 0 && TA != 0 && TC != 0;
    const int version = TV ? TV : g.version;
    const int A = TA ? TA : g.a, C = TC ? TC : g.c;
    const int bs = version == 2 ? 5 + C : 5;
    const int cf = version == 2 ? A * (5 + C) : 5 * A + C;
    const int cells = g.cells;
    const int W = kFixed ? kMaxWarps : 1 << p.warps_log2;
    const int wlog = kFixed ? 3 : p.warps_log2;
    static_assert(kMaxWarps == 8, "wlog above assumes 8 warps");
    const int mc = kFixed ? mini_cells(cf) : p.mc;
    const int S = mc * cf;  // floats per stage (multiple of 4)

    // this CTA's contiguous cell range (in quads of cells so chunk starts stay 16-B aligned);
    // everything below is 32-bit: train_impl checks that the tensor has < 2^31 floats
    const int bid = blockIdx.x;
    const int q0 = bid * p.q_base + min(bid, p.q_extra);
    const int q1 = q0 + p.q_base + (bid < p.q_extra ? 1 : 0);
    const int cta_cell0 = q0 * 4;
    const int cta_cell1 = min(q1 * 4, (int)p.total_cells);
    const int cta_cells = cta_cell1 - cta_cell0;

    // ---- metadata first (before the bulk loads flood the memory system): the CTA's slice of the
    //      CSR offsets, and -- in the same round trip -- a speculative window of records placed
    //      where the CTA's records sit if boxes are spread evenly over the images
    const int n_first = cta_cells > 0 ? cta_cell0 / cells : 0;
    const int n_last = cta_cells > 0 ? (cta_cell1 - 1) / cells : -1;
    const int n_imgs = n_last - n_first + 1;
    const bool off_cached = n_imgs + 1 <= min(kOffCap, (int)blockDim.x);
    int w0 = 0, wn = 0;
    if (cta_cells > 0 && p.m_local > 0) {
        const int cap = min(kGtCap, (int)blockDim.x * kWinRegs / 3);
        const int mid = (int)(p.rec_per_cell * (float)(cta_cell0 + (cta_cells >> 1)));
        wn = min(cap, p.m_local);
        w0 = max(0, min(mid - (cap >> 1), p.m_local - wn));
    }
    int4 wv[kWinRegs];
    {
        const int4* src = reinterpret_cast<const int4*>(p.gt + w0);
#pragma unroll
        for (int q = 0; q < kWinRegs; ++q) {
            const int i = tid + q * blockDim.x;
            wv[q] = i < wn * 3 ? __ldg(src + i) : make_int4(0, 0, 0, 0);
        }
    }
    const int offv = (off_cached && tid <= n_imgs) ? __ldg(p.gt_off + n_first + tid) : 0;
    YH_TR(58);

    // shared-memory carve-up: per warp kInStages input stages (+ kOutStages output stages)
    float* in_base = reinterpret_cast<float*>(smem_raw) + warp * kInStages * S;
    float* out_base = reinterpret_cast<float*>(smem_raw) + W * kInStages * S + warp * kOutStages * S;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) +
                                                 W * (kInStages + (WRITE_DY ? kOutStages : 0)) * S) +
                     warp * kInStages;
    const int nmini = cta_cells > 0 ? (cta_cells + mc - 1) / mc : 0;
    const int my_n = nmini > warp ? (nmini - warp + W - 1) >> wlog : 0;  // mini-chunks of this warp
    const int step = W * mc;  // cells between consecutive mini-chunks of one warp

    auto issue_load = [&](int k, int st) {  // lane 0 only: mini-chunk k of this warp into stage st
        const int c0 = cta_cell0 + warp * mc + k * step;
        const int nc = min(mc, cta_cell1 - c0);
        const uint32_t bytes = ((uint32_t)nc * cf * 4u) & ~15u;
        uint64_t* bar = &bars[st];
        if (bytes) {
            yh_mbar_expect_tx(bar, bytes);
            yh_bulk_load(in_base + st * S, p.y + (size_t)c0 * cf, bytes, bar);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(bar)) : "memory");
        }
    };
    // every warp's FIRST chunk goes out before anybody's second one: work can start sooner
    if (p.tma_in && lane == 0) {
#pragma unroll
        for (int s = 0; s < kInStages; ++s) yh_mbar_init(&bars[s], 1);
        yh_mbar_fence_init();
        if (my_n > 0) issue_load(0, 0);
    }
    YH_TR(49);
    if (WRITE_DY) {  // output stages start out all-zero and are kept that way between chunks
        float4* o4 = reinterpret_cast<float4*>(out_base);
        const int n4 = (kOutStages * S) >> 2;
        for (int i = lane; i < n4; i += 32) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.tma_in && lane == 0) {
#pragma unroll
        for (int k = 1; k < kInStages; ++k)
            if (k < my_n) issue_load(k, k);
    }
    if (tid < kMaskWords) s_chunkmask[tid] = 0u;
    YH_TR(59);
    __syncthreads();
    YH_TR(60);

    const bool use_mask = nmini <= 32 * kMaskWords;
    // flat cell index of a record header (or -1), and "this chunk has records" bookkeeping
    auto note_record = [&](const int4& h, int slot) {
        const bool ok = h.y >= 0 && h.y < g.s_h && h.z >= 0 && h.z < g.s_w && h.x >= 0 && h.x < g.n;
        const int gc = ok ? h.x * cells + h.y * g.s_w + h.z : -1;
        s_cell[slot] = gc;
        if (use_mask && gc >= cta_cell0 && gc < cta_cell1) {
            const int ch = (gc - cta_cell0) / mc;
            atomicOr(&s_chunkmask[ch >> 5], 1u << (ch & 31));
        }
    };
    if (off_cached && tid <= n_imgs) s_off[tid] = offv;
#pragma unroll
    for (int q = 0; q < kWinRegs; ++q) {
        const int i = tid + q * blockDim.x;
        if (i < wn * 3) {
            reinterpret_cast<int4*>(s_gt)[i] = wv[q];
            if (i % 3 == 0) note_record(wv[q], i / 3);
        }
    }
    YH_TR(61);
    __syncthreads();
    YH_TR(62);
    const int rec0 = n_imgs > 0 ? (off_cached ? s_off[0] : __ldg(p.gt_off + n_first)) : 0;
    const int rec1 = n_imgs > 0 ? (off_cached ? s_off[n_imgs] : __ldg(p.gt_off + n_last + 1)) : 0;
    int gt_base = w0;
    bool gt_cached = rec0 >= w0 && rec1 <= w0 + wn;
    if (!gt_cached && rec1 - rec0 <= kGtCap) {  // the guess missed: fetch exactly the CTA's records
        __syncthreads();
        if (tid < kMaskWords) s_chunkmask[tid] = 0u;
        __syncthreads();
        const int4* src = reinterpret_cast<const int4*>(p.gt + rec0);
        for (int i = tid; i < (rec1 - rec0) * 3; i += blockDim.x) {
            const int4 v = __ldg(src + i);
            reinterpret_cast<int4*>(s_gt)[i] = v;
            if (i % 3 == 0) note_record(v, i / 3);
        }
        gt_base = rec0;
        gt_cached = true;
        __syncthreads();
    }
    const bool have_mask = use_mask && gt_cached;
    auto off_at = [&](int n) -> int { return off_cached ? s_off[n - n_first] : __ldg(p.gt_off + n); };
    YH_TR(1);

    WarpSums sums = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float my_pw = lane < 5 * A ? g.pw[lane / 5] : 0.f;  // anchor multipliers of lane / 5
    const float my_ph = lane < 5 * A ? g.ph[lane / 5] : 0.f;
    unsigned touched0 = 0u, touched1 = 0u;  // cells of output stage 0 / 1 that hold sparse gradient rows
    static_assert(kOutStages == 2, "touched0/touched1 track exactly two output stages");

    // running state of this warp's current chunk: image n0, first cell rem0 inside it, stages
    int cell0 = cta_cell0 + warp * mc;
    int n0 = my_n > 0 ? cell0 / cells : 0;
    int rem0 = cell0 - n0 * cells;
    int ist = 0, ost = 0;
    uint32_t in_phase = 0;
    const int rows_per_cell = version == 2 ? A : 1;  // dense-pass rows: predictors (v2) / cells (v1)
    const bool two_img = mc <= cells;                   // a chunk then touches at most two images

    for (int k = 0; k < my_n; ++k) {
        const int ncell = min(mc, cta_cell1 - cell0);
        const int nfl = ncell * cf;
        const int ntail = nfl & 3;  // floats past the last 16-byte boundary (last chunk of the tensor only)
        const int nfl16 = nfl - ntail;
        float* in = in_base + ist * S;
        float* out = out_base + ost * S;
        const float* ysrc = p.y + (size_t)cell0 * cf;

        if (p.tma_in) {
            if (lane < ntail) in[nfl16 + lane] = __ldg(ysrc + nfl16 + lane);
            yh_mbar_wait(&bars[ist], in_phase);
        } else {
            for (int i = lane; i < nfl; i += 32) in[i] = __ldg(ysrc + i);
        }
        YH_TR(4 + 5 * k);
        if (WRITE_DY) {
            // the out stage is reused every kOutStages chunks: its bulk store must have drained,
            // then the sparse rows it carried are cleared again
            if (p.tma_out && lane == 0) yh_bulk_wait_read<kOutStages - 1>();
            __syncwarp();
            unsigned m = ost ? touched1 : touched0;
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                for (int q = lane; q < cf; q += 32) out[c * cf + q] = 0.f;
            }
            if (ost) touched1 = 0u; else touched0 = 0u;
        }
        __syncwarp();
        YH_TR(5 + 5 * k);

        // ---------------- dense pass: no-object term, one row owner per lane ----------------
        {
            const int o0 = off_at(n0), o1 = off_at(n0 + 1);
            const float kn0 = (float)(o1 - o0);
            const float kn1 = n0 < n_last ? (float)(off_at(n0 + 2) - o1) : kn0;
            const int rows_in_n0 = (cells - rem0) * rows_per_cell;  // rows before the next image starts
            const int nrows = ncell * rows_per_cell;
            auto kn_of = [&](int u) -> float {
                if (two_img) return u < rows_in_n0 ? kn0 : kn1;
                const int n = n0 + (rem0 + u / rows_per_cell) / cells;
                return (float)(off_at(n + 1) - off_at(n));
            };
            if (version == 2) {
                // two rows per lane and step: two independent exp/div chains in flight
                float acc1 = 0.f;
                for (int u = lane; u < nrows; u += 64) {
                    const int u2 = u + 32;
                    const bool has2 = u2 < nrows;
                    const float ta = in[u * bs + 4];
                    const float tb = has2 ? in[u2 * bs + 4] : 0.f;
                    const float ca_ = yh_sigmoid(ta), cb_ = yh_sigmoid(tb);
                    const float ka = kn_of(u), kb = has2 ? kn_of(u2) : 0.f;
                    const float a2 = ca_ * ca_, b2 = cb_ * cb_;
                    sums.no += ka * a2;
                    acc1 += kb * b2;
                    if (WRITE_DY) {
                        out[u * bs + 4] = p.cno * ka * a2 * (1.f - ca_);
                        if (has2) out[u2 * bs + 4] = p.cno * kb * b2 * (1.f - cb_);
                    }
                }
                sums.no += acc1;
            } else {
                for (int u = lane; u < nrows; u += 32) {
                    const float kn = kn_of(u);
                    const float* irow = in + u * cf;
                    float* orow = out + u * cf;
                    for (int b = 0; b < A; ++b) {
                        const float conf = yh_sigmoid(irow[b * 5 + 4]);
                        const float c2 = conf * conf;
                        sums.no += kn * c2;
                        if (WRITE_DY) orow[b * 5 + 4] = p.cno * kn * c2 * (1.f - conf);
                    }
                }
            }
        }
        __syncwarp();  // dense rows written before the sparse read-modify-writes
        YH_TR(6 + 5 * k);

        // ---------------- sparse pass: ground-truth records of this chunk ----------------
        const int ci = warp + (k << wlog);  // index of this mini-chunk inside the CTA
        if (!have_mask || ((s_chunkmask[ci >> 5] >> (ci & 31)) & 1u)) {
            const int n_hi = two_img ? (rem0 + ncell > cells ? n0 + 1 : n0) : n0 + (rem0 + ncell - 1) / cells;
            const int r0 = off_at(n0), r1 = off_at(n_hi + 1);
            for (int base = r0; base < r1; base += 32) {
                const int j = base + lane;
                int lc = -1;
                if (j < r1) {
                    int gc;
                    if (gt_cached) {
                        gc = s_cell[j - gt_base];
                    } else {
                        const int4 h = __ldg(reinterpret_cast<const int4*>(p.gt + j));
                        const bool ok = h.y >= 0 && h.y < g.s_h && h.z >= 0 && h.z < g.s_w && h.x >= 0 && h.x < g.n;
                        gc = ok ? h.x * cells + h.y * g.s_w + h.z : -1;
                    }
                    if (gc >= cell0 && gc < cell0 + ncell) lc = gc - cell0;
                }
                unsigned bal = __ballot_sync(0xffffffffu, lc >= 0);
                while (bal) {
                    const int b = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int jj = base + b;
                    const int lcell = __shfl_sync(0xffffffffu, lc, b);
                    RecordRegs rr;
                    if (gt_cached) {
                        const int4* rp = reinterpret_cast<const int4*>(s_gt + (jj - gt_base));
                        rr.hd = rp[0];
                        rr.tt = *reinterpret_cast<const float4*>(rp + 1);
                        rr.bb = *reinterpret_cast<const float4*>(rp + 2);
                    } else {
                        const int4* rp = reinterpret_cast<const int4*>(p.gt + jj);
                        rr.hd = __ldg(rp);
                        rr.tt = __ldg(reinterpret_cast<const float4*>(rp + 1));
                        rr.bb = __ldg(reinterpret_cast<const float4*>(rp + 2));
                    }
                    process_record<WRITE_DY>(p, version, A, C, rr, jj, in + lcell * cf, out + lcell * cf, lane, my_pw, my_ph, sums);
                    if (ost) touched1 |= 1u << lcell; else touched0 |= 1u << lcell;
                }
            }
        }

        YH_TR(7 + 5 * k);
        // ---------------- push the chunk's dL/dy, refill the input stage ----------------
        if (WRITE_DY) {
            float* ydst = p.dy + (size_t)cell0 * cf;
            if (p.tma_out) {
                yh_fence_proxy_async();  // every lane: its generic-proxy writes -> async proxy
                __syncwarp();
                if (lane == 0) {
                    if (nfl16) yh_bulk_store(ydst, out, (uint32_t)nfl16 * 4u);
                    yh_bulk_commit();
                }
                if (lane < ntail) ydst[nfl16 + lane] = out[nfl16 + lane];
            } else {
                __syncwarp();
                for (int i = lane; i < nfl; i += 32) ydst[i] = out[i];
            }
        } else {
            __syncwarp();
        }
        if (p.tma_in && lane == 0 && k + kInStages < my_n) issue_load(k + kInStages, ist);

        YH_TR(8 + 5 * k);
        cell0 += step;
        rem0 += step;
        while (rem0 >= cells) { rem0 -= cells; ++n0; }
        ost ^= 1;
        if (++ist == kInStages) { ist = 0; in_phase ^= 1u; }
    }
    YH_TR(2);
    if (WRITE_DY && p.tma_out && lane == 0) yh_bulk_wait_all<0>();
    YH_TR(3);

    // ---------------- block reduction of the six partial sums ----------------
    // every lane carries partial sums (dense rows; record channels by lane role): fold the warp
    const float s_no = yh_warp_sum(sums.no);
    const float s_xy = yh_warp_sum(sums.xy), s_wh = yh_warp_sum(sums.wh), s_conf = yh_warp_sum(sums.conf);
    const float s_nr = yh_warp_sum(sums.nr), s_cls = yh_warp_sum(sums.cls);
    if (lane == 0) {
        float* r = red + warp * 6;
        r[0] = s_xy; r[1] = s_wh; r[2] = s_conf; r[3] = s_no; r[4] = s_nr; r[5] = s_cls;
    }
    __syncthreads();
    if (tid == 0) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < W; ++w)
            for (int q = 0; q < 6; ++q) acc[q] += red[w * 6 + q];
        float* dst = p.partials + (size_t)blockIdx.x * kPartials;
        for (int q = 0; q < 6; ++q) dst[q] = acc[q];
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (unsigned b = lane; b < gridDim.x; b += 32) {
            const float* src = p.partials + (size_t)b * kPartials;
            for (int q = 0; q < 6; ++q) acc[q] += (double)__ldcg(src + q);
        }
        for (int q = 0; q < 6; ++q)
            for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) {
            const double t0 = acc[0] * p.inv_den[0];
            const double t1 = acc[1] * p.inv_den[1];
            const double t2 = acc[2] * p.inv_den[2];
            const double t3 = (acc[3] - acc[4]) * p.inv_den[3];
            const double t4 = acc[5] * p.inv_den[4];
            p.terms[0] = (float)t0; p.terms[1] = (float)t1; p.terms[2] = (float)t2;
            p.terms[3] = (float)t3; p.terms[4] = (float)t4;
            p.loss[0] = (float)(p.lam[0] * t0 + p.lam[1] * t1 + p.lam[2] * t2 + p.lam[3] * t3 + p.lam[4] * t4);
            *p.ticket = 0u;  // ready for the next launch
        }
    }
    YH_TR(63);
}

template <bool WDY, int TV, int TA, int TC>
int launch_variant(const TrainParams& p, int grid, size_t smem, int warps, cudaStream_t stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_train_kernel<WDY, TV, TA, TC>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(train)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    yh_train_kernel<WDY, TV, TA, TC><<<grid, warps * 32, smem, stream>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_train launch");
}

template <bool WDY>
int launch_geometry(const TrainParams& p, int grid, size_t smem, int warps, cudaStream_t stream) {
    const YhGeom& g = p.g;
    // compile-time geometries for the shapes the reference trains: YOLOv2 5 anchors x 20 classes
    // (VOC, models/yolov2.py:49-70) and YOLOv1 B=2, C=20 (config.py:7-11); anything else runs the
    // same kernel with run-time geometry
    if (warps == kMaxWarps && p.mc == mini_cells(g.cell_floats)) {
        if (g.version == 2 && g.a == 5 && g.c == 20) return launch_variant<WDY, 2, 5, 20>(p, grid, smem, warps, stream);
        if (g.version == 1 && g.a == 2 && g.c == 20) return launch_variant<WDY, 1, 2, 20>(p, grid, smem, warps, stream);
    }
    return launch_variant<WDY, 0, 0, 0>(p, grid, smem, warps, stream);
}

int train_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
               const float* anchors_wh_host, float img_h, float img_w, const YhGt* gt,
               const int32_t* gt_off, int m_local, int m_global, const float* lambdas_host,
               float* dy, float* terms, float* loss, int32_t* resp, float* iou_resp, void* ws,
               size_t ws_bytes, void* stream) {
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
    p.partials = reinterpret_cast<float*>(ws);
    p.ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + (size_t)kMaxGrid * kPartials * 4);
    p.total_cells = (long long)n * p.g.cells;
    p.m_local = m_local;
    p.quads_total = (p.total_cells + 3) / 4;
    const int cf = p.g.cell_floats;
    p.mc = mini_cells(cf);
    const int qpc = p.mc / 4;  // quads of cells per mini-chunk
    YH_REQUIRE(p.total_cells * cf < (1ll << 31), YH_ERR_UNSUPPORTED, "head tensor has 2^31 or more floats");
    p.tma_in = ((uintptr_t)y & 15) == 0;
    p.tma_out = dy && ((uintptr_t)dy & 15) == 0;

    const double M = (double)m_global;
    const double P1 = (double)p.g.preds - 1.0;
    for (int i = 0; i < 5; ++i) p.lam[i] = lambdas_host[i];
    p.inv_den[0] = 1.0 / (2.0 * M);
    p.inv_den[1] = 1.0 / (2.0 * M);
    p.inv_den[2] = 1.0 / M;
    p.inv_den[3] = 1.0 / (M * P1);
    p.inv_den[4] = 1.0 / M;
    // d(mean)/d(activation):  xy: 2(s-t)/(2M);  wh: 2(q-T)/(2M) * dq/dt with dq/dt = q/2;
    // conf: 2(conf-iou)/M;  noobj: 2 conf /(M(P-1));  cls: 2/M
    p.cxy = (float)(lambdas_host[0] / M);
    p.cwh = (float)(lambdas_host[1] / (2.0 * M));
    p.cconf = (float)(lambdas_host[2] * 2.0 / M);
    p.cno = (float)(lambdas_host[3] * 2.0 / (M * P1));
    p.ccls = (float)(lambdas_host[4] * 2.0 / M);

    // warps per CTA: as many private pipelines as fit in shared memory
    const size_t stage_bytes = (size_t)p.mc * cf * 4;
    const int stages = kInStages + (dy ? kOutStages : 0);
    const size_t budget = 200 * 1024;
    int warps = kMaxWarps;
    while (warps > 1 && (size_t)warps * stages * stage_bytes + 8 * kInStages * warps + 16 > budget) warps >>= 1;
    const size_t smem = (size_t)warps * stages * stage_bytes + 8 * kInStages * warps + 16;
    YH_REQUIRE(smem <= budget, YH_ERR_UNSUPPORTED, "cell too wide for shared memory (%d floats per cell)", cf);
    p.warps_log2 = 0;
    while ((1 << p.warps_log2) < warps) ++p.warps_log2;

    const long long nmini_total = (p.quads_total + qpc - 1) / qpc;
    long long grid = (nmini_total + warps - 1) / warps;  // at least one mini-chunk per warp
    const int sms = yh_sm_count();
    if (grid > sms) grid = sms;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (grid < 1) grid = 1;
    p.q_base = (int)(p.quads_total / grid);
    p.q_extra = (int)(p.quads_total % grid);
    p.rec_per_cell = (float)((double)m_local / (double)p.total_cells);
    cudaStream_t st = (cudaStream_t)stream;
    return dy ? launch_geometry<true>(p, (int)grid, smem, warps, st) : launch_geometry<false>(p, (int)grid, smem, warps, st);
}

}  // namespace

extern "C" {

size_t yh_train_workspace_bytes(void) { return (size_t)kMaxGrid * kPartials * 4 + 128; }

int yh_v2_train(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                float img_h, float img_w, const YhGt* gt, const int32_t* gt_off, int m_local,
                int m_global, const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return train_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, gt, gt_off, m_local,
                      m_global, lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream);
}

int yh_v1_train(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss, int32_t* resp,
                float* iou_resp, void* ws, size_t ws_bytes, void* stream) {
    return train_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, gt, gt_off, m_local, m_global,
                      lambdas_host, dy, terms, loss, resp, iou_resp, ws, ws_bytes, stream);
}

}  // extern "C"
