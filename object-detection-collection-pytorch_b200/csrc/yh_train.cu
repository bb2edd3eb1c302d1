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
constexpr int kClsRegs = 4;      // class logits per lane kept in registers (C <= 128)

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
    int warps;              // warps per CTA
    int tma_in, tma_out;    // base pointers 16-byte aligned
    float lam[5];
    double inv_den[5];      // 1/(2M), 1/(2M), 1/M, 1/(M(P-1)), 1/M
    float cxy, cwh, cconf, cno, ccls;  // gradient coefficients (see train_impl)
};

// "a beats b" for torch.max semantics: larger wins, NaN beats everything, first index on ties
__device__ __forceinline__ bool yh_better(float va, int ia, float vb, int ib) {
    const bool na = va != va, nb = vb != vb;
    if (na || nb) return na && (!nb || ia < ib);
    return va > vb || (va == vb && ia < ib);
}

struct WarpSums {
    float no, xy, wh, conf, nr, cls;
};

// One ground-truth record against its cell (all 32 lanes cooperate).  `cellp` / `ocell` point at
// the cell's floats in the input / output stage.
template <bool WRITE_DY>
__device__ __forceinline__ void process_record(const TrainParams& p, const YhGt* rec, int jj,
                                               const float* cellp, float* ocell, int lane, WarpSums& s) {
    const YhGeom& g = p.g;
    const int A = g.a, C = g.c, bs = g.box_stride;
    const int4 hd = *reinterpret_cast<const int4*>(rec);             // img, cy, cx, cls
    const float4 tt = *(reinterpret_cast<const float4*>(rec) + 1);   // stx, sty, tw, th
    const float4 bb = *(reinterpret_cast<const float4*>(rec) + 2);   // x1, y1, x2, y2

    // lanes < A: decode the box of anchor `lane` and its IoU with the record
    float sx = 0.f, sy = 0.f, wa = 0.f, ha = 0.f, conf = 0.f;
    float iou = -INFINITY;
    int best = 1 << 20;
    if (lane < A) {
        const float* bp = cellp + lane * bs;
        sx = yh_sigmoid(bp[0]);
        sy = yh_sigmoid(bp[1]);
        if (g.version == 2) {
            wa = expf(bp[2]);
            ha = expf(bp[3]);
        } else {
            wa = yh_sigmoid(bp[2]);
            ha = yh_sigmoid(bp[3]);
        }
        conf = yh_sigmoid(bp[4]);
        const YhBox pb = yh_decode_box(sx, sy, wa, ha, g.pw[lane], g.ph[lane], hd.z, hd.y, g.gw, g.gh);
        const YhBox gb{bb.x, bb.y, bb.z, bb.w};
        iou = yh_iou_xyxy(pb, gb);
        best = lane;
    }
    float bv = iou;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best, o);
        if (yh_better(ov, oi, bv, best)) { bv = ov; best = oi; }
    }
    const int r = best;  // responsible predictor (uniform across the warp)
    sx = __shfl_sync(0xffffffffu, sx, r);
    sy = __shfl_sync(0xffffffffu, sy, r);
    wa = __shfl_sync(0xffffffffu, wa, r);
    ha = __shfl_sync(0xffffffffu, ha, r);
    conf = __shfl_sync(0xffffffffu, conf, r);
    const float iou_r = bv;

    // class term: softmax over C logits, lanes stride the classes (logits cached in registers)
    const int coff = g.version == 2 ? r * bs + 5 : 5 * A;
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
    mx = yh_warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        lg[k] = lane + 32 * k < C ? expf(lg[k] - mx) : 0.f;
        se += lg[k];
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) se += expf(cl[c] - mx);
    se = yh_warp_sum(se);
    float sq = 0.f, dot = 0.f;
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
            lg[k] = __fdiv_rn(lg[k], se);  // p_c
            const float gg = lg[k] - (c == hd.w ? 1.f : 0.f);
            sq += gg * gg;
            dot += gg * lg[k];
        }
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
        const float pc = __fdiv_rn(expf(cl[c] - mx), se);
        const float gg = pc - (c == hd.w ? 1.f : 0.f);
        sq += gg * gg;
        dot += gg * pc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }

    if (lane == 0) {
        float tw_t, th_t;
        if (g.version == 2) {  // sqrt(bwbh / pwph), models/yolov2.py:946-947
            tw_t = __fsqrt_rn(__fdiv_rn(tt.z, g.pw[r]));
            th_t = __fsqrt_rn(__fdiv_rn(tt.w, g.ph[r]));
        } else {               // sqrt(sig_twth), models/yolov1.py:760-761
            tw_t = __fsqrt_rn(tt.z);
            th_t = __fsqrt_rn(tt.w);
        }
        const float sqw = __fsqrt_rn(wa), sqh = __fsqrt_rn(ha);
        const float dx = sx - tt.x, dyv = sy - tt.y;
        const float dw = sqw - tw_t, dh = sqh - th_t;
        const float dc = conf - iou_r;
        s.xy += dx * dx + dyv * dyv;
        s.wh += dw * dw + dh * dh;
        s.conf += dc * dc;
        s.nr += conf * conf;
        s.cls += sq;
        if (p.resp) p.resp[jj] = r;
        if (p.iou_resp) p.iou_resp[jj] = iou_r;
        if (WRITE_DY) {
            float* row = ocell + r * bs;
            row[0] += p.cxy * dx * sx * (1.f - sx);
            row[1] += p.cxy * dyv * sy * (1.f - sy);
            if (g.version == 2) {
                row[2] += p.cwh * dw * sqw;
                row[3] += p.cwh * dh * sqh;
            } else {
                row[2] += p.cwh * dw * sqw * (1.f - wa);
                row[3] += p.cwh * dh * sqh * (1.f - ha);
            }
            row[4] += (p.cconf * dc - p.cno * conf) * conf * (1.f - conf);
        }
    }
    if (WRITE_DY) {
        float* ocl = ocell + coff;
#pragma unroll
        for (int k = 0; k < kClsRegs; ++k) {
            const int c = lane + 32 * k;
            if (c < C) ocl[c] += p.ccls * lg[k] * (lg[k] - (c == hd.w ? 1.f : 0.f) - dot);
        }
        for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
            const float pc = __fdiv_rn(expf(cl[c] - mx), se);
            ocl[c] += p.ccls * pc * (pc - (c == hd.w ? 1.f : 0.f) - dot);
        }
    }
    __syncwarp();
}

template <bool WRITE_DY>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) yh_train_kernel(const TrainParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_off[kOffCap];
    __shared__ __align__(16) YhGt s_gt[kGtCap];
    __shared__ int s_cell[kGtCap];  // flat cell index (image * cells + cy * s_w + cx) of each cached record
    __shared__ float red[kMaxWarps * 6];
    __shared__ bool is_last;

    const YhGeom& g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = p.warps;
    const int cf = g.cell_floats, bs = g.box_stride, A = g.a, cells = g.cells;
    const int mc = p.mc;
    const int S = mc * cf;  // floats per stage (multiple of 4)

    // shared-memory carve-up: per warp kInStages input stages (+ kOutStages output stages)
    float* in_base = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kInStages * S;
    float* out_base = reinterpret_cast<float*>(smem_raw) + (size_t)W * kInStages * S + (size_t)warp * kOutStages * S;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) +
                                                 (size_t)W * (kInStages + (WRITE_DY ? kOutStages : 0)) * S) +
                     warp * kInStages;

    // this CTA's contiguous cell range (in quads of cells so chunk starts stay 16-B aligned);
    // everything below is 32-bit: train_impl checks that the tensor has < 2^31 floats
    const int q0 = (int)((long long)p.quads_total * blockIdx.x / gridDim.x);
    const int q1 = (int)((long long)p.quads_total * (blockIdx.x + 1) / gridDim.x);
    const int cta_cell0 = q0 * 4;
    const int cta_cell1 = min(q1 * 4, (int)p.total_cells);
    const int cta_cells = cta_cell1 - cta_cell0;
    const int nmini = cta_cells > 0 ? (cta_cells + mc - 1) / mc : 0;
    const int my_n = nmini > warp ? (nmini - warp + W - 1) / W : 0;  // mini-chunks of this warp

    auto issue_load = [&](int k) {  // lane 0 only: mini-chunk k of this warp into stage k % kInStages
        const int c0 = cta_cell0 + (warp + k * W) * mc;
        const int nc = min(mc, cta_cell1 - c0);
        const uint32_t bytes = ((uint32_t)nc * cf * 4u) & ~15u;
        uint64_t* bar = &bars[k % kInStages];
        if (bytes) {
            yh_mbar_expect_tx(bar, bytes);
            yh_bulk_load(in_base + (k % kInStages) * S, p.y + (size_t)c0 * cf, bytes, bar);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(bar)) : "memory");
        }
    };
    if (p.tma_in && lane == 0) {
        for (int s = 0; s < kInStages; ++s) yh_mbar_init(&bars[s], 1);
        yh_mbar_fence_init();
        const int pre = my_n < kInStages ? my_n : kInStages;
        for (int k = 0; k < pre; ++k) issue_load(k);
    }
    if (WRITE_DY) {  // output stages start out all-zero and are kept that way between chunks
        float4* o4 = reinterpret_cast<float4*>(out_base);
        for (int i = lane; i < kOutStages * S / 4; i += 32) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // ---- stage the CTA's CSR offsets and ground-truth records in shared memory ----
    const int n_first = cta_cells > 0 ? cta_cell0 / cells : 0;
    const int n_last = cta_cells > 0 ? (cta_cell1 - 1) / cells : -1;
    const int n_imgs = n_last - n_first + 1;
    const bool off_cached = n_imgs + 1 <= kOffCap;
    if (off_cached)
        for (int i = tid; i <= n_imgs; i += blockDim.x) s_off[i] = __ldg(p.gt_off + n_first + i);
    __syncthreads();
    const int rec0 = n_imgs > 0 ? (off_cached ? s_off[0] : __ldg(p.gt_off + n_first)) : 0;
    const int rec1 = n_imgs > 0 ? (off_cached ? s_off[n_imgs] : __ldg(p.gt_off + n_last + 1)) : 0;
    const bool gt_cached = rec1 - rec0 <= kGtCap;
    if (gt_cached) {
        const int4* src = reinterpret_cast<const int4*>(p.gt + rec0);
        int4* dst = reinterpret_cast<int4*>(s_gt);
        for (int i = tid; i < (rec1 - rec0) * 3; i += blockDim.x) {
            const int4 v = __ldg(src + i);
            dst[i] = v;
            if (i % 3 == 0) {  // header: img, cy, cx, cls
                const bool ok = v.y >= 0 && v.y < g.s_h && v.z >= 0 && v.z < g.s_w && v.x >= 0 && v.x < g.n;
                s_cell[i / 3] = ok ? v.x * cells + v.y * g.s_w + v.z : -1;
            }
        }
    }
    __syncthreads();
    auto off_at = [&](int n) -> int { return off_cached ? s_off[n - n_first] : __ldg(p.gt_off + n); };

    WarpSums sums = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned touched[kOutStages];  // cells of each output stage that hold sparse gradient rows
#pragma unroll
    for (int s = 0; s < kOutStages; ++s) touched[s] = 0u;

    // running position of this warp's current chunk: image n0, first cell rem0 inside it
    int cell0 = cta_cell0 + warp * mc;
    int n0 = my_n > 0 ? cell0 / cells : 0;
    int rem0 = cell0 - n0 * cells;
    const int step = W * mc;
    const int rows_per_cell = g.version == 2 ? A : 1;  // dense-pass rows: predictors (v2) / cells (v1)
    const int row_floats = g.version == 2 ? bs : cf;

    for (int k = 0; k < my_n; ++k) {
        const int ncell = min(mc, cta_cell1 - cell0);
        const int nfl = ncell * cf;
        const int nfl16 = nfl & ~3;  // floats covered by the bulk copies
        const int ist = k % kInStages, ost = k % kOutStages;
        float* in = in_base + ist * S;
        float* out = out_base + ost * S;
        const float* ysrc = p.y + (size_t)cell0 * cf;

        if (p.tma_in) {
            if (lane < nfl - nfl16) in[nfl16 + lane] = __ldg(ysrc + nfl16 + lane);  // < 16-B tail
            yh_mbar_wait(&bars[ist], (uint32_t)((k / kInStages) & 1));
        } else {
            for (int i = lane; i < nfl; i += 32) in[i] = __ldg(ysrc + i);
        }
        if (WRITE_DY) {
            // the out stage is reused every kOutStages chunks: its bulk store must have drained,
            // then the sparse rows it carried are cleared again
            if (p.tma_out && lane == 0) yh_bulk_wait_read<kOutStages - 1>();
            __syncwarp();
            unsigned m = touched[ost];
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                for (int q = lane; q < cf; q += 32) out[c * cf + q] = 0.f;
            }
            touched[ost] = 0u;
        }
        __syncwarp();

        // ---------------- dense pass: no-object term, one row owner per lane ----------------
        {
            const int n1 = n0 < n_last ? n0 + 1 : n0;
            const float kn0 = (float)(off_at(n0 + 1) - off_at(n0));
            const float kn1 = n1 != n0 ? (float)(off_at(n1 + 1) - off_at(n1)) : kn0;
            const int rows_in_n0 = (cells - rem0) * rows_per_cell;  // rows before the next image starts
            const bool two_img = mc <= cells;                        // a chunk then touches <= 2 images
            const int nrows = ncell * rows_per_cell;
            for (int u = lane; u < nrows; u += 32) {
                float kn;
                if (two_img) kn = u < rows_in_n0 ? kn0 : kn1;
                else { const int n = n0 + (rem0 + u / rows_per_cell) / cells; kn = (float)(off_at(n + 1) - off_at(n)); }
                const float* irow = in + u * row_floats;
                float* orow = out + u * row_floats;
                if (g.version == 2) {
                    const float conf = yh_sigmoid(irow[4]);
                    const float c2 = conf * conf;
                    sums.no += kn * c2;
                    if (WRITE_DY) orow[4] = p.cno * kn * c2 * (1.f - conf);
                } else {
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

        // ---------------- sparse pass: ground-truth records of this chunk ----------------
        {
            const int n_hi = n0 + (rem0 + ncell - 1) / cells;
            const int r0 = off_at(n0), r1 = off_at(n_hi + 1);
            for (int base = r0; base < r1; base += 32) {
                const int j = base + lane;
                int lc = -1;
                if (j < r1) {
                    int gc;
                    if (gt_cached) {
                        gc = s_cell[j - rec0];
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
                    const YhGt* rec = gt_cached ? s_gt + (jj - rec0) : p.gt + jj;
                    process_record<WRITE_DY>(p, rec, jj, in + lcell * cf, out + lcell * cf, lane, sums);
                    touched[ost] |= 1u << lcell;
                }
            }
        }

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
                if (lane < nfl - nfl16) ydst[nfl16 + lane] = out[nfl16 + lane];
            } else {
                __syncwarp();
                for (int i = lane; i < nfl; i += 32) ydst[i] = out[i];
            }
        } else {
            __syncwarp();
        }
        if (p.tma_in && lane == 0 && k + kInStages < my_n) issue_load(k + kInStages);

        cell0 += step;
        rem0 += step;
        while (rem0 >= cells) { rem0 -= cells; ++n0; }
    }
    if (WRITE_DY && p.tma_out && lane == 0) yh_bulk_wait_all<0>();

    // ---------------- block reduction of the six partial sums ----------------
    const float s_no = yh_warp_sum(sums.no);
    if (lane == 0) {
        float* r = red + warp * 6;
        r[0] = sums.xy; r[1] = sums.wh; r[2] = sums.conf; r[3] = s_no; r[4] = sums.nr; r[5] = sums.cls;
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
}

template <bool W>
int launch_variant(const TrainParams& p, int grid, size_t smem, cudaStream_t stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_train_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(train)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    yh_train_kernel<W><<<grid, p.warps * 32, smem, stream>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_train launch");
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
    YH_REQUIRE(ws && ws_bytes >= yh_train_workspace_bytes(), YH_ERR_WORKSPACE,
               "workspace too small: %zu < %zu", ws_bytes, yh_train_workspace_bytes());
    YH_REQUIRE(((uintptr_t)y & 3) == 0 && ((uintptr_t)dy & 3) == 0 && ((uintptr_t)gt & 15) == 0,
               YH_ERR_INVALID, "misaligned pointer (y/dy need 4-byte, gt 16-byte alignment)");

    p.y = y; p.dy = dy; p.gt = gt; p.gt_off = gt_off;
    p.terms = terms; p.loss = loss; p.resp = resp; p.iou_resp = iou_resp;
    p.partials = reinterpret_cast<float*>(ws);
    p.ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + (size_t)kMaxGrid * kPartials * 4);
    p.total_cells = (long long)n * p.g.cells;
    p.quads_total = (p.total_cells + 3) / 4;
    const int cf = p.g.cell_floats;
    int qpc = kStageBytesTarget / (16 * cf);  // quads of cells per mini-chunk
    if (qpc < 1) qpc = 1;
    if (qpc > 8) qpc = 8;  // <= 32 cells per mini-chunk: the kernel tracks touched cells in one word
    p.mc = qpc * 4;
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
    p.warps = warps;

    const long long nmini_total = (p.quads_total + qpc - 1) / qpc;
    long long grid = (nmini_total + warps - 1) / warps;  // at least one mini-chunk per warp
    const int sms = yh_sm_count();
    if (grid > sms) grid = sms;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    return dy ? launch_variant<true>(p, (int)grid, smem, st) : launch_variant<false>(p, (int)grid, smem, st);
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
