// Fused YOLO train head: decode + responsible-predictor assignment + five loss terms + dL/dy
// in one pass over the head tensor.
//
// Replaces YOLOv2.get_loss / YOLOv1.get_loss and their autograd backward
// (reference models/yolov2.py:747-1140, models/yolov1.py:556-931, models/utils.py:5-65).
//
// Data movement (this kernel is HBM-bound: ~60 flop per 200 bytes):
//   * the head tensor is treated as a flat stream of cells; each persistent CTA owns a
//     contiguous range of cells and walks it in chunks of <= 16 KB;
//   * chunks are pulled into a shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier) several chunks ahead of the compute;
//   * dL/dy for the chunk is assembled in a second shared-memory ring and pushed back with
//     TMA bulk stores, so every byte of y is read once and every byte of dy written once,
//     fully coalesced, with no zero-fill pass (each row owner writes its whole row);
//   * the five partial sums go through a last-block-done reduction (fixed order ->
//     run-to-run deterministic), so the whole step is ONE launch.
//
// Work split inside a chunk:
//   phase 1: one thread per predictor row (v2) / per cell (v1): conf = sigmoid(to), the dense
//            no-object term and its gradient, zeros everywhere else;
//   phase 2: one warp per ground-truth record that falls into the chunk (records of one cell
//            always go to the same warp, in CSR order, so collisions accumulate
//            deterministically): lanes decode the A boxes of the cell, compute IoU against the
//            record, shuffle-argmax the responsible predictor, lanes then cover the C classes
//            for the softmax/class term, and the warp adds its sparse gradient rows.
#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkBytes = 16 * 1024;
constexpr int kMaxGrid = 1024;
constexpr int kPartials = 8;  // floats per CTA in the workspace (6 used)

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
    int cells_per_chunk;    // multiple of 4
    int tma_in, tma_out;    // base pointers 16-byte aligned
    float lam[5];
    double inv_den[5];      // 1/(2M), 1/(2M), 1/M, 1/(M(P-1)), 1/M
    float cxy, cwh, cconf, cno, ccls;  // gradient coefficients (see below)
};

// "a beats b" for torch.max semantics: larger wins, NaN beats everything, first index on ties
__device__ __forceinline__ bool yh_better(float va, int ia, float vb, int ib) {
    const bool na = va != va, nb = vb != vb;
    if (na || nb) return na && (!nb || ia < ib);
    return va > vb || (va == vb && ia < ib);
}

template <int STAGES_IN, int STAGES_OUT, bool WRITE_DY>
__global__ void __launch_bounds__(kThreads)
yh_train_kernel(const TrainParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const YhGeom& g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cf = g.cell_floats;
    const int chunk_floats = p.cells_per_chunk * cf;  // multiple of 4 floats

    float* in_ring = reinterpret_cast<float*>(smem_raw);
    float* out_ring = in_ring + (size_t)STAGES_IN * chunk_floats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_ring + (WRITE_DY ? (size_t)STAGES_OUT * chunk_floats : 0));
    float* red = reinterpret_cast<float*>(bars + STAGES_IN);  // [kWarps][6]

    // this CTA's contiguous cell range (in quads of cells so chunk starts stay 16-B aligned)
    const long long q0 = p.quads_total * blockIdx.x / gridDim.x;
    const long long q1 = p.quads_total * (blockIdx.x + 1) / gridDim.x;
    const long long cta_cell0 = q0 * 4;
    const long long cta_cell1 = min(q1 * 4, p.total_cells);
    const long long cta_cells = cta_cell1 - cta_cell0;
    const int nchunks = cta_cells > 0 ? (int)((cta_cells + p.cells_per_chunk - 1) / p.cells_per_chunk) : 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES_IN; ++s) yh_mbar_init(&bars[s], 1);
        yh_mbar_fence_init();
    }
    __syncthreads();

    auto issue_load = [&](int ci) {  // thread 0 only
        const long long c0 = cta_cell0 + (long long)ci * p.cells_per_chunk;
        const int nc = (int)min((long long)p.cells_per_chunk, cta_cell1 - c0);
        const uint32_t bytes = ((uint32_t)nc * cf * 4u) & ~15u;
        uint64_t* bar = &bars[ci % STAGES_IN];
        if (bytes) {
            yh_mbar_expect_tx(bar, bytes);
            yh_bulk_load(in_ring + (size_t)(ci % STAGES_IN) * chunk_floats, p.y + c0 * cf, bytes, bar);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(bar)) : "memory");
        }
    };
    if (p.tma_in && tid == 0) {
        const int pre = nchunks < STAGES_IN ? nchunks : STAGES_IN;
        for (int ci = 0; ci < pre; ++ci) issue_load(ci);
    }

    // partial sums: s_no is per thread, the rest live in lane 0 of each warp
    float s_no = 0.f, s_xy = 0.f, s_wh = 0.f, s_conf = 0.f, s_nr = 0.f, s_cls = 0.f;
    const int cells = g.cells;
    const int bs = g.box_stride;
    const int A = g.a, C = g.c;

    for (int ci = 0; ci < nchunks; ++ci) {
        const long long cell0 = cta_cell0 + (long long)ci * p.cells_per_chunk;
        const int ncell = (int)min((long long)p.cells_per_chunk, cta_cell1 - cell0);
        const int nfl = ncell * cf;
        const int nfl16 = nfl & ~3;  // floats covered by the bulk copy
        float* in = in_ring + (size_t)(ci % STAGES_IN) * chunk_floats;
        float* out = out_ring + (size_t)(ci % STAGES_OUT) * chunk_floats;
        const float* ysrc = p.y + cell0 * cf;
        const int n0 = (int)(cell0 / cells);                      // first image of the chunk
        const int rem0 = (int)(cell0 - (long long)n0 * cells);    // its first cell inside that image

        // the out slot is reused every STAGES_OUT chunks: its bulk store must have drained
        if (WRITE_DY && p.tma_out && tid == 0) yh_bulk_wait_read<STAGES_OUT - 1>();
        if (p.tma_in) {
            if (tid < nfl - nfl16) in[nfl16 + tid] = __ldg(ysrc + nfl16 + tid);  // <16-B tail
            yh_mbar_wait(&bars[ci % STAGES_IN], (uint32_t)((ci / STAGES_IN) & 1));
        } else {
            for (int i = tid; i < nfl; i += kThreads) in[i] = __ldg(ysrc + i);
        }
        __syncthreads();  // [A] chunk visible to everyone, out slot free

        // ---------------- phase 1: dense no-object term, one row owner per thread ------------
        if (g.version == 2) {
            const int nrows = ncell * A;
            for (int u = tid; u < nrows; u += kThreads) {
                const int lcell = u / A;
                const int n = n0 + (rem0 + lcell) / cells;
                const float kn = (float)(__ldg(p.gt_off + n + 1) - __ldg(p.gt_off + n));
                const float conf = yh_sigmoid(in[u * bs + 4]);
                const float c2 = conf * conf;
                s_no += kn * c2;
                if (WRITE_DY) {
                    float* row = out + u * bs;
                    for (int k = 0; k < bs; ++k) row[k] = 0.f;
                    row[4] = p.cno * kn * c2 * (1.f - conf);
                }
            }
        } else {
            for (int u = tid; u < ncell; u += kThreads) {
                const int n = n0 + (rem0 + u) / cells;
                const float kn = (float)(__ldg(p.gt_off + n + 1) - __ldg(p.gt_off + n));
                float* row = out + u * cf;
                if (WRITE_DY)
                    for (int k = 0; k < cf; ++k) row[k] = 0.f;
                for (int b = 0; b < A; ++b) {
                    const float conf = yh_sigmoid(in[u * cf + b * 5 + 4]);
                    const float c2 = conf * conf;
                    s_no += kn * c2;
                    if (WRITE_DY) row[b * 5 + 4] = p.cno * kn * c2 * (1.f - conf);
                }
            }
        }
        __syncthreads();  // [C] dense rows written before the sparse read-modify-writes

        // ---------------- phase 2: ground-truth records of this chunk, one warp each ---------
        {
            const int n_first = n0;
            const int n_last = n0 + (rem0 + ncell - 1) / cells;
            const int r0 = __ldg(p.gt_off + n_first), r1 = __ldg(p.gt_off + n_last + 1);
            for (int base = r0; base < r1; base += 32) {
                const int j = base + lane;
                int lc = -1;
                bool mine = false;
                if (j < r1) {
                    const int4 h = __ldg(reinterpret_cast<const int4*>(p.gt + j));  // img,cy,cx,cls
                    if (h.y >= 0 && h.y < g.s_h && h.z >= 0 && h.z < g.s_w) {
                        const long long gc = (long long)h.x * cells + (long long)h.y * g.s_w + h.z - cell0;
                        if (gc >= 0 && gc < ncell) {
                            lc = (int)gc;
                            mine = (lc % kWarps) == warp;
                        }
                    }
                }
                unsigned bal = __ballot_sync(0xffffffffu, mine);
                while (bal) {
                    const int b = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int jj = base + b;
                    const int lcell = __shfl_sync(0xffffffffu, lc, b);
                    const float4* gp = reinterpret_cast<const float4*>(p.gt + jj);
                    const int4 hd = __ldg(reinterpret_cast<const int4*>(gp));
                    const float4 tt = __ldg(gp + 1);  // stx, sty, tw, th
                    const float4 bb = __ldg(gp + 2);  // x1, y1, x2, y2
                    const float* cellp = in + lcell * cf;

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
                        s_xy += dx * dx + dyv * dyv;
                        s_wh += dw * dw + dh * dh;
                        s_conf += dc * dc;
                        s_nr += conf * conf;
                        if (p.resp) p.resp[jj] = r;
                        if (p.iou_resp) p.iou_resp[jj] = iou_r;
                        if (WRITE_DY) {
                            float* row = out + lcell * cf + r * bs;
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

                    // class term: softmax over C logits, lanes stride the classes
                    const int coff = g.version == 2 ? r * bs + 5 : 5 * A;
                    const float* cl = cellp + coff;
                    float mx = -INFINITY;
                    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, cl[c]);
                    mx = yh_warp_max(mx);
                    float se = 0.f;
                    for (int c = lane; c < C; c += 32) se += expf(cl[c] - mx);
                    se = yh_warp_sum(se);
                    float sq = 0.f, dot = 0.f;
                    for (int c = lane; c < C; c += 32) {
                        const float pc = __fdiv_rn(expf(cl[c] - mx), se);
                        const float gg = pc - (c == hd.w ? 1.f : 0.f);
                        sq += gg * gg;
                        dot += gg * pc;
                    }
                    sq = yh_warp_sum(sq);
                    dot = yh_warp_sum(dot);
                    if (lane == 0) s_cls += sq;
                    if (WRITE_DY) {
                        float* ocl = out + lcell * cf + coff;
                        for (int c = lane; c < C; c += 32) {
                            const float pc = __fdiv_rn(expf(cl[c] - mx), se);
                            const float gg = pc - (c == hd.w ? 1.f : 0.f);
                            ocl[c] += p.ccls * pc * (gg - dot);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();  // [D] dy chunk complete; nobody reads `in` any more

        if (WRITE_DY) {
            float* ydst = p.dy + cell0 * cf;
            if (p.tma_out) {
                if (tid == 0) {
                    yh_fence_proxy_async();
                    if (nfl16) yh_bulk_store(ydst, out, (uint32_t)nfl16 * 4u);
                    yh_bulk_commit();
                }
                if (tid < nfl - nfl16) ydst[nfl16 + tid] = out[nfl16 + tid];
            } else {
                for (int i = tid; i < nfl; i += kThreads) ydst[i] = out[i];
            }
        }
        if (p.tma_in && tid == 0 && ci + STAGES_IN < nchunks) issue_load(ci + STAGES_IN);
    }
    if (WRITE_DY && p.tma_out && tid == 0) yh_bulk_wait_all<0>();

    // ---------------- block reduction of the six partial sums ----------------
    s_no = yh_warp_sum(s_no);
    if (lane == 0) {
        float* r = red + warp * 6;
        r[0] = s_xy; r[1] = s_wh; r[2] = s_conf; r[3] = s_no; r[4] = s_nr; r[5] = s_cls;
    }
    __syncthreads();
    __shared__ bool is_last;
    if (tid == 0) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < kWarps; ++w)
            for (int k = 0; k < 6; ++k) acc[k] += red[w * 6 + k];
        float* dst = p.partials + (size_t)blockIdx.x * kPartials;
        for (int k = 0; k < 6; ++k) dst[k] = acc[k];
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
            for (int k = 0; k < 6; ++k) acc[k] += (double)__ldcg(src + k);
        }
        for (int k = 0; k < 6; ++k)
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
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

template <int SI, int SO, bool W>
int launch_variant(const TrainParams& p, int grid, size_t smem, cudaStream_t stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_train_kernel<SI, SO, W>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(train)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    yh_train_kernel<SI, SO, W><<<grid, kThreads, smem, stream>>>(p);
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
    int qpc = kChunkBytes / (16 * cf);
    if (qpc < 1) qpc = 1;
    p.cells_per_chunk = qpc * 4;
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

    const size_t chunk_bytes = (size_t)p.cells_per_chunk * cf * 4;
    const long long nchunks_total = (p.quads_total + qpc - 1) / qpc;
    const int sms = yh_sm_count();
    const size_t tail = 8 * 8 + kWarps * 6 * 4 + 64;
    cudaStream_t st = (cudaStream_t)stream;

    // ring depth: as deep as fits two CTAs per SM, shallower for very wide cells
    if (dy) {
        if ((4 + 2) * chunk_bytes + tail <= 110 * 1024) {
            int grid = (int)min((long long)sms * 2, nchunks_total);
            if (grid > kMaxGrid) grid = kMaxGrid;
            return launch_variant<4, 2, true>(p, grid, 6 * chunk_bytes + tail, st);
        }
        YH_REQUIRE((2 + 2) * chunk_bytes + tail <= 227 * 1024, YH_ERR_UNSUPPORTED,
                   "cell too wide for shared memory (%d floats per cell)", cf);
        int grid = (int)min((long long)sms, nchunks_total);
        if (grid > kMaxGrid) grid = kMaxGrid;
        return launch_variant<2, 2, true>(p, grid, 4 * chunk_bytes + tail, st);
    }
    if (6 * chunk_bytes + tail <= 110 * 1024) {
        int grid = (int)min((long long)sms * 2, nchunks_total);
        if (grid > kMaxGrid) grid = kMaxGrid;
        return launch_variant<6, 2, false>(p, grid, 6 * chunk_bytes + tail, st);
    }
    YH_REQUIRE(2 * chunk_bytes + tail <= 227 * 1024, YH_ERR_UNSUPPORTED,
               "cell too wide for shared memory (%d floats per cell)", cf);
    int grid = (int)min((long long)sms, nchunks_total);
    if (grid > kMaxGrid) grid = kMaxGrid;
    return launch_variant<2, 2, false>(p, grid, 2 * chunk_bytes + tail, st);
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
