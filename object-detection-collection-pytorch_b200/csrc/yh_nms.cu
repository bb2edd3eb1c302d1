// Inference post-process: confidence threshold + greedy NMS per image (+ class pick), either
// straight from the head tensor (yh_v{1,2}_postprocess) or from decoded boxes (yh_nms).
//
// Replaces models.utils.nms (reference models/utils.py:68-164) applied per image, and the
// predict -> nms -> argmax chain of detect() (reference models/yolov2.py:694-731,
// models/yolov1.py:491-534).
//
// One CTA per image, one launch for the whole batch, no sort:
//   A  stage+threshold : the image's slice of the head tensor is pulled into shared memory with
//                  1-D TMA bulk copies (16 KB stages, one mbarrier each; the whole image stays
//                  resident when it fits, otherwise a 4-stage ring is recycled and candidate rows
//                  are copied aside).  As stages land, every thread tests
//                  `sigmoid(to) >= conf_thre` for its predictors and appends survivors to a
//                  candidate list (order irrelevant).  All later phases read shared memory only.
//   B  rank      : a candidate's position in the descending-confidence order is the number of
//                  candidates that beat it (ties: lower predictor index first) -> the order is
//                  unique and deterministic without a sort;
//   C  decode    : each candidate's box is decoded from its 5 logits directly into its ranked
//                  slot (same rounding sequence as the train head / predict kernels);
//   D  suppress  : tiles of 256 ranked candidates: (1) test the tile against the boxes kept so
//                  far, (2) build the intra-tile suppression bitmask with one ballot per 32
//                  pairs, (3) one warp walks the KEPT boxes of the tile (ffs over the live bits),
//                  OR-ing their mask rows;
//   E  emit      : one warp per kept box: class softmax of its row, cls_spec = p * conf, argmax
//                  label / max score, and the box record.
// The reference's greedy rule (models/utils.py:124-158): candidate j is dropped iff an earlier
// KEPT candidate i has iou(i, j) >= iou_thre.  class_aware additionally requires equal labels.
//
// Candidate lists live in shared memory up to 256 candidates per image (the common case by a wide
// margin); images with more spill to the caller's workspace and take the same code path through
// generic pointers.
#include <math.h>
#include <string.h>

#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 256;             // ranked candidates per suppression tile
constexpr int kTileWords = kTile / 32;
constexpr int kSmemCand = 256;         // candidates held in shared memory
constexpr int kRowCache = 256;         // candidate rows copied aside in ring mode
constexpr int kStageFloats = 4096;     // 16 KB TMA stages
constexpr int kRingStages = 4;
constexpr int kMaxStages = 8;          // resident mode: image <= 8 stages (128 KB)

enum { SRC_TMA = 0, SRC_GLOBAL = 1, SRC_DECODED = 2 };

struct NmsParams {
    YhGeom g;                // head sources only
    int src;
    const float* y;
    long long total_floats;  // of the head tensor
    const float4* bbox;      // [N,P] decoded
    const float* conf;       // [N,P]
    const int32_t* labels;   // [N,P] or NULL
    int n, p, c;
    float conf_thre, iou_thre;
    float to_reject;         // objectness logits below this can never reach conf_thre
    int class_aware, max_out;
    int32_t* keep_idx;
    int32_t* keep_cnt;
    float4* out_bbox;
    float* out_conf;
    float* out_cls_spec;
    int32_t* out_label;
    float* out_score;
    unsigned char* ws;       // global candidate storage for images with > kSmemCand candidates
    size_t ws_per_image;
    // SRC_TMA staging
    int resident;            // whole image stays in shared memory
    int nst;                 // shared-memory stages
    int img_smem_floats;     // floats of shared memory set aside for the staged image
    int img_floats;          // floats per image
    int unit_floats;         // v2: 5+C (a predictor row); v1: 5B+C (a cell)
    int units;               // v2: P; v1: cells
    int row_floats;          // 5 + C
    int row_cache;           // rows cached in ring mode
};

struct Cand {
    float* u_conf;     // unsorted
    int32_t* u_idx;
    int32_t* s_slot;   // ranked -> unsorted slot
    int32_t* s_idx;    // ranked
    float* s_conf;
    float4* s_box;
    int32_t* s_lab;
    int32_t* keep;     // ranked positions of kept boxes
};

__host__ __device__ inline size_t cand_bytes(int cap) {
    const size_t c = ((size_t)cap + 3) & ~(size_t)3;
    return c * (16 + 4 * 7);
}

__device__ __forceinline__ Cand carve(unsigned char* base, int cap) {
    const size_t c = ((size_t)cap + 3) & ~(size_t)3;
    Cand a;
    a.s_box = reinterpret_cast<float4*>(base);
    a.u_conf = reinterpret_cast<float*>(base + c * 16);
    a.u_idx = reinterpret_cast<int32_t*>(a.u_conf + c);
    a.s_slot = a.u_idx + c;
    a.s_idx = a.s_slot + c;
    a.s_conf = reinterpret_cast<float*>(a.s_idx + c);
    a.s_lab = reinterpret_cast<int32_t*>(a.s_conf + c);
    a.keep = a.s_lab + c;
    return a;
}

// Class pick of one predictor by a group of 4 adjacent lanes (8 predictors per warp at a time):
// softmax over its C logits, cls_spec = p * conf (reference models/yolov2.py:625-640),
// label = first argmax of cls_spec, score = its max (models/yolov2.py:726-731).  Optionally stores
// the cls_spec row.  `cl` may point to shared or global memory; inactive groups pass active=false
// (they still take part in the shuffles).
__device__ __forceinline__ void quad_class_pick(const float* cl, int C, float conf, int sub, bool active,
                                                float* spec_out, int* label, float* score) {
    float mx = -INFINITY;
    if (active)
        for (int c = sub; c < C; c += 4) mx = fmaxf(mx, cl[c]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float se = 0.f;
    if (active)
        for (int c = sub; c < C; c += 4) se += expf(cl[c] - mx);
    se += __shfl_xor_sync(0xffffffffu, se, 1);
    se += __shfl_xor_sync(0xffffffffu, se, 2);
    float bv = -INFINITY;
    int bi = 1 << 30;
    if (active) {
        for (int c = sub; c < C; c += 4) {
            const float sp = __fmul_rn(__fdiv_rn(expf(cl[c] - mx), se), conf);
            if (spec_out) spec_out[c] = sp;
            if (sp > bv || (sp != sp && bv == bv)) { bv = sp; bi = c; }  // first max per lane (c ascending)
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool on = ov != ov, bn = bv != bv;
        bool take;
        if (on || bn) take = on && (!bn || oi < bi);
        else take = ov > bv || (ov == bv && oi < bi);
        if (take) { bv = ov; bi = oi; }
    }
    *label = bi;
    *score = bv;
}

// bit = "box i suppresses box j" (models/utils.py:133: j survives iff iou < thr).  Boxes that do
// not overlap have iou == 0 exactly, which skips the division for the vast majority of pairs.
__device__ __forceinline__ bool suppresses(const float4& bi, const float4& bj, float thr) {
    if (thr > 0.f) {
        const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
        const float ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        if (!(iw > 0.f) || !(ih > 0.f)) {
            // inter == 0 (or NaN coordinates): iou is 0, -0 or NaN, none of which reaches thr > 0
            if (iw == iw && ih == ih) return false;
        }
    }
    const YhBox qi{bi.x, bi.y, bi.z, bi.w}, qj{bj.x, bj.y, bj.z, bj.w};
    return yh_iou_xyxy(qi, qj) >= thr;
}

__global__ void __launch_bounds__(kThreads) yh_nms_kernel(const NmsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned int mask[kTile * kTileWords];
    __shared__ unsigned int rem0[kTileWords];
    __shared__ unsigned int nzrow[kTileWords];  // rows of `mask` with at least one bit set
    __shared__ __align__(8) uint64_t bars[kMaxStages];
    __shared__ int s_count, s_kept;

    const YhGeom& g = p.g;
    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.p;

    // dynamic shared memory: [image stages | candidate arrays (kSmemCand) | row cache (ring mode)]
    float* sm_img = reinterpret_cast<float*>(smem_raw);
    const size_t img_smem_floats = p.src == SRC_TMA ? (size_t)p.img_smem_floats : 0;
    unsigned char* cand_smem = smem_raw + img_smem_floats * 4;
    float* row_cache = reinterpret_cast<float*>(cand_smem + cand_bytes(kSmemCand));
    Cand ca = carve(cand_smem, kSmemCand);
    Cand cw = ca;  // workspace copy for images that overflow shared memory
    if (p.ws) cw = carve(p.ws + (size_t)img * p.ws_per_image, P);

    if (tid == 0) { s_count = 0; s_kept = 0; }

    // candidate append: warp-aggregated slot reservation; first kSmemCand slots in shared memory
    auto append = [&](bool pass, float conf, int idx) -> int {
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        int slot = -1;
        if (bal) {
            int slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(&s_count, __popc(bal));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (pass) {
                slot = slot0 + __popc(bal & ((1u << lane) - 1u));
                if (slot < kSmemCand) { ca.u_conf[slot] = conf; ca.u_idx[slot] = idx; }
                else { cw.u_conf[slot] = conf; cw.u_idx[slot] = idx; }
            }
        }
        return slot;
    };

    // ---------------- A: stage + threshold ----------------
    int shift = 0;  // float offset of the image inside the staged (16-byte aligned) window
    if (p.src == SRC_TMA) {
        const long long f_start = (long long)img * p.img_floats;
        const long long f_end = f_start + p.img_floats;
        const long long a0 = f_start & ~3ll;
        long long a1 = (f_end + 3) & ~3ll;
        const long long lim = p.total_floats & ~3ll;
        if (a1 > lim) a1 = lim;
        shift = (int)(f_start - a0);
        const int win = (int)(a1 - a0);  // floats the bulk copies bring in
        const int nstages = (win + kStageFloats - 1) / kStageFloats;
        const unsigned ring_mask = p.resident ? 0xffffffffu : (unsigned)(p.nst * kStageFloats - 1);
        auto issue = [&](int s) {  // thread 0
            const int fl = min(kStageFloats, win - s * kStageFloats);
            uint64_t* bar = &bars[s % p.nst];
            yh_mbar_expect_tx(bar, (uint32_t)fl * 4u);
            yh_bulk_load(sm_img + (size_t)(s % p.nst) * kStageFloats, p.y + a0 + (long long)s * kStageFloats,
                         (uint32_t)fl * 4u, bar);
        };
        if (tid == 0) {
            for (int s = 0; s < p.nst; ++s) yh_mbar_init(&bars[s], 1);
            yh_mbar_fence_init();
            for (int s = 0; s < min(p.nst, nstages); ++s) issue(s);
        }
        // floats past the last 16-byte boundary of the tensor (last image only, < 4 of them)
        if (tid < (int)(f_end - a1) && p.resident) sm_img[win + tid] = __ldg(p.y + a1 + tid);
        __syncthreads();

        const int UF = p.unit_floats;
        const int upp = g.version == 2 ? 1 : g.a;
        int u_begin = 0;
        for (int s = 0; s < nstages; ++s) {
            yh_mbar_wait(&bars[s % p.nst], (uint32_t)((s / p.nst) & 1));
            int u_end = p.units;
            if (s + 1 < nstages) {
                u_end = ((s + 1) * kStageFloats - shift) / UF;
                if (u_end > p.units) u_end = p.units;
            } else if (f_end > a1 && !p.resident) {
                u_end = (win - shift) / UF;  // ring mode: the unit holding the tail is read from global below
            }
            for (int base = u_begin; base < u_end; base += kThreads) {
                const int u = base + tid;
                const unsigned rel = (unsigned)(shift + u * UF);
                for (int b = 0; b < upp; ++b) {
                    float conf = 0.f;
                    bool pass = false;
                    if (u < u_end) {
                        const float to = sm_img[(rel + 5 * b + 4) & ring_mask];
                        if (!(to < p.to_reject)) {  // far below the threshold: sigmoid not needed
                            conf = yh_sigmoid(to);
                            pass = conf >= p.conf_thre;  // models/utils.py:92
                        }
                    }
                    const int slot = append(pass, conf, u * upp + b);
                    if (!p.resident && slot >= 0 && slot < p.row_cache) {  // copy the row aside
                        float* dst = row_cache + (size_t)slot * p.row_floats;
                        for (int q = 0; q < 5; ++q) dst[q] = sm_img[(rel + 5 * b + q) & ring_mask];
                        const unsigned coff = rel + (g.version == 2 ? 5 : 5 * g.a);
                        for (int q = 0; q < p.c; ++q) dst[5 + q] = sm_img[(coff + q) & ring_mask];
                    }
                }
            }
            u_begin = u_end;
            if (!p.resident && s >= 1 && s - 1 + p.nst < nstages) {
                __syncthreads();  // everyone is done with stage s-1 (rows may straddle s-1 | s)
                if (tid == 0) issue(s - 1 + p.nst);
            }
        }
        // ring mode, last image of an unaligned tensor: its final unit straight from global memory
        for (int u = u_begin + tid; u < p.units; u += kThreads) {  // at most one unit
            const float* up = p.y + f_start + (long long)u * UF;
            for (int b = 0; b < upp; ++b) {
                const float conf = yh_sigmoid(__ldg(up + 5 * b + 4));
                if (conf >= p.conf_thre) {
                    const int slot = atomicAdd(&s_count, 1);
                    if (slot < kSmemCand) { ca.u_conf[slot] = conf; ca.u_idx[slot] = u * upp + b; }
                    else { cw.u_conf[slot] = conf; cw.u_idx[slot] = u * upp + b; }
                    // slot >= row_cache semantics: rows of such slots are re-read from global memory
                    if (slot < p.row_cache) {
                        float* dst = row_cache + (size_t)slot * p.row_floats;
                        for (int q = 0; q < 5; ++q) dst[q] = __ldg(up + 5 * b + q);
                        const float* cp = up + (g.version == 2 ? 5 : 5 * g.a);
                        for (int q = 0; q < p.c; ++q) dst[5 + q] = __ldg(cp + q);
                    }
                }
            }
        }
    } else {
        __syncthreads();
        for (int base = 0; base < P; base += kThreads) {
            const int i = base + tid;
            float conf = 0.f;
            bool pass = false;
            if (i < P) {
                if (p.src == SRC_GLOBAL) {
                    const float* bp = g.version == 2
                        ? p.y + ((size_t)img * g.preds + i) * g.box_stride
                        : p.y + ((size_t)img * g.cells + i / g.a) * g.cell_floats + (i % g.a) * 5;
                    conf = yh_sigmoid(__ldg(bp + 4));
                } else {
                    conf = __ldg(p.conf + (size_t)img * P + i);
                }
                pass = conf >= p.conf_thre;
            }
            append(pass, conf, i);
        }
    }
    __syncthreads();
    const int K = s_count;
    if (K > kSmemCand) {  // overflow: continue in the workspace arrays
        for (int k = tid; k < kSmemCand; k += kThreads) { cw.u_conf[k] = ca.u_conf[k]; cw.u_idx[k] = ca.u_idx[k]; }
        ca = cw;
        __syncthreads();
    }

    // pointers to the 5 box logits / C class logits of unsorted candidate `slot` (predictor idx)
    auto box_ptr = [&](int slot, int idx) -> const float* {
        if (p.src == SRC_TMA) {
            if (p.resident) {
                return g.version == 2 ? sm_img + shift + idx * p.unit_floats
                                      : sm_img + shift + (idx / g.a) * p.unit_floats + (idx % g.a) * 5;
            }
            if (slot < p.row_cache) return row_cache + (size_t)slot * p.row_floats;
        }
        return g.version == 2 ? p.y + ((size_t)img * g.preds + idx) * g.box_stride
                              : p.y + ((size_t)img * g.cells + idx / g.a) * g.cell_floats + (idx % g.a) * 5;
    };
    auto cls_ptr = [&](int slot, int idx) -> const float* {
        if (p.src == SRC_TMA) {
            if (p.resident) {
                return g.version == 2 ? sm_img + shift + idx * p.unit_floats + 5
                                      : sm_img + shift + (idx / g.a) * p.unit_floats + 5 * g.a;
            }
            if (slot < p.row_cache) return row_cache + (size_t)slot * p.row_floats + 5;
        }
        return g.version == 2 ? p.y + ((size_t)img * g.preds + idx) * g.box_stride + 5
                              : p.y + ((size_t)img * g.cells + idx / g.a) * g.cell_floats + 5 * g.a;
    };

    // ---------------- B + C: rank, then decode into the ranked slot ----------------
    for (int k = tid; k < K; k += kThreads) {
        const float ck = ca.u_conf[k];
        const int ik = ca.u_idx[k];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float cj = ca.u_conf[j];
            rank += (cj > ck || (cj == ck && ca.u_idx[j] < ik)) ? 1 : 0;
        }
        ca.s_slot[rank] = k;
        ca.s_idx[rank] = ik;
        ca.s_conf[rank] = ck;
        float4 bx;
        if (p.src != SRC_DECODED) {
            const float* bp = box_ptr(k, ik);
            const int cell = ik / g.a, a = ik - cell * g.a;
            const int cy = cell / g.s_w, cx = cell - cy * g.s_w;
            const float sx = yh_sigmoid(bp[0]), sy = yh_sigmoid(bp[1]);
            float wa, ha;
            if (g.version == 2) {
                wa = expf(bp[2]);
                ha = expf(bp[3]);
            } else {
                wa = yh_sigmoid(bp[2]);
                ha = yh_sigmoid(bp[3]);
            }
            const YhBox b = yh_decode_box(sx, sy, wa, ha, g.pw[a], g.ph[a], cx, cy, g.gw, g.gh);
            bx = make_float4(b.x1, b.y1, b.x2, b.y2);
        } else {
            bx = __ldg(p.bbox + (size_t)img * P + ik);
            if (p.labels) ca.s_lab[rank] = __ldg(p.labels + (size_t)img * P + ik);
        }
        ca.s_box[rank] = bx;
    }
    __syncthreads();

    const bool use_lab = p.src != SRC_DECODED ? (p.class_aware != 0) : (p.labels != nullptr);
    const int sub = tid & 3;
    if (use_lab && p.src != SRC_DECODED) {  // label of every candidate (argmax of cls_spec)
        for (int k0 = 0; k0 < K; k0 += kThreads / 4) {
            const int k = k0 + (tid >> 2);
            const bool act = k < K;
            int lab;
            float sc;
            quad_class_pick(act ? cls_ptr(ca.s_slot[k], ca.s_idx[k]) : nullptr, p.c, act ? ca.s_conf[k] : 0.f, sub,
                            act, nullptr, &lab, &sc);
            if (act && sub == 0) ca.s_lab[k] = lab;
        }
        __syncthreads();
    }

    // ---------------- D: greedy suppression, tile by tile ----------------
    const float thr = p.iou_thre;
    for (int base = 0; base < K; base += kTile) {
        const int tn = min(kTile, K - base);
        const int W = (tn + 31) >> 5;
        // (1) against the boxes kept in earlier tiles
        {
            bool dead = false;
            if (tid < tn && base > 0) {
                const float4 bj = ca.s_box[base + tid];
                const YhBox qj{bj.x, bj.y, bj.z, bj.w};
                const int lj = use_lab ? ca.s_lab[base + tid] : 0;
                const int kept = s_kept;
                for (int q = 0; q < kept; ++q) {
                    const int i = ca.keep[q];
                    const float4 bi = ca.s_box[i];
                    const YhBox qi{bi.x, bi.y, bi.z, bi.w};
                    if (yh_iou_xyxy(qi, qj) >= thr && (!use_lab || ca.s_lab[i] == lj)) { dead = true; break; }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, dead);
            if (lane == 0) rem0[warp] = bal;  // kThreads == kTile: warp w covers word w
        }
        if (tid < kTileWords) nzrow[tid] = 0u;
        __syncthreads();
        // (2) intra-tile mask: word (i, w) = candidates j in [32w, 32w+32) suppressed by i; each
        //     warp takes rows i = warp, warp + 8, ... and walks the words to the right of i
        for (int i = warp; i < tn; i += kWarps) {
            const float4 bi = ca.s_box[base + i];
            const int li = use_lab ? ca.s_lab[base + i] : 0;
            unsigned any = 0u;
            for (int w = i >> 5; w < W; ++w) {
                const int j = w * 32 + lane;
                bool bit = false;
                if (j > i && j < tn) {
                    bit = suppresses(bi, ca.s_box[base + j], thr);
                    if (use_lab) bit = bit && li == ca.s_lab[base + j];
                }
                const unsigned m = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) mask[i * kTileWords + w] = m;
                any |= m;
            }
            if (any && lane == 0) atomicOr(&nzrow[i >> 5], 1u << (i & 31));
        }
        __syncthreads();
        // (3) one warp resolves the greedy order: only rows that are still alive AND suppress
        //     something need a sequential step; lane l < W owns removed-word l
        if (warp == 0) {
            unsigned rem = lane < W ? rem0[lane] : 0xffffffffu;
            if (lane == W - 1 && (tn & 31)) rem |= ~0u << (tn & 31);  // bits past the tile end
            const unsigned nz = lane < W ? nzrow[lane] : 0u;
            for (int w = 0; w < W; ++w) {
                const unsigned nzw = __shfl_sync(0xffffffffu, nz, w);
                unsigned done = 0u;
                while (true) {
                    const unsigned cur = __shfl_sync(0xffffffffu, rem, w);
                    const unsigned todo = ~cur & nzw & ~done;
                    if (!todo) break;
                    const int b = __ffs(todo) - 1;
                    done |= 1u << b;
                    const int i = w * 32 + b;
                    if (lane >= w && lane < W) rem |= mask[i * kTileWords + lane];
                }
            }
            if (lane < W) rem0[lane] = rem;
        }
        __syncthreads();
        // survivors of the tile, in rank order, appended to the keep list
        {
            int before = s_kept, total = 0;
            for (int w = 0; w < W; ++w) {
                const int cnt = __popc(~rem0[w]);
                if (w < (tid >> 5)) before += cnt;
                total += cnt;
            }
            if (tid < tn) {
                const unsigned word = rem0[tid >> 5];
                if (!((word >> (tid & 31)) & 1u)) ca.keep[before + __popc(~word & ((1u << (tid & 31)) - 1u))] = base + tid;
            }
            __syncthreads();
            if (tid == 0) s_kept += total;
            __syncthreads();
        }
    }

    // ---------------- E: emit ----------------
    const int kept = s_kept;
    if (tid == 0) p.keep_cnt[img] = kept;
    const int nout = min(kept, p.max_out);
    const bool want_cls = p.src != SRC_DECODED && (p.out_cls_spec || p.out_label || p.out_score);
    for (int t0 = 0; t0 < nout; t0 += kThreads / 4) {
        const int t = t0 + (tid >> 2);
        const bool act = t < nout;
        int i = 0, idx = 0;
        float conf = 0.f;
        size_t o = 0;
        if (act) {
            i = ca.keep[t];
            idx = ca.s_idx[i];
            conf = ca.s_conf[i];
            o = (size_t)img * p.max_out + t;
            if (sub == 0) {
                p.keep_idx[o] = idx;
                if (p.out_conf) p.out_conf[o] = conf;
            } else if (sub == 1) {
                if (p.out_bbox) p.out_bbox[o] = ca.s_box[i];
            }
        }
        if (want_cls) {
            int lab;
            float sc;
            quad_class_pick(act ? cls_ptr(ca.s_slot[i], idx) : nullptr, p.c, conf, sub, act,
                            (act && p.out_cls_spec) ? p.out_cls_spec + o * p.c : nullptr, &lab, &sc);
            if (act && sub == 2) {
                if (p.out_label) p.out_label[o] = lab;
                if (p.out_score) p.out_score[o] = sc;
            }
        }
    }
}

int launch(NmsParams& p, void* ws, size_t ws_bytes, void* stream) {
    YH_REQUIRE(p.n > 0 && p.p > 0, YH_ERR_INVALID, "n and predictors per image must be positive");
    YH_REQUIRE(p.max_out >= 0, YH_ERR_INVALID, "max_out < 0");
    YH_REQUIRE(p.keep_cnt && (p.max_out == 0 || p.keep_idx), YH_ERR_INVALID, "keep_idx / keep_cnt is NULL");
    YH_REQUIRE(((uintptr_t)p.out_bbox & 15) == 0, YH_ERR_INVALID, "out_bbox must be 16-byte aligned");
    p.ws = nullptr;
    p.ws_per_image = 0;
    if (p.p > kSmemCand) {  // an image may produce more candidates than shared memory holds
        const size_t need = yh_postprocess_workspace_bytes(p.n, p.p);
        YH_REQUIRE(ws && ws_bytes >= need, YH_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
        YH_REQUIRE(((uintptr_t)ws & 15) == 0, YH_ERR_INVALID, "workspace must be 16-byte aligned");
        p.ws = reinterpret_cast<unsigned char*>(ws);
        p.ws_per_image = cand_bytes(p.p);
    }
    size_t smem = cand_bytes(kSmemCand);
    if (p.src == SRC_TMA) {
        smem += (size_t)p.img_smem_floats * 4;
        if (!p.resident) smem += (size_t)p.row_cache * p.row_floats * 4;
    }
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > 32 * 1024 && smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(nms)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    yh_nms_kernel<<<(unsigned)p.n, kThreads, smem, (cudaStream_t)stream>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_nms launch");
}

int postprocess_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
                     const float* anchors_wh_host, float img_h, float img_w, float conf_thre,
                     float iou_thre, int class_aware, int max_out, int32_t* keep_idx, int32_t* keep_cnt,
                     float* out_bbox, float* out_conf, float* out_cls_spec, int32_t* out_label,
                     float* out_score, void* ws, size_t ws_bytes, void* stream) {
    NmsParams p;
    memset(&p, 0, sizeof(p));
    int rc = yh_make_geom(&p.g, version, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w);
    if (rc) return rc;
    YH_REQUIRE(y, YH_ERR_INVALID, "y is NULL");
    YH_REQUIRE(((uintptr_t)y & 3) == 0, YH_ERR_INVALID, "y must be 4-byte aligned");
    p.y = y;
    p.n = n; p.p = p.g.preds; p.c = c;
    p.conf_thre = conf_thre; p.iou_thre = iou_thre;
    // sigmoid(t) >= thr needs t >= logit(thr) up to a few ulp; reject only with a wide margin
    if (!(conf_thre > 0.f)) p.to_reject = -INFINITY;          // everything passes (or thr is NaN)
    else if (conf_thre > 1.f) p.to_reject = INFINITY;         // nothing can pass
    else {
        const double t = conf_thre < 0.9999999 ? log((double)conf_thre / (1.0 - (double)conf_thre)) : 16.0;
        p.to_reject = (float)((t < 16.0 ? t : 16.0) - 0.01);
    }
    p.class_aware = class_aware; p.max_out = max_out;
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt;
    p.out_bbox = reinterpret_cast<float4*>(out_bbox);
    p.out_conf = out_conf; p.out_cls_spec = out_cls_spec; p.out_label = out_label; p.out_score = out_score;

    p.img_floats = p.g.cells * p.g.cell_floats;
    p.total_floats = (long long)n * p.img_floats;
    p.unit_floats = version == 2 ? p.g.box_stride : p.g.cell_floats;
    p.units = version == 2 ? p.g.preds : p.g.cells;
    p.row_floats = 5 + c;
    // staged (TMA) source needs a 16-byte aligned tensor and units that fit one stage
    p.src = SRC_GLOBAL;
    if (((uintptr_t)y & 15) == 0 && p.unit_floats * 2 <= kStageFloats) {
        p.src = SRC_TMA;
        const int stages_needed = (p.img_floats + 3 + 4 + kStageFloats - 1) / kStageFloats;
        if (stages_needed <= 6) {  // <= 96 KB: keep the whole image resident (two CTAs per SM)
            p.resident = 1;
            p.nst = stages_needed;
            p.img_smem_floats = (p.img_floats + 8 + 31) & ~31;  // window + tail, not whole stages
        } else {
            p.resident = 0;
            p.nst = kRingStages;
            p.img_smem_floats = kRingStages * kStageFloats;
            int rows = kRowCache;
            while (rows > 32 && (size_t)rows * p.row_floats * 4 > 32 * 1024) rows >>= 1;
            p.row_cache = rows;
        }
    }
    return launch(p, ws, ws_bytes, stream);
}

}  // namespace

extern "C" {

size_t yh_postprocess_workspace_bytes(int n, int preds_per_image) {
    if (n <= 0 || preds_per_image <= kSmemCand) return 16;
    return (size_t)n * cand_bytes(preds_per_image);
}

int yh_v2_postprocess(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                      float img_h, float img_w, float conf_thre, float iou_thre, int class_aware,
                      int max_out, int32_t* keep_idx, int32_t* keep_cnt, float* out_bbox, float* out_conf,
                      float* out_cls_spec, int32_t* out_label, float* out_score, void* ws,
                      size_t ws_bytes, void* stream) {
    return postprocess_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, conf_thre, iou_thre,
                            class_aware, max_out, keep_idx, keep_cnt, out_bbox, out_conf, out_cls_spec,
                            out_label, out_score, ws, ws_bytes, stream);
}

int yh_v1_postprocess(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                      float conf_thre, float iou_thre, int class_aware, int max_out, int32_t* keep_idx,
                      int32_t* keep_cnt, float* out_bbox, float* out_conf, float* out_cls_spec,
                      int32_t* out_label, float* out_score, void* ws, size_t ws_bytes, void* stream) {
    return postprocess_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, conf_thre, iou_thre,
                            class_aware, max_out, keep_idx, keep_cnt, out_bbox, out_conf, out_cls_spec,
                            out_label, out_score, ws, ws_bytes, stream);
}

int yh_nms(const float* bbox, const float* conf, const int32_t* labels, int n, int p_, float conf_thre,
           float iou_thre, int max_out, int32_t* keep_idx, int32_t* keep_cnt, void* ws, size_t ws_bytes,
           void* stream) {
    NmsParams p;
    memset(&p, 0, sizeof(p));
    YH_REQUIRE(bbox && conf, YH_ERR_INVALID, "bbox / conf is NULL");
    YH_REQUIRE(((uintptr_t)bbox & 15) == 0, YH_ERR_INVALID, "bbox must be 16-byte aligned");
    p.src = SRC_DECODED;
    p.bbox = reinterpret_cast<const float4*>(bbox);
    p.conf = conf;
    p.labels = labels;
    p.n = n; p.p = p_; p.c = 0;
    p.conf_thre = conf_thre; p.iou_thre = iou_thre;
    p.class_aware = labels != nullptr; p.max_out = max_out;
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt;
    return launch(p, ws, ws_bytes, stream);
}

}  // extern "C"
