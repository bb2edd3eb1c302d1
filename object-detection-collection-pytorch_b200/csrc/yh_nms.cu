// Inference post-process: confidence threshold + greedy NMS per image (+ class pick), either
// straight from the head tensor (yh_v{1,2}_postprocess) or from decoded boxes (yh_nms).
//
// Replaces models.utils.nms (reference models/utils.py:68-164) applied per image, and the
// predict -> nms -> argmax chain of detect() (reference models/yolov2.py:694-731,
// models/yolov1.py:491-534).
//
// One CTA per image, one launch for the whole batch, no sort:
//   A  threshold : every thread tests `sigmoid(to) >= conf_thre` for its predictors and appends
//                  survivors to a candidate list in shared memory (order irrelevant);
//   B  rank      : a candidate's position in the descending-confidence order is the number of
//                  candidates that beat it (ties: lower predictor index first) -> the order is
//                  unique and deterministic without a sort;
//   C  decode    : each candidate's box is decoded from its 5 logits directly into its ranked
//                  slot (same rounding sequence as the train head / predict kernels);
//   D  suppress  : tiles of 256 ranked candidates: (1) test the tile against the boxes kept so
//                  far, (2) build the intra-tile suppression bitmask with one ballot per 32
//                  pairs, (3) one warp walks the tile in order, OR-ing mask rows of kept boxes;
//   E  emit      : one warp per kept box: class softmax of its row, cls_spec = p * conf, argmax
//                  label / max score, and the box record.
// The reference's greedy rule (models/utils.py:124-158): candidate j is dropped iff an earlier
// KEPT candidate i has iou(i, j) >= iou_thre.  class_aware additionally requires equal labels.
#include <string.h>

#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 256;             // ranked candidates per suppression tile
constexpr int kTileWords = kTile / 32;
constexpr int kSmemCandCap = 2048;     // candidate lists live in shared memory up to this P


struct NmsParams {
    YhGeom g;                // head mode only
    int mode;                // 0: from the head tensor, 1: from decoded bbox/conf
    const float* y;
    const float4* bbox;      // [N,P] decoded
    const float* conf;       // [N,P]
    const int32_t* labels;   // [N,P] or NULL
    int n, p, c;
    float conf_thre, iou_thre;
    int class_aware, max_out;
    int32_t* keep_idx;
    int32_t* keep_cnt;
    float4* out_bbox;
    float* out_conf;
    float* out_cls_spec;
    int32_t* out_label;
    float* out_score;
    unsigned char* ws;       // global candidate storage when p > kSmemCandCap
    size_t ws_per_image;
};

struct CandArrays {
    float* u_conf;     // [cap] unsorted
    int32_t* u_idx;    // [cap]
    int32_t* s_idx;    // [cap] ranked
    float* s_conf;     // [cap]
    float4* s_box;     // [cap]
    int32_t* s_lab;    // [cap]
    int32_t* keep;     // [cap] ranked positions of kept boxes
};

__host__ __device__ inline size_t cand_bytes(int cap) {
    const size_t c = ((size_t)cap + 3) & ~(size_t)3;
    return c * (16 + 4 * 6);
}

__device__ __forceinline__ CandArrays carve(unsigned char* base, int cap) {
    const size_t c = ((size_t)cap + 3) & ~(size_t)3;
    CandArrays a;
    a.s_box = reinterpret_cast<float4*>(base);
    a.u_conf = reinterpret_cast<float*>(base + c * 16);
    a.u_idx = reinterpret_cast<int32_t*>(a.u_conf + c);
    a.s_idx = a.u_idx + c;
    a.s_conf = reinterpret_cast<float*>(a.s_idx + c);
    a.s_lab = reinterpret_cast<int32_t*>(a.s_conf + c);
    a.keep = a.s_lab + c;
    return a;
}

// pointer to the 5 box logits / the C class logits of predictor `idx` of image `img`
__device__ __forceinline__ const float* box_logits(const NmsParams& p, int img, int idx) {
    const YhGeom& g = p.g;
    if (g.version == 2) return p.y + ((size_t)img * g.preds + idx) * g.box_stride;
    const int cell = idx / g.a, b = idx - cell * g.a;
    return p.y + ((size_t)img * g.cells + cell) * g.cell_floats + b * 5;
}
__device__ __forceinline__ const float* cls_logits(const NmsParams& p, int img, int idx) {
    const YhGeom& g = p.g;
    if (g.version == 2) return p.y + ((size_t)img * g.preds + idx) * g.box_stride + 5;
    const int cell = idx / g.a;
    return p.y + ((size_t)img * g.cells + cell) * g.cell_floats + 5 * g.a;
}

// Warp-cooperative class pick of one predictor: softmax over its C logits, cls_spec = p * conf
// (reference models/yolov2.py:625-640), label = first argmax of cls_spec, score = its max
// (models/yolov2.py:726-731).  Optionally stores the cls_spec row.
__device__ __forceinline__ void warp_class_pick(const float* cl, int C, float conf, int lane,
                                                float* spec_out, int* label, float* score) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(cl + c));
    mx = yh_warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(__ldg(cl + c) - mx);
    se = yh_warp_sum(se);
    float bv = -INFINITY;
    int bi = 1 << 30;
    for (int c = lane; c < C; c += 32) {
        const float pc = __fdiv_rn(expf(__ldg(cl + c) - mx), se);
        const float sp = __fmul_rn(pc, conf);
        if (spec_out) spec_out[c] = sp;
        if (sp > bv || (sp != sp && bv == bv)) { bv = sp; bi = c; }  // first max per lane (c ascending)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
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

__global__ void __launch_bounds__(kThreads) yh_nms_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int mask[kTile * kTileWords];
    __shared__ unsigned int rem0[kTileWords];
    __shared__ int s_count, s_kept;

    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.p;
    const CandArrays ca = carve(p.ws ? p.ws + (size_t)img * p.ws_per_image : smem_raw, P);

    if (tid == 0) { s_count = 0; s_kept = 0; }
    __syncthreads();

    // ---------------- A: threshold ----------------
    for (int base = 0; base < P; base += kThreads) {
        const int i = base + tid;
        float conf = 0.f;
        bool pass = false;
        if (i < P) {
            conf = p.mode == 0 ? yh_sigmoid(__ldg(box_logits(p, img, i) + 4))
                               : __ldg(p.conf + (size_t)img * P + i);
            pass = conf >= p.conf_thre;  // models/utils.py:92
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            int slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(&s_count, __popc(bal));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (pass) {
                const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
                ca.u_conf[slot] = conf;
                ca.u_idx[slot] = i;
            }
        }
    }
    __syncthreads();
    const int K = s_count;

    // ---------------- B + C: rank, then decode into the ranked slot ----------------
    for (int k = tid; k < K; k += kThreads) {
        const float ck = ca.u_conf[k];
        const int ik = ca.u_idx[k];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float cj = ca.u_conf[j];
            rank += (cj > ck || (cj == ck && ca.u_idx[j] < ik)) ? 1 : 0;
        }
        ca.s_idx[rank] = ik;
        ca.s_conf[rank] = ck;
        float4 bx;
        if (p.mode == 0) {
            const YhGeom& g = p.g;
            const float* bp = box_logits(p, img, ik);
            const int cell = ik / g.a, a = ik - cell * g.a;
            const int cy = cell / g.s_w, cx = cell - cy * g.s_w;
            const float sx = yh_sigmoid(__ldg(bp + 0)), sy = yh_sigmoid(__ldg(bp + 1));
            float wa, ha;
            if (g.version == 2) {
                wa = expf(__ldg(bp + 2));
                ha = expf(__ldg(bp + 3));
            } else {
                wa = yh_sigmoid(__ldg(bp + 2));
                ha = yh_sigmoid(__ldg(bp + 3));
            }
            const YhBox b = yh_decode_box(sx, sy, wa, ha, g.pw[a], g.ph[a], cx, cy, g.gw, g.gh);
            bx = make_float4(b.x1, b.y1, b.x2, b.y2);
        } else {
            bx = __ldg(p.bbox + (size_t)img * P + ik);
        }
        ca.s_box[rank] = bx;
        if (p.mode == 1 && p.labels) ca.s_lab[rank] = __ldg(p.labels + (size_t)img * P + ik);
    }
    __syncthreads();

    const bool use_lab = p.mode == 0 ? (p.class_aware != 0) : (p.labels != nullptr);
    if (use_lab && p.mode == 0) {  // label of every candidate (argmax of cls_spec)
        for (int k = warp; k < K; k += kWarps) {
            int lab;
            float sc;
            warp_class_pick(cls_logits(p, img, ca.s_idx[k]), p.c, ca.s_conf[k], lane, nullptr, &lab, &sc);
            if (lane == 0) ca.s_lab[k] = lab;
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
        // (2) intra-tile mask: word (i, w) = candidates j in [32w, 32w+32) suppressed by i
        for (int task = warp; task < tn * W; task += kWarps) {
            const int i = task / W, w = task - i * W;
            if (w < (i >> 5)) continue;
            const int j = w * 32 + lane;
            bool bit = false;
            if (j > i && j < tn) {
                const float4 bi = ca.s_box[base + i], bj = ca.s_box[base + j];
                const YhBox qi{bi.x, bi.y, bi.z, bi.w}, qj{bj.x, bj.y, bj.z, bj.w};
                bit = yh_iou_xyxy(qi, qj) >= thr;  // survive iff iou < thr, models/utils.py:133
                if (use_lab) bit = bit && ca.s_lab[base + i] == ca.s_lab[base + j];
            }
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) mask[i * kTileWords + w] = m;
        }
        __syncthreads();
        // (3) in-order walk, one warp; lane l < W owns removed-word l
        if (warp == 0) {
            unsigned rem = lane < W ? rem0[lane] : 0u;
            int kept = s_kept;
            for (int i = 0; i < tn; ++i) {
                const unsigned word = __shfl_sync(0xffffffffu, rem, i >> 5);
                if (!((word >> (i & 31)) & 1u)) {
                    if (lane == 0) ca.keep[kept] = base + i;
                    ++kept;
                    if (lane >= (i >> 5) && lane < W) rem |= mask[i * kTileWords + lane];
                }
            }
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
    }

    // ---------------- E: emit ----------------
    const int kept = s_kept;
    if (tid == 0) p.keep_cnt[img] = kept;
    const int nout = min(kept, p.max_out);
    for (int t = warp; t < nout; t += kWarps) {
        const int i = ca.keep[t];
        const int idx = ca.s_idx[i];
        const float conf = ca.s_conf[i];
        const size_t o = (size_t)img * p.max_out + t;
        if (lane == 0) {
            p.keep_idx[o] = idx;
            if (p.out_conf) p.out_conf[o] = conf;
            if (p.out_bbox) p.out_bbox[o] = ca.s_box[i];
        }
        if (p.mode == 0 && (p.out_cls_spec || p.out_label || p.out_score)) {
            int lab;
            float sc;
            warp_class_pick(cls_logits(p, img, idx), p.c, conf, lane,
                            p.out_cls_spec ? p.out_cls_spec + o * p.c : nullptr, &lab, &sc);
            if (lane == 0) {
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
    size_t smem = 0;
    if (p.p <= kSmemCandCap) {
        smem = cand_bytes(p.p);
        p.ws = nullptr;
        p.ws_per_image = 0;
    } else {
        const size_t need = yh_postprocess_workspace_bytes(p.n, p.p);
        YH_REQUIRE(ws && ws_bytes >= need, YH_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
        YH_REQUIRE(((uintptr_t)ws & 15) == 0, YH_ERR_INVALID, "workspace must be 16-byte aligned");
        p.ws = reinterpret_cast<unsigned char*>(ws);
        p.ws_per_image = cand_bytes(p.p);
    }
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > 40 * 1024 && smem > configured[dev]) {
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
    p.mode = 0;
    p.y = y;
    p.n = n; p.p = p.g.preds; p.c = c;
    p.conf_thre = conf_thre; p.iou_thre = iou_thre;
    p.class_aware = class_aware; p.max_out = max_out;
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt;
    p.out_bbox = reinterpret_cast<float4*>(out_bbox);
    p.out_conf = out_conf; p.out_cls_spec = out_cls_spec; p.out_label = out_label; p.out_score = out_score;
    return launch(p, ws, ws_bytes, stream);
}

}  // namespace

extern "C" {

size_t yh_postprocess_workspace_bytes(int n, int preds_per_image) {
    if (n <= 0 || preds_per_image <= kSmemCandCap) return 16;
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
    p.mode = 1;
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
