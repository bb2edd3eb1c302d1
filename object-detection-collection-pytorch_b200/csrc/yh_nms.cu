// Inference post-process: confidence threshold + greedy NMS per image (+ class pick), either
// straight from the head tensor (yh_v{1,2}_postprocess) or from decoded boxes (yh_nms).
//
// Replaces models.utils.nms (reference models/utils.py:68-164) applied per image, and the
// predict -> nms -> argmax chain of detect() (reference models/yolov2.py:694-731,
// models/yolov1.py:491-534).
//
// One CTA per image (512 threads; 1024 where an image leaves room for one CTA per SM only), one launch for the
// whole batch, no sort.  Measured on B200: the
// kernel is DRAM-bound while the images arrive (all CTAs are resident) and bound by instruction
// issue afterwards (two CTAs per SM, each phase a few hundred nanoseconds), so it is organised
// around ONE global round trip and few warp instructions:
//   A  threshold : (whole-image mode, `IMG`: the head tensor is 16-byte aligned and an image fits
//                  in shared memory -- both reference grids) thread 0 issues four TMA bulk copies
//                  of the image, one mbarrier each; each group of four warps thresholds its piece as
//                  it lands.  `sigmoid(to) >= conf_thre` is decided on the logit outside a band of
//                  1e-3 around logit(conf_thre); survivors are listed as (logit, predictor) pairs,
//                  one shared atomic per 64 predictors; their sigmoids are taken afterwards, by
//                  as many threads, which leave 64-bit sort keys (confidence | ~predictor).
//                  (general mode: strided loads of the objectness logits, a bulk copy per
//                  candidate row; rows beyond the staged slots are read from global memory);
//   B  rank      : a candidate's position in the descending-confidence order is the number of
//                  candidates that beat it (ties: lower predictor index first) -> the order is
//                  unique and deterministic without a sort; two lanes per candidate split the
//                  comparisons and then decode one axis of the box each (same rounding sequence
//                  as the train head / predict kernels);
//   D  suppress  : the i < j pairs of the ranked candidates, enumerated densely over all threads;
//                  "i suppresses j" bits into a bitmask (rare: atomicOr); one warp resolves the
//                  greedy order by fixed-point iteration over the mask and lists the survivors;
//   E  emit      : four lanes per kept box: class softmax of its row, cls_spec = p * conf,
//                  argmax label / max score, and the box record.
// The reference's greedy rule (models/utils.py:124-158): candidate j is dropped iff an earlier
// KEPT candidate i has iou(i, j) >= iou_thre.  class_aware additionally requires equal labels.
//
// Up to 256 candidates per image everything lives in shared memory (the common case by a wide
// margin: rest_img below); images with more continue in the caller's workspace, in suppression
// tiles of 256 candidates, through generic pointers (rest below, also the path of yh_nms and of
// unaligned or oversized head tensors).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "yh_common.cuh"
#include "yh_finalize.cuh"
#include "yh_record.cuh"

namespace {

#ifndef YH_X_NMS_THREADS
#define YH_X_NMS_THREADS 512
#endif
constexpr int kThreadsFull = YH_X_NMS_THREADS;  // threads per CTA of the kernels that read the head tensor
constexpr int kTeam = 256;    // fused step: threads of a CTA that resolve the image's NMS while the others process its records
constexpr int kRecSmem = 32;  // fused step: records of the image staged in shared memory (more: read from global memory)
#ifndef YH_X_REC4_MIN
#define YH_X_REC4_MIN 8
#endif
constexpr int kRec4Min = YH_X_REC4_MIN;  // fused step: images with more records than this process four per warp at a time (yh_record.cuh)
constexpr int kTile = 256;             // ranked candidates per suppression tile
constexpr int kTileWords = kTile / 32;
constexpr int kSmemCand = 256;         // candidates held in shared memory
constexpr int kStageBytesMax = 64 * 1024;  // shared memory for staged candidate rows
constexpr int kLoadUnroll = 2;         // objectness loads in flight per thread (845 predictors = 1.65 per thread)
constexpr int kGroups = 4;             // whole-image mode: the image arrives in kGroups bulk copies, one warp group each
#ifndef YH_X_NMS_IMG_BYTES
#define YH_X_NMS_IMG_BYTES (200 * 1024)
#endif
// images up to this size are staged whole: two CTAs per SM up to ~96 KB (13x13x5x25: 84.5 KB), one beyond
// (19x19x5x25: 180.5 KB -- measured 5 % faster than staging only the candidates' rows with two CTAs per SM)
constexpr int kImgBytesMax = YH_X_NMS_IMG_BYTES;

enum { SRC_HEAD = 0, SRC_DECODED = 2 };
enum { MODE_GENERAL = 0, MODE_IMG = 1 };

#ifdef YH_X_TRACE
__device__ unsigned long long g_ntrace[4096 * 16];
__device__ __forceinline__ unsigned long long nt_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NT(slot) do { if (threadIdx.x == 0) g_ntrace[blockIdx.x * 16 + (slot)] = nt_now(); } while (0)
#define NTR(slot) do { if (threadIdx.x == kTeam) g_ntrace[blockIdx.x * 16 + (slot)] = nt_now(); } while (0)
extern "C" YH_API int yh_x_ntrace_copy(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_ntrace, (size_t)n * 8);
}
#else
#define NT(slot) do { } while (0)
#define NTR(slot) do { } while (0)
#endif

template <bool B>
struct FastTag {
    static constexpr bool value = B;
};
template <int N>
struct IntTag {
    static constexpr int value = N;
};

struct NmsParams {
    YhGeom g;                // head source only
    int src;
    const float* y;
    long long total_floats;  // of the head tensor
    const float4* bbox;      // [N,P] decoded
    const float* conf;       // [N,P]
    const int32_t* labels;   // [N,P] or NULL
    int n, p, c;
    float conf_thre, iou_thre;
    float to_reject;         // objectness logits below this can never reach conf_thre
    float to_accept;         // objectness logits from this on always reach conf_thre (+inf: no such shortcut)
    int class_aware, max_out;
    int late_wait;           // YH_POST_INPUT_READY: wait for the previous kernel at the end, not at the start
    int32_t* keep_idx;
    int32_t* keep_cnt;
    float4* out_bbox;
    float* out_conf;
    float* out_cls_spec;
    int32_t* out_label;
    float* out_score;
    unsigned char* ws;       // global candidate storage for images with > kSmemCand candidates
    size_t ws_per_image;
    int img_floats;          // floats per image
    int use_tma;             // y is 16-byte aligned: candidate rows are staged with bulk copies
    int stage_slots;         // candidate rows staged in shared memory (<= kSmemCand)
    int stage_bytes;         // shared memory in front of the candidate arrays (staged rows, or the whole image)
    int slot_floats;         // floats per staged row slot (multiple of 4)
    unsigned magic_sw;       // ceil(2^32 / s_w): cell / s_w == umulhi(cell, magic_sw) for cell < 2^16
    // ---- the fused step (TRAIN kernels): the train head's work on the image this CTA holds anyway
    float* dy;               // [N, ...] gradient (NULL: loss only)
    const YhGt* gt;          // records sorted by image
    const int32_t* gt_off;   // [N+1]
    int m_local;
    unsigned long long* acc; // the six fixed-point loss sums (train workspace)
    float cxy, cwh, cconf, cno, ccls;
    int32_t* resp;
    float* iou_resp;
};

struct Cand {
    float* u_conf;     // unsorted
    int32_t* u_idx;
    int32_t* s_slot;   // ranked -> unsorted slot
    int32_t* s_idx;    // ranked
    float* s_conf;
    float4* s_box;
    float* s_area;     // ranked: (x2-x1)*(y2-y1), the rounding yh_iou_xyxy uses
    int32_t* s_lab;
    int32_t* keep;     // ranked positions of kept boxes
};

__host__ __device__ inline size_t cand_bytes(int cap) {
    const size_t c = ((size_t)cap + 3) & ~(size_t)3;
    return c * (16 + 4 * 8);
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
    a.s_area = a.s_conf + c;
    a.s_lab = reinterpret_cast<int32_t*>(a.s_area + c);
    a.keep = a.s_lab + c;
    return a;
}

// Candidate beyond the shared-memory lists -> the image's workspace arrays (rare: kept out of line so that
// its 64-bit address arithmetic is not hoisted into the common path).
__device__ __noinline__ void overflow_put(unsigned char* ws_img, int cap, int slot, float conf, int idx) {
    const Cand cw = carve(ws_img, cap);
    cw.u_conf[slot] = conf;
    cw.u_idx[slot] = idx;
}

// Class pick of one predictor by a group of G adjacent lanes (32 / G predictors per warp at a time):
// softmax over its C logits, cls_spec = p * conf (reference models/yolov2.py:625-640),
// label = first argmax of cls_spec, score = its max (models/yolov2.py:726-731).  Optionally stores
// the cls_spec row.  `cl` may point to shared or global memory; inactive groups pass active=false
// (they still take part in the shuffles).  The exponentials of up to G*kQuadRegs classes stay in
// registers between the two passes.
constexpr int kQuadRegs = 4;
template <int G>
__device__ __forceinline__ void group_class_pick(const float* cl, int C, float conf, int sub, bool active,
                                                 float* spec_out, int* label, float* score) {
    float e[kQuadRegs];
    float mx = -INFINITY;
    if (active) {
#pragma unroll
        for (int k = 0; k < kQuadRegs; ++k) {
            const int c = sub + G * k;
            e[k] = c < C ? cl[c] : -INFINITY;
            mx = fmaxf(mx, e[k]);
        }
        for (int c = sub + G * kQuadRegs; c < C; c += G) mx = fmaxf(mx, cl[c]);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    if (active) {
#pragma unroll
        for (int k = 0; k < kQuadRegs; ++k) {
            const int c = sub + G * k;
            e[k] = c < C ? expf(e[k] - mx) : 0.f;
        }
        // (same summation order as a plain loop over c = sub, sub + 4, ...)
#pragma unroll
        for (int k = 0; k < kQuadRegs; ++k)
            if (sub + G * k < C) se += e[k];
        for (int c = sub + G * kQuadRegs; c < C; c += G) se += expf(cl[c] - mx);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    float bv = -INFINITY;
    int bi = 1 << 30;
    if (active) {
#pragma unroll
        for (int k = 0; k < kQuadRegs; ++k) {
            const int c = sub + G * k;
            if (c < C) {
                const float sp = __fmul_rn(__fdiv_rn(e[k], se), conf);
                if (spec_out) spec_out[c] = sp;
                if (sp > bv || (sp != sp && bv == bv)) { bv = sp; bi = c; }  // first max per lane (c ascending)
            }
        }
        for (int c = sub + G * kQuadRegs; c < C; c += G) {
            const float sp = __fmul_rn(__fdiv_rn(expf(cl[c] - mx), se), conf);
            if (spec_out) spec_out[c] = sp;
            if (sp > bv || (sp != sp && bv == bv)) { bv = sp; bi = c; }
        }
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
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

// Class pick for outputs that only need label and score, by 4 adjacent lanes, classes in registers (C <= 32):
// the label is the first maximum of cls_spec = softmax * conf, and only classes whose logit is within 1e-4 of
// the largest one can reach it (a smaller logit gives a cls_spec smaller by a factor 1 - 1e-4, far beyond
// the rounding of expf, the division and the product), so only those classes -- almost always one -- take the
// division.  Same sums, in the same order, as group_class_pick<4>: the score has the same bits.  *ok is false
// where that argument does not hold (non-finite sum, cls_spec below the normal range): the caller
// then runs group_class_pick for the warp.
constexpr int kPickRegs = 8;
__device__ __forceinline__ void quad_class_pick_fast(const float* cl, int C, float conf, int sub, bool active,
                                                     int* label, float* score, bool* ok) {
    float l[kPickRegs];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kPickRegs; ++k) {
        const int c = sub + 4 * k;
        l[k] = (active && c < C) ? cl[c] : -INFINITY;
        mx = fmaxf(mx, l[k]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    // this lane's class that can be the maximum (more than one: the caller takes the full pick)
    const float near_ = mx - 1e-4f;
    float se = 0.f, eb = -1.f;
    int cb = 1 << 30, ncand = 0;
#pragma unroll
    for (int k = 0; k < kPickRegs; ++k) {
        const int c = sub + 4 * k;
        if (c < C) {  // (uniform per k for the lanes of a group up to the last k)
            const float e = expf(l[k] - mx);
            se += e;
            if (l[k] >= near_) { eb = e; cb = c; ++ncand; }
        }
    }
    se += __shfl_xor_sync(0xffffffffu, se, 1);
    se += __shfl_xor_sync(0xffffffffu, se, 2);
    float bv = ncand ? __fmul_rn(__fdiv_rn(eb, se), conf) : -INFINITY;
    int bi = cb;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    *label = bi;
    *score = bv;
    *ok = !active || (ncand <= 1 && se < INFINITY && bv >= 1e-30f && bv < INFINITY);  // (false for NaN)
}

// bit = "box i suppresses box j": NOT (iou(i, j) < thr) with the reference's arithmetic
// (models/utils.py:47-63, 133: j survives iff iou < thr; a NaN IoU therefore removes j).  ai/aj are the boxes' areas as yh_iou_xyxy
// rounds them.  Boxes that do not overlap have iou == 0 exactly, and away from the threshold the
// comparison is decided by one multiplication (margin 2^-20 >> the rounding of the product); the
// IEEE division only runs for ratios within that margin of thr, so the result is bit-exact.
__device__ __forceinline__ bool suppresses(const float4& bi, float ai, const float4& bj, float aj, float thr) {
    const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
    const float ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
    if (thr > 0.f && (!(iw > 0.f) || !(ih > 0.f)) && iw == iw && ih == ih) return false;  // inter == 0 -> iou is 0 / -0
    const float inter = __fmul_rn(fmaxf(iw, 0.0f), fmaxf(ih, 0.0f));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(ai, aj), inter), 1e-6f);
    const float t = __fmul_rn(thr, den);
    if (den > 0.f && thr > 0.f && t < INFINITY) {
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    return !(__fdiv_rn(inter, den) < thr);  // (a NaN ratio -- two boxes of infinite size -- does NOT survive: utils.py:133)
}

// The same decision for the dense pair enumeration, where nearly every pair is far below the threshold:
// with thr > 0 (`thr_pos`, uniform) a pair whose intersection is below 0.999999 * thr * (union + 1e-6) cannot
// reach thr (a union that is not positive, or NaNs anywhere, fail this test and take the full one).
__device__ __forceinline__ bool suppresses_dense(const float4& bi, float ai, const float4& bj, float aj, float thr, bool thr_pos) {
    const float iw = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
    const float ih = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
    const float inter = __fmul_rn(fmaxf(iw, 0.0f), fmaxf(ih, 0.0f));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(ai, aj), inter), 1e-6f);
    if (thr_pos && inter < __fmul_rn(__fmul_rn(thr, den), 0.999999f)) return false;
    return suppresses(bi, ai, bj, aj, thr);
}

// TV/TA/TC != 0 fix version / boxes per cell / classes at compile time (index arithmetic folds,
// divisions become multiplies); 0 keeps them as run-time values from the geometry.
// IMG: the image's whole slice of the head tensor is staged in shared memory by kGroups bulk copies
// issued at the start (ONE global round trip, no per-candidate copies); otherwise the objectness
// logits are read with strided loads and only the candidates' rows are staged.
// TRAIN != 0 (the fused step, yh_v2_train_post; whole-image mode, v2; TRAIN == 2: with the four-records-per-warp form): the CTA that holds an image in shared memory for
// the post-process ALSO does the train head's work on it -- y is read once per step, by one kernel:
//   * as each piece of the image lands, its warp group first lists the piece's NMS candidates (arriving on bar_list),
//     then runs the dense pass over it: the no-object term and its gradient per objectness logit, dL/dy written with
//     16-byte stores straight from registers (zero elsewhere), and arrives on bar_dense;
//   * the first 256 threads -- their pieces land first -- resolve the image's NMS as soon as the list is complete
//     (keys, rank, decode, pair tests, greedy order, class pick, emit: the lean path below on a 256-thread team),
//     i.e. while the last pieces' gradients are still being written; the other warps wait for the complete dense
//     gradient (bar_dense) and process the image's ground-truth records on top of it (one warp per record, or four
//     records per warp, dealt by cell so that records sharing a cell accumulate in CSR order; yh_record.cuh -- the
//     same arithmetic, the same bits as the train head);
//   * the CTA's six loss sums go onto the 64-bit fixed-point accumulators; yh_train_finalize_kernel follows.
template <int TV, int TA, int TC, int MODE, int NTH, int TRAIN>
__global__ void __launch_bounds__(NTH, NTH >= 1024 ? 1 : (NTH >= 512 ? 2 : 3)) yh_nms_kernel(const NmsParams p) {
    constexpr int kThreads = NTH, kWarps = NTH / 32;
    constexpr bool IMG = MODE == MODE_IMG;
    static_assert(!TRAIN || (MODE == MODE_IMG && NTH > kTeam), "the fused step runs on the whole-image kernel");
    __shared__ float red[TRAIN ? kWarps * 6 : 1];                       // TRAIN: the warps' six loss sums
    __shared__ __align__(16) int4 s_rec[TRAIN ? 3 * kRecSmem : 1];      // TRAIN: the image's first records
    __shared__ __align__(8) uint64_t bar_rec;                           // TRAIN: ... have landed
    extern __shared__ __align__(128) unsigned char smem_raw[];  // [staged rows or image | candidate arrays]
    __shared__ __align__(16) unsigned int mask[kTile * kTileWords];  // (zeroed with 16-byte stores)
    __shared__ unsigned int rem0[kTileWords];
    __shared__ __align__(8) uint64_t bar;      // staged rows have landed (transaction bytes)
    __shared__ __align__(8) uint64_t bar_list; // every thread has listed its candidates
    __shared__ __align__(8) uint64_t bar_dense; // TRAIN: every thread has written its share of the dense dL/dy
    __shared__ __align__(8) uint64_t bar_img[kGroups];  // IMG: one per bulk copy of the image
    __shared__ int s_count, s_kept;

    const YhGeom& g = p.g;
    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.p;
    const bool head = p.src == SRC_HEAD;
    const int A = TA ? TA : g.a, C = TC ? TC : p.c;
    const bool v2 = (TV ? TV : g.version) == 2;
    const int bs = v2 ? 5 + C : 5, cf = v2 ? A * (5 + C) : 5 * A + C;

    float* stage = reinterpret_cast<float*>(smem_raw);
    unsigned char* cand_smem = smem_raw + p.stage_bytes;
    Cand ca = carve(cand_smem, kSmemCand);
    // workspace copy for images that overflow shared memory (carved where it is needed: the common case never is)
    auto carve_ws = [&]() -> Cand { return p.ws ? carve(p.ws + (size_t)img * p.ws_per_image, P) : ca; };

    NT(0);
    if (IMG) {  // the single-tile path sets suppression bits with atomicOr
        for (int q = tid; q < kTile * kTileWords / 4; q += kThreads) reinterpret_cast<uint4*>(mask)[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        s_count = 0;
        s_kept = 0;
        yh_mbar_init(&bar, kThreads);
        yh_mbar_init(&bar_list, kThreads);
        if (TRAIN) yh_mbar_init(&bar_dense, kThreads);
        if (IMG) {
#pragma unroll
            for (int q = 0; q < kGroups; ++q) yh_mbar_init(&bar_img[q], 1);
        }
        if (TRAIN) yh_mbar_init(&bar_rec, 1);
        yh_mbar_fence_init();
    }
    // (programmatic dependent launch: the prologue above overlaps the previous kernel's tail; global
    // memory is only touched once that kernel has completed)
    if (!p.late_wait) yh_grid_dependency_wait();
    yh_grid_launch_dependents();
    // IMG: the image's aligned window [image start - fsh, ...) goes to shared memory in kGroups pieces cut at
    // unit boundaries (v2: predictors, v1: cells) rounded down to 16 bytes; window float w is image float w - fsh
    const int img_units = (TV ? TV : p.g.version) == 2 ? p.p : p.g.cells;
    const int img_unit_floats = (TV ? TV : p.g.version) == 2 ? 5 + (TC ? TC : p.c) : p.g.cell_floats;
    if (IMG && tid == 0) {
        const int fsh0 = (int)(((long long)img * p.img_floats) & 3);
        const float* wsrc = p.y + ((long long)img * p.img_floats - fsh0);
        const long long limw = (p.total_floats & ~3ll) - ((long long)img * p.img_floats - fsh0);  // bulk copies end here
        const int wend = (int)min((long long)((fsh0 + p.img_floats + 3) & ~3), limw);
#pragma unroll
        for (int q = 0; q < kGroups; ++q) {
            const int lo = q == 0 ? 0 : (fsh0 + ((img_units * q) / kGroups) * img_unit_floats) & ~3;
            const int hi = q == kGroups - 1 ? wend : min(wend, (fsh0 + ((img_units * (q + 1)) / kGroups) * img_unit_floats) & ~3);
            if (hi > lo) {
                yh_mbar_expect_tx(&bar_img[q], (uint32_t)(hi - lo) * 4u);
                yh_bulk_load(reinterpret_cast<float*>(smem_raw) + lo, wsrc + lo, (uint32_t)(hi - lo) * 4u, &bar_img[q]);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar_img[q])) : "memory");
            }
        }
    }
    // TRAIN: the image's record range (every thread: the dense pass needs the box count), and its first records
    // into shared memory with one bulk copy -- both round trips are hidden behind the arrival of the image
    int o0 = 0, o1 = 0;
    WarpSums sums = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (TRAIN) {
        o0 = __ldg(p.gt_off + img);
        o1 = min(__ldg(p.gt_off + img + 1), p.m_local);
    }
    __syncthreads();
    if (TRAIN && tid == kThreads - 32) {
        const int nsm = min(o1 - o0, kRecSmem);
        if (nsm > 0) {
            yh_mbar_expect_tx(&bar_rec, (uint32_t)nsm * 48u);
            yh_bulk_load(s_rec, p.gt + o0, (uint32_t)nsm * 48u, &bar_rec);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar_rec)) : "memory");
        }
    }
    const float kn = (float)(o1 - o0);

    // float offsets, inside the image, of a predictor's 5 box logits and C class logits (32-bit: one
    // image of the head tensor is far below 2^31 floats); `fsh` is the image's own shift inside its
    // 16-byte window, so (fsh + offset) & 3 is a logit's shift inside ITS window
    const float* yimg = p.y + (long long)img * p.img_floats;
    const int fsh = (int)(((long long)img * p.img_floats) & 3);
    auto box_off = [&](int idx) -> int { return v2 ? idx * bs : (idx / A) * cf + (idx % A) * 5; };
    auto cls_off = [&](int idx) -> int { return v2 ? idx * bs + 5 : (idx / A) * cf + 5 * A; };
    // where the logits of unsorted candidate `slot` (predictor idx) can be read: its staged row, or
    // global memory for candidates beyond the staged slots
    auto box_ptr = [&](int slot, int idx) -> const float* {
        const int f = box_off(idx);
        if (slot < p.stage_slots) return stage + (size_t)slot * p.slot_floats + ((fsh + f) & 3);
        return yimg + f;
    };
    auto cls_ptr = [&](int slot, int idx) -> const float* {
        if (slot < p.stage_slots) {
            if (v2) return stage + (size_t)slot * p.slot_floats + ((fsh + box_off(idx)) & 3) + 5;
            return stage + (size_t)slot * p.slot_floats + 8 + ((fsh + cls_off(idx)) & 3);
        }
        return yimg + cls_off(idx);
    };
    // stage `len` floats starting at image offset f into dst (+ the shift of f inside its 16-byte
    // window): one bulk copy of the aligned window; returns the bytes the mbarrier has to expect
    const long long lim4 = (p.total_floats & ~3ll) - (long long)img * p.img_floats;  // (relative to the image)
    auto stage_span = [&](float* dst, int f, int len) -> uint32_t {
        const int shift = (fsh + f) & 3;
        const int win = (shift + len + 3) & ~3;
        if (p.use_tma && (long long)(f - shift + win) <= lim4) {
            yh_bulk_load(dst, yimg + (f - shift), (uint32_t)win * 4u, &bar);
            return (uint32_t)win * 4u;
        }
        for (int q = 0; q < len; ++q) dst[shift + q] = __ldg(yimg + f + q);  // unaligned tensor / its very end
        return 0u;
    };

    NT(1);
    uint32_t tx = 0;
    if (IMG) {
        // ---------------- A (whole image): each warp group thresholds its piece as it lands ----------------
        const float* win = reinterpret_cast<const float*>(smem_raw);
        constexpr int kGroupThreads = kThreads / kGroups;
        const int grp = tid / kGroupThreads, gtid = tid - grp * kGroupThreads;
        const int upp = v2 ? 1 : A;  // predictors per unit
        const int i_lo = ((img_units * grp) / kGroups) * upp, i_hi = ((img_units * (grp + 1)) / kGroups) * upp;
        if (img == p.n - 1) {  // floats past the last whole 16 bytes of the tensor: plain loads
            const int wend = (int)min((long long)(fsh + p.img_floats), lim4 + fsh);
            if (tid < fsh + p.img_floats - wend) stage[wend + tid] = __ldg(yimg + (wend - fsh) + tid);
        }
        // one warp of the group polls the mbarrier, the others sleep on a hardware barrier (a polling warp
        // takes issue slots from the CTA next door); their own test afterwards succeeds at once and makes
        // the bulk copy's bytes visible to them
        if (gtid < 32) yh_mbar_wait(&bar_img[grp], 0);
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kGroupThreads) : "memory");
        yh_mbar_wait(&bar_img[grp], 0);
        auto dense_pass = [&]() {
            // ---- dense pass over this group's piece (the train head's, yh_train.cu: same helpers, same bits): window
            // floats 4*i4 .. 4*i4+3 are image floats f0 .. f0+3 and sit at positions m .. m+3 of a predictor row; the
            // objectness logit (position 4) is among them iff 1 <= m <= 4.  dL/dy is zero but the objectness channel.
            const float4* win4 = reinterpret_cast<const float4*>(smem_raw);
            float4* dwin4 = p.dy ? reinterpret_cast<float4*>(p.dy + ((long long)img * p.img_floats - fsh)) : nullptr;
            const int q4lo = grp == 0 ? 0 : (fsh + ((img_units * grp) / kGroups) * img_unit_floats) >> 2;
            const int q4hi = grp == kGroups - 1 ? (fsh + p.img_floats + 3) >> 2
                                                : (fsh + ((img_units * (grp + 1)) / kGroups) * img_unit_floats) >> 2;
            const int mstep = (4 * kGroupThreads) % bs;
            int f0 = 4 * (q4lo + gtid) - fsh;
            int m = (f0 + bs) % bs;
            for (int i4 = q4lo + gtid; i4 < q4hi; i4 += kGroupThreads, f0 += 4 * kGroupThreads) {
                if (f0 >= 0 && f0 + 3 < p.img_floats) {
                    const float4 v = win4[i4];
                    const bool has = (unsigned)(m - 1) < 4u;
                    float tl = v.w;
                    tl = m == 2 ? v.z : tl;
                    tl = m == 3 ? v.y : tl;
                    tl = m == 4 ? v.x : tl;
                    float conf;
                    float w = noobj_term(tl, kn, &conf);
                    w = has ? w : 0.f;
                    sums.no += w;
                    const float val = noobj_grad(w, conf, p.cno);
                    float4 o;
                    o.x = m == 4 ? val : 0.f;
                    o.y = m == 3 ? val : 0.f;
                    o.z = m == 2 ? val : 0.f;
                    o.w = m == 1 ? val : 0.f;
                    if (dwin4) dwin4[i4] = o;
                } else {
                    // the image's first / last 16 bytes shared with its neighbours (images are not multiples of 16 bytes)
                    for (int j = 0; j < 4; ++j) {
                        const int f = f0 + j;
                        if (f < 0 || f >= p.img_floats) continue;
                        float o = 0.f;
                        if (f % bs == 4) {
                            float conf;
                            const float w = noobj_term(win[4 * i4 + j], kn, &conf);
                            sums.no += w;
                            o = noobj_grad(w, conf, p.cno);
                        }
                        if (p.dy) p.dy[(long long)img * p.img_floats + f] = o;
                    }
                }
                m += mstep;
                m = m >= bs ? m - bs : m;
            }
        };
        // conf >= conf_thre is decided on the logit wherever that is safe (to_reject / to_accept leave a
        // band around logit(conf_thre) in which the sigmoid is evaluated); survivors are listed with
        // their LOGIT -- the sigmoid of the ~50 survivors is taken later by as many threads, instead
        // of here by every warp that holds one
        for (int base = i_lo; base < i_hi; base += 2 * kGroupThreads) {  // (warp-uniform trip count)
            const int i0 = base + gtid, i1 = i0 + kGroupThreads;
            float t0 = __int_as_float(0x7fc00000), t1 = t0;  // (idle lanes: NaN fails every test below)
            if (i0 < i_hi) t0 = win[fsh + box_off(i0) + 4];
            if (i1 < i_hi) t1 = win[fsh + box_off(i1) + 4];
            bool pass0 = t0 >= p.to_accept, pass1 = t1 >= p.to_accept;
            if (!pass0 && t0 >= p.to_reject) pass0 = yh_sigmoid(t0) >= p.conf_thre;  // models/utils.py:92
            if (!pass1 && t1 >= p.to_reject) pass1 = yh_sigmoid(t1) >= p.conf_thre;
            const unsigned bal0 = __ballot_sync(0xffffffffu, pass0), bal1 = __ballot_sync(0xffffffffu, pass1);
            if (bal0 | bal1) {
                int slot0 = 0;
                if (lane == 0)
                    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(slot0) : "r"(yh_smem_u32(&s_count)), "r"(__popc(bal0) + __popc(bal1)) : "memory");
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                const unsigned below = (1u << lane) - 1u;
                if (pass0) {
                    const int slot = slot0 + __popc(bal0 & below);
                    if (slot < kSmemCand) reinterpret_cast<int2*>(ca.u_conf)[slot] = make_int2(__float_as_int(t0), i0);
                    else overflow_put(p.ws + (size_t)img * p.ws_per_image, P, slot, t0, i0);
                }
                if (pass1) {
                    const int slot = slot0 + __popc(bal0) + __popc(bal1 & below);
                    if (slot < kSmemCand) reinterpret_cast<int2*>(ca.u_conf)[slot] = make_int2(__float_as_int(t1), i1);
                    else overflow_put(p.ws + (size_t)img * p.ws_per_image, P, slot, t1, i1);
                }
            }
        }
        if (TRAIN) {
            // Fused step: the candidates of this piece are listed BEFORE its dense pass runs, so that the NMS team (whose
            // own pieces landed first) can rank and decode while the last pieces' dense passes are still writing dL/dy;
            // what needs the complete gradient -- the records, which overwrite its rows -- waits on bar_dense instead.
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar_list)) : "memory");
            dense_pass();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar_dense)) : "memory");
        }
    } else
    // ---------------- A: threshold + stage the survivors' rows ----------------
    for (int base = 0; base < P; base += kThreads * kLoadUnroll) {
        float val[kLoadUnroll];
#pragma unroll
        for (int u = 0; u < kLoadUnroll; ++u) {
            const int i = base + u * kThreads + tid;
            val[u] = 0.f;
            if (i < P) val[u] = head ? __ldg(yimg + box_off(i) + 4) : __ldg(p.conf + (size_t)img * P + i);
        }
#pragma unroll
        for (int u = 0; u < kLoadUnroll; ++u) {
            const int i = base + u * kThreads + tid;
            float conf = 0.f;
            bool pass = false;
            if (i < P) {
                if (!head) {
                    conf = val[u];
                    pass = conf >= p.conf_thre;
                } else if (!(val[u] < p.to_reject)) {  // far below the threshold: sigmoid not needed
                    conf = yh_sigmoid(val[u]);
                    pass = conf >= p.conf_thre;  // models/utils.py:92
                }
            }
            // warp-aggregated slot reservation; the first kSmemCand slots live in shared memory
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&s_count, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (pass) {
                    const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
                    if (slot < kSmemCand) { ca.u_conf[slot] = conf; ca.u_idx[slot] = i; }
                    else overflow_put(p.ws + (size_t)img * p.ws_per_image, P, slot, conf, i);
                    if (head && slot < p.stage_slots) {
                        float* dst = stage + (size_t)slot * p.slot_floats;
                        if (v2) {
                            tx += stage_span(dst, box_off(i), 5 + C);
                        } else {
                            tx += stage_span(dst, box_off(i), 5);
                            tx += stage_span(dst + 8, cls_off(i), C);
                        }
                    }
                }
            }
        }
    }
    // two arrivals per thread: "my candidates are listed" and, carrying the bytes its bulk copies
    // will deliver, "my rows are on their way".  Ranking only needs the list, so it runs while the
    // rows are still in flight.
    if (!TRAIN) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar_list)) : "memory");
    if (!IMG) {
        if (tx) yh_mbar_expect_tx(&bar, tx);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(&bar)) : "memory");
    }
    NT(2);
    yh_mbar_wait(&bar_list, 0);
    if (IMG) {  // (all pieces have landed by now: every group has gone through its own; this makes them visible)
#pragma unroll
        for (int q = 0; q < kGroups; ++q) yh_mbar_wait(&bar_img[q], 0);
    }

    const int K = s_count;
    const bool overflow = K > kSmemCand;
    const bool with_labels = head ? (p.class_aware != 0) : (p.labels != nullptr);
    int2* const pairs = reinterpret_cast<int2*>(ca.u_conf);  // IMG: (u_conf | u_idx) hold 64-bit entries instead

    // ---- fused step: the image's ground-truth records, by the warps that do not resolve the NMS.  Records are dealt
    // by cell (cell % team warps), so the records of one cell stay on one warp, in CSR order: collisions on a
    // predictor accumulate deterministically.  The rows hold the dense pass' values (complete: every thread has
    // passed the list barrier); a cell's first record overwrites without reading them back.
    auto do_records = [&]() {
        constexpr int kRecWarps = kWarps - kTeam / 32;
        const int tw = warp - kTeam / 32;
        const int nrec = o1 - o0;
        if (nrec <= 0) return;
        const int nsm = min(nrec, kRecSmem);
        yh_mbar_wait(&bar_rec, 0);
        const float* win = reinterpret_cast<const float*>(smem_raw);
        float* dimg = p.dy ? p.dy + (long long)img * p.img_floats : nullptr;
        const float my_pw = lane < 5 * A ? g.pw[lane / 5] : 0.f, my_ph = lane < 5 * A ? g.ph[lane / 5] : 0.f;
        const bool track = g.cells <= 64 * kRecWarps;  // cells this warp has updated fit one 64-bit mask
        unsigned long long seen = 0ull;
        for (int base = 0; base < nrec; base += 32) {
            int lc = -1;
            const int r = base + lane;
            if (r < nrec) {
                const int4 h = r < nsm ? s_rec[3 * r] : __ldg(reinterpret_cast<const int4*>(p.gt + o0 + r));
                const bool ok = h.x == img && h.y >= 0 && h.y < g.s_h && h.z >= 0 && h.z < g.s_w;
                lc = ok ? h.y * g.s_w + h.z : -1;
                if (lc >= 0 && lc % kRecWarps != tw) lc = -1;
            }
            unsigned bal = __ballot_sync(0xffffffffu, lc >= 0);
            auto load_record = [&](int rj) -> RecordRegs {
                RecordRegs rr;
                if (rj < nsm) {
                    rr.hd = s_rec[3 * rj];
                    rr.tt = *reinterpret_cast<const float4*>(s_rec + 3 * rj + 1);
                    rr.bb = *reinterpret_cast<const float4*>(s_rec + 3 * rj + 2);
                } else {
                    const int4* rp = reinterpret_cast<const int4*>(p.gt + o0 + rj);
                    rr.hd = __ldg(rp);
                    rr.tt = __ldg(reinterpret_cast<const float4*>(rp + 1));
                    rr.bb = __ldg(reinterpret_cast<const float4*>(rp + 2));
                }
                return rr;
            };
            if (TRAIN == 2 && nrec > kRec4Min) {
                // an image dense with ground truth: four records per warp at a time, eight lanes each (yh_record.cuh)
                while (bal) {
                    const int b = yh_batch4(bal, lc, lane);
                    const bool on = b >= 0;
                    const int lcell_ = __shfl_sync(0xffffffffu, lc, on ? b : 0);  // (every lane takes part)
                    const int lcell = on ? lcell_ : 0;
                    const int rj = base + (on ? b : 0);
                    RecordRegs rr;
                    rr.hd = make_int4(0, 0, 0, 0);
                    rr.tt = rr.bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (on) rr = load_record(rj);
                    const unsigned long long bit = 1ull << ((lcell / kRecWarps) & 63);
                    const bool again = !track || (seen & bit) != 0ull;
                    const unsigned lo = __reduce_or_sync(0xffffffffu, on ? (unsigned)bit : 0u);
                    const unsigned hi = __reduce_or_sync(0xffffffffu, on ? (unsigned)(bit >> 32) : 0u);
                    seen |= ((unsigned long long)hi << 32) | lo;
                    if (dimg) process_records4<1>(p, 2, A, C, on, rr, o0 + rj, win + fsh + lcell * cf, dimg + lcell * cf, again, kn, lane, sums);
                    else process_records4<0>(p, 2, A, C, on, rr, o0 + rj, win + fsh + lcell * cf, nullptr, false, kn, lane, sums);
                }
                continue;
            }
            while (bal) {
                const int b = __ffs(bal) - 1;
                bal &= bal - 1u;
                const int lcell = __shfl_sync(0xffffffffu, lc, b);
                const int rj = base + b;
                const RecordRegs rr = load_record(rj);
                const unsigned long long bit = 1ull << ((lcell / kRecWarps) & 63);
                const bool again = !track || (seen & bit) != 0ull;
                seen |= bit;
                if (dimg) process_record<1>(p, 2, A, C, rr, o0 + rj, win + fsh + lcell * cf, dimg + lcell * cf, again, nullptr, nullptr,
                                            kn, lane, my_pw, my_ph, sums);
                else process_record<0>(p, 2, A, C, rr, o0 + rj, win + fsh + lcell * cf, nullptr, false, nullptr, nullptr,
                                       kn, lane, my_pw, my_ph, sums);
            }
        }
    };
    // ---- fused step: the CTA's six loss sums -> the 64-bit fixed-point accumulators (exact, order-independent;
    // yh_train.cu has the full story); the finalize kernel behind this one turns the totals into terms and loss
    auto publish_sums = [&]() {
        const float s_no = yh_warp_sum(sums.no);
        const float s_xy = yh_warp_sum(sums.xy), s_wh = yh_warp_sum(sums.wh), s_conf = yh_warp_sum(sums.conf);
        const float s_nr = yh_warp_sum(sums.nr), s_cls = yh_warp_sum(sums.cls);
        if (lane == 0) {
            float* r = red + warp * 6;
            r[0] = s_xy; r[1] = s_wh; r[2] = s_conf; r[3] = s_no; r[4] = s_nr; r[5] = s_cls;
        }
        __syncthreads();
        if (tid == kThreads - 1) {
            // (an overlapped call shares nothing with the calls in front of it but must not complete before them:
            //  the wait at the end of the kernel covers that; the accumulators are this call's own)
            constexpr double kFix = 4294967296.0;  // 2^32
            unsigned flags = 0u;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                float a = 0.f;
                for (int w = 0; w < kWarps; ++w) a += red[w * 6 + q];
                if (a >= 0.f && a < 1073741824.f) atomicAdd(p.acc + q, (unsigned long long)((double)a * kFix + 0.5));
                else flags |= 1u << q;  // NaN / inf / out of range: the term becomes NaN, like the reference's
            }
            if (flags) atomicOr(p.acc + 6, (unsigned long long)flags);
        }
    };

    // Whole-image mode with at most one tile of candidates, all in shared memory: the lean path.  The
    // kernel's tail is bound by instruction issue (two CTAs per SM, and the next kernel of the stream
    // streaming next to them), so every phase is laid out for few warp instructions:
    //   B  two lanes per candidate: each counts half of the candidates that beat it (16-byte loads of
    //      the unsorted lists), then decodes one axis of the box (x: tx, tw; y: ty, th);
    //   D  the i < j pairs are enumerated densely over all threads (the triangle folded into a
    //      K/2 x (K-1) rectangle), set bits go to the mask with atomicOr (rare), then the fixed-point
    //      resolution of the greedy order by one warp, as in the general path -- while the other warps pick
    //      label and score of every candidate (four lanes per box, without the divisions that cannot matter);
    //   E  four lanes per kept box copy the record out.
    // (fused step: run by the first kTeam threads of the CTA only -- `team_tag` carries the team size, the phases then
    //  meet on a named barrier instead of the CTA barrier -- while the other warps process the image's records)
    auto rest_img = [&](auto lab_tag, auto team_tag) {
        constexpr bool LAB = decltype(lab_tag)::value;
        constexpr int kThreads = decltype(team_tag)::value, kWarps = kThreads / 32;
        auto team_sync = [&]() {
            if (kThreads == NTH) __syncthreads();
            else asm volatile("bar.sync 6, %0;" ::"n"(kThreads) : "memory");
        };
        const float* win = reinterpret_cast<const float*>(smem_raw);
        Cand ca = carve(smem_raw + p.stage_bytes, kSmemCand);
        const bool use_lab = LAB && p.class_aware != 0;
        auto box_row = [&](int k, int ik) -> const float* { return win + (fsh + box_off(ik)); };
        auto cls_row = [&](int ranked) -> const float* { return win + (fsh + cls_off(ca.s_idx[ranked])); };

        // ---------------- B: rank + decode ----------------
        for (int kb = 0; kb < K; kb += kThreads / 2) {
            const int k = kb + (tid >> 1), ax = tid & 1;  // candidate, axis
            if (kb + ((tid & ~31) >> 1) < K) {  // (whole warps)
                const bool on = k < K;
                const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(ca.u_conf);
                const unsigned long long kk = on ? keys[k] : 0ull;
                const float ck = __int_as_float((int)(kk >> 32));
                const int ik = (int)~(unsigned)kk;
                int rank = 0;
                const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(ca.u_conf);
                const int K2 = K >> 1;
                for (int q = ax; q < K2; q += 2) {  // (keys are unique: the predictor index is part of them)
                    const ulonglong2 kj = k2[q];
                    rank += kj.x > kk ? 1 : 0;
                    rank += kj.y > kk ? 1 : 0;
                }
                if (ax == 0 && (K & 1)) rank += keys[K - 1] > kk ? 1 : 0;
                rank += __shfl_xor_sync(0xffffffffu, rank, 1);
                // this lane's axis of the box: lo/hi corner coordinate (same roundings as yh_decode_box)
                float lo = 0.f, hi = 0.f;
                if (on) {
                    const float* row = box_row(k, ik);
                    const float tc = row[ax], ts = row[2 + ax];
                    const float sc = yh_sigmoid(tc);
                    const float sa = v2 ? expf(ts) : yh_sigmoid(ts);
                    const int cell = ik / A, a = ik - cell * A;
                    const int cy = (int)__umulhi((unsigned)cell, p.magic_sw), cx = cell - cy * g.s_w;  // cell / s_w
                    const float bsz = __fmul_rn(ax ? g.ph[a] : g.pw[a], sa);
                    const float bc = __fadd_rn(sc, (float)(ax ? cy : cx));
                    const float hs = __fmul_rn(bsz, 0.5f);
                    const float gsz = ax ? g.gh : g.gw;
                    lo = __fmul_rn(__fsub_rn(bc, hs), gsz);
                    hi = __fmul_rn(__fadd_rn(bc, hs), gsz);
                }
                const float olo = __shfl_xor_sync(0xffffffffu, lo, 1), ohi = __shfl_xor_sync(0xffffffffu, hi, 1);
                if (on && ax == 0) {
                    ca.s_box[rank] = make_float4(lo, olo, hi, ohi);
                    ca.s_area[rank] = __fmul_rn(__fsub_rn(hi, lo), __fsub_rn(ohi, olo));
                    ca.s_idx[rank] = ik;
                    ca.s_conf[rank] = ck;
                }
            }
        }
        team_sync();
        NT(4);

        constexpr int kPick = 4;  // lanes per class pick
        const int sub = tid & (kPick - 1);
        // label and score of ranked candidate / kept box: fast pick, the full one where it does not apply
        // (`with_spec` is uniform over the warp: both functions shuffle across all 32 lanes)
        auto pick = [&](const float* cl, float conf, bool act, bool with_spec, float* spec_out, int* lab, float* sc) {
            bool ok = false;
            if (C <= 4 * kPickRegs && !with_spec) quad_class_pick_fast(cl, C, conf, sub, act, lab, sc, &ok);
            if (!__all_sync(0xffffffffu, ok)) group_class_pick<kPick>(cl, C, conf, sub, act, spec_out, lab, sc);
        };
        // Label and score of EVERY ranked candidate, by the warps w0 .. w0 + nw - 1 (eight candidates per warp and
        // trip).  Class-aware suppression needs the labels before the pair tests; otherwise the pick runs on
        // warps 1.. while warp 0 resolves the greedy order -- a latency-bound chain of ballots during which the
        // other warps would idle -- so that the emit phase only copies.  (~10 % of the candidates are not kept
        // and get picked for nothing: cheaper than a serial pick phase at the end of the kernel's critical path.)
        const bool with_spec = p.out_cls_spec != nullptr;
        const bool want_ls = p.out_label != nullptr || p.out_score != nullptr;
        float* s_score = ca.u_conf;  // (the sort keys there are dead once phase B is through)
        auto pick_all = [&](int w0, int nw) {
            for (int k0 = 8 * (warp - w0); k0 < K; k0 += 8 * nw) {  // (uniform over the warp)
                const int k = k0 + (lane >> 2);
                const bool act = k < K;
                int lab;
                float sc;
                pick(act ? cls_row(k) : nullptr, act ? ca.s_conf[k] : 0.f, act, false, nullptr, &lab, &sc);
                if (act && sub == 0) { ca.s_lab[k] = lab; s_score[k] = sc; }
            }
        };
        const bool picked = !with_spec && (use_lab || want_ls);  // emit finds label and score in shared memory
        if (use_lab) {
            pick_all(0, kWarps);
            team_sync();
        }

        // ---------------- D: greedy suppression (one tile) ----------------
        const float thr = p.iou_thre;
        const bool thr_pos = thr > 0.f;
        const int tn = K;
        const int W = (tn + 31) >> 5;
        {
            const int Ke = K + (K & 1), cols = Ke - 1, total = (Ke >> 1) * cols;
            // pair / cols through the float reciprocal: (pair + 0.5) / cols is at least 0.5 / cols >= 2^-9 away from
            // an integer and the quotient is below 2^7, so the few ulp of the approximate division cannot move
            // its integer part (pair < 2^15: exact in float)
            const float inv_cols = __fdividef(1.0f, (float)cols);
            for (int pr = tid; pr < total; pr += kThreads) {
                const int r = (int)(((float)pr + 0.5f) * inv_cols);
                const int c = pr - r * cols;
                const int i = c >= r ? r : Ke - 1 - r;
                const int j = c >= r ? c + 1 : Ke - 1 - c;
                if (j < K) {
                    bool bit = suppresses_dense(ca.s_box[i], ca.s_area[i], ca.s_box[j], ca.s_area[j], thr, thr_pos);
                    if (use_lab) bit = bit && ca.s_lab[i] == ca.s_lab[j];
                    if (bit) atomicOr(&mask[j * kTileWords + (i >> 5)], 1u << (i & 31));
                }
            }
        }
        team_sync();
        NT(7);
        if (warp != 0 && picked && !use_lab) pick_all(1, kWarps - 1);
        if (warp == 0) {
            unsigned alive[kTileWords], dead0[kTileWords];
#pragma unroll
            for (int w = 0; w < kTileWords; ++w) {
                dead0[w] = w < W ? 0u : 0xffffffffu;
                if (w == W - 1 && (tn & 31)) dead0[w] |= ~0u << (tn & 31);  // bits past the tile end
                alive[w] = ~dead0[w];
            }
            if (W <= 2) {
                // common case (<= 64 candidates): the lane's two columns live in registers
                const unsigned c00 = lane < tn ? mask[lane * kTileWords] : 0u;
                const unsigned c10 = 32 + lane < tn ? mask[(32 + lane) * kTileWords] : 0u;
                const unsigned c11 = 32 + lane < tn ? mask[(32 + lane) * kTileWords + 1] : 0u;
                for (int sweep = 0; sweep <= tn; ++sweep) {
                    const unsigned n0 = __ballot_sync(0xffffffffu, (c00 & alive[0]) == 0u) & ~dead0[0];
                    const unsigned n1 = __ballot_sync(0xffffffffu, ((c10 & n0) | (c11 & alive[1])) == 0u) & ~dead0[1];
                    const bool same = n0 == alive[0] && n1 == alive[1];
                    alive[0] = n0;
                    alive[1] = n1;
                    if (same) break;
                }
            } else {
                for (int sweep = 0; sweep <= tn; ++sweep) {
                    bool changed = false;
#pragma unroll
                    for (int m = 0; m < kTileWords; ++m) {
                        if (m < W) {  // (warp-uniform)
                            const int j = 32 * m + lane;
                            bool a = false;
                            if (j < tn) {
                                unsigned hit = 0u;
#pragma unroll
                                for (int w = 0; w < kTileWords; ++w)
                                    if (w <= m) hit |= mask[j * kTileWords + w] & alive[w];
                                a = hit == 0u;
                            }
                            const unsigned nw = __ballot_sync(0xffffffffu, a) & ~dead0[m];
                            changed = changed || nw != alive[m];
                            alive[m] = nw;  // (later words of this sweep already see it)
                        }
                    }
                    if (!changed) break;
                }
            }
            int kept_n = 0;
#pragma unroll
            for (int m = 0; m < kTileWords; ++m) {
                if (m < W) {
                    if ((alive[m] >> lane) & 1u) ca.keep[kept_n + __popc(alive[m] & ((1u << lane) - 1u))] = 32 * m + lane;
                    kept_n += __popc(alive[m]);
                }
            }
            if (lane == 0) s_kept = kept_n;
        }
        team_sync();
        NT(8);

        NT(12);
        // ---------------- E: emit ----------------
        const int kept = s_kept;
        if (tid == 0) p.keep_cnt[img] = kept;
        const int nout = min(kept, p.max_out);
        const bool want_cls = p.out_cls_spec || p.out_label || p.out_score;
        if (picked || !want_cls) {
            // nothing left to compute: one thread per kept box copies its record out
            for (int t = tid; t < nout; t += kThreads) {
                const int i = ca.keep[t];
                const size_t o = (size_t)img * p.max_out + t;
                p.keep_idx[o] = ca.s_idx[i];
                if (p.out_conf) p.out_conf[o] = ca.s_conf[i];
                if (p.out_bbox) p.out_bbox[o] = ca.s_box[i];
                if (picked) {
                    if (p.out_label) p.out_label[o] = ca.s_lab[i];
                    if (p.out_score) p.out_score[o] = s_score[i];
                }
            }
            NT(13);
            return;
        }
        for (int t0 = 0; t0 < nout; t0 += kThreads / kPick) {
            if (t0 + (32 / kPick) * warp >= nout) break;
            const int t = t0 + tid / kPick;
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
            if (want_cls) {  // (cls_spec rows requested: the full pick, four lanes per kept box)
                int lab;
                float sc;
                pick(act ? cls_row(i) : nullptr, conf, act, with_spec,
                     (act && with_spec) ? p.out_cls_spec + o * C : nullptr, &lab, &sc);
                if (act && sub == 2) {
                    if (p.out_label) p.out_label[o] = lab;
                    if (p.out_score) p.out_score[o] = sc;
                }
            }
        }
        NT(13);
    };

    if (IMG && !overflow) {
        // the common case first, and out of the way of everything below
        // (logit, predictor) -> sort key: confidence bits in the high word (positive floats order like
        // integers), complement of the predictor index in the low word (ties: lower index first)
        if (TRAIN) {
            if (warp < kTeam / 32) {
                for (int k = tid; k < K; k += kTeam) {
                    const int2 e = pairs[k];
                    pairs[k] = make_int2(~e.y, __float_as_int(yh_sigmoid(__int_as_float(e.x))));
                }
                asm volatile("bar.sync 6, %0;" ::"n"(kTeam) : "memory");
                if (!with_labels) rest_img(FastTag<false>{}, IntTag<kTeam>{});
                else rest_img(FastTag<true>{}, IntTag<kTeam>{});
            } else {
                yh_mbar_wait(&bar_dense, 0);  // the whole image's dense dL/dy is written: the records go on top of it
                NTR(9);
                do_records();
                NTR(10);
            }
            publish_sums();
            NT(11);
        } else {
            for (int k = tid; k < K; k += kThreads) {
                const int2 e = pairs[k];
                pairs[k] = make_int2(~e.y, __float_as_int(yh_sigmoid(__int_as_float(e.x))));
            }
            __syncthreads();
            if (!with_labels) rest_img(FastTag<false>{}, IntTag<NTH>{});
            else rest_img(FastTag<true>{}, IntTag<NTH>{});
        }
        if (p.late_wait && img == 0 && tid == 0) yh_grid_dependency_wait();  // (see the end of the kernel)
        return;
    }
    if (TRAIN) yh_mbar_wait(&bar_dense, 0);  // (an image with more candidates than shared memory holds: the whole CTA goes on together)
    if (TRAIN && warp >= kTeam / 32) do_records();  // (... records first)

    Cand cw = ca;
    if (overflow) {  // continue in the workspace arrays
        cw = carve_ws();
        if (IMG) {
            // phase A left (logit, predictor) pairs: confidence = sigmoid(logit), into the separate arrays
            for (int k = tid; k < K; k += kThreads) {
                float t;
                int i;
                if (k < kSmemCand) { t = __int_as_float(pairs[k].x); i = pairs[k].y; }
                else { t = cw.u_conf[k]; i = cw.u_idx[k]; }
                cw.u_conf[k] = yh_sigmoid(t);
                cw.u_idx[k] = i;
            }
        } else {
            for (int k = tid; k < kSmemCand; k += kThreads) { cw.u_conf[k] = ca.u_conf[k]; cw.u_idx[k] = ca.u_idx[k]; }
        }
        __syncthreads();
    }

    // Everything after the candidate list exists twice: the common case -- all candidates and all their
    // rows in shared memory -- is compiled with pointers the compiler can PROVE to be shared (32-bit
    // addresses, LDS/STS); the general case (workspace arrays, rows beyond the staged slots) goes
    // through generic pointers.  The kernel is bound by instruction issue, and generic accesses with
    // their 64-bit address arithmetic were a quarter of it.
    auto rest = [&](auto fast_tag, auto lab_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
        constexpr bool LAB = decltype(lab_tag)::value;  // labels take part in the suppression test
        Cand ca = carve(smem_raw + p.stage_bytes, kSmemCand);
        if (!FAST && overflow) ca = cw;
        auto box_ptr = [&](int slot, int idx) -> const float* {
            const int f = box_off(idx);
            if (IMG) return reinterpret_cast<const float*>(smem_raw) + (fsh + f);
            if (FAST || slot < p.stage_slots) return reinterpret_cast<const float*>(smem_raw) + slot * p.slot_floats + ((fsh + f) & 3);
            return yimg + f;
        };
        auto cls_ptr = [&](int slot, int idx) -> const float* {
            if (IMG) return reinterpret_cast<const float*>(smem_raw) + (fsh + cls_off(idx));
            if (FAST || slot < p.stage_slots) {
                const float* st = reinterpret_cast<const float*>(smem_raw) + slot * p.slot_floats;
                if (v2) return st + ((fsh + box_off(idx)) & 3) + 5;
                return st + 8 + ((fsh + cls_off(idx)) & 3);
            }
            return yimg + cls_off(idx);
        };

        // ---------------- B: rank + decode, eight lanes per candidate ----------------
        // A candidate's rank = how many candidates beat it: the eight lanes split the comparisons and
        // fold with three shuffles.  Then (rows landed) four of them activate one box logit each and the
        // group's first lane decodes the box into the ranked slot.
        const int sub8 = tid & 7;
        bool rows_in = false;
        // Images that overflowed the shared-memory lists keep their candidates in the workspace -- but the 12 KB of the
        // shared lists are idle then, and they hold a 64-bit sort key (order-preserving image of the confidence |
        // complemented predictor index, as in the lean path) for up to 1536 candidates: the K^2 comparisons of the
        // ranking read those with 16-byte shared loads instead of two global loads each (K = 420: 22 -> 3 us).
        constexpr int kKeyMax = kSmemCand * (16 + 4 * 8) / 8;  // cand_bytes(kSmemCand) / 8
        unsigned long long* const skeys = reinterpret_cast<unsigned long long*>(smem_raw + p.stage_bytes);
        const bool keyed = !FAST && overflow && K <= kKeyMax;
        auto sort_key = [](float c, int idx) -> unsigned long long {
            const unsigned u = __float_as_uint(c + 0.0f);  // (-0 -> +0: equal confidences must give equal high words)
            const unsigned o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            return ((unsigned long long)o << 32) | (unsigned)~idx;
        };
        if (keyed) {
            for (int k = tid; k < K; k += kThreads) skeys[k] = sort_key(ca.u_conf[k], ca.u_idx[k]);
            __syncthreads();
        }
        for (int k0 = 0; k0 < K; k0 += kThreads / 8) {
            if (k0 + 4 * warp >= K) break;  // no candidate left for this warp: it stays out of the issue slots
            const int k = k0 + (tid >> 3);
            const bool on = k < K;
            const float ck = on ? ca.u_conf[k] : 0.f;
            const int ik = on ? ca.u_idx[k] : 0;
            int rank = 0;
            if (on && keyed) {
                const unsigned long long kk = skeys[k];
                const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(skeys);
                const int K2 = K >> 1;
                for (int q = sub8; q < K2; q += 8) {  // (keys are unique: the predictor index is part of them)
                    const ulonglong2 kj = k2[q];
                    rank += kj.x > kk ? 1 : 0;
                    rank += kj.y > kk ? 1 : 0;
                }
                if (sub8 == 0 && (K & 1)) rank += skeys[K - 1] > kk ? 1 : 0;
            } else if (on) {
                for (int j = sub8; j < K; j += 8) {
                    const float cj = ca.u_conf[j];
                    rank += (cj > ck || (cj == ck && ca.u_idx[j] < ik)) ? 1 : 0;
                }
            }
            rank += __shfl_xor_sync(0xffffffffu, rank, 1);
            rank += __shfl_xor_sync(0xffffffffu, rank, 2);
            rank += __shfl_xor_sync(0xffffffffu, rank, 4);
            if (!rows_in) {
                NT(3);
                if (!IMG) yh_mbar_wait(&bar, 0);  // the staged rows
                rows_in = true;
            }
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            int lab = 0;
            if (head) {
                float act = 0.f;
                if (on && sub8 < 4) {
                    const float tq = box_ptr(k, ik)[sub8];
                    act = (sub8 < 2 || !v2) ? yh_sigmoid(tq) : expf(tq);
                }
                const int l0 = lane & ~7;
                const float sx = __shfl_sync(0xffffffffu, act, l0), sy = __shfl_sync(0xffffffffu, act, l0 + 1);
                const float wa = __shfl_sync(0xffffffffu, act, l0 + 2), ha = __shfl_sync(0xffffffffu, act, l0 + 3);
                if (on && sub8 == 0) {
                    const int cell = ik / A, a = ik - cell * A;
                    const int cy = (int)__umulhi((unsigned)cell, p.magic_sw), cx = cell - cy * g.s_w;  // cell / s_w
                    const YhBox b = yh_decode_box(sx, sy, wa, ha, g.pw[a], g.ph[a], cx, cy, g.gw, g.gh);
                    bx = make_float4(b.x1, b.y1, b.x2, b.y2);
                }
            } else if (on && sub8 == 0) {
                bx = __ldg(p.bbox + (size_t)img * P + ik);
                if (p.labels) lab = __ldg(p.labels + (size_t)img * P + ik);
            }
            if (on && sub8 == 0) {
                ca.s_box[rank] = bx;
                ca.s_area[rank] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
                ca.s_slot[rank] = k;
                ca.s_idx[rank] = ik;
                ca.s_conf[rank] = ck;
                if (!head && p.labels) ca.s_lab[rank] = lab;
            }
        }
        __syncthreads();

        NT(4);
        const bool use_lab = LAB && (head ? (p.class_aware != 0) : (p.labels != nullptr));
        constexpr int kPick = 4;  // lanes per class pick (as in the whole-image path: same sums, same bits)
        const int sub = tid & (kPick - 1);
        if (use_lab && head) {  // label of every candidate (argmax of cls_spec)
            for (int k0 = 0; k0 < K; k0 += kThreads / kPick) {
                if (k0 + (32 / kPick) * warp >= K) break;
                const int k = k0 + tid / kPick;
                const bool act = k < K;
                int lab;
                float sc;
                group_class_pick<kPick>(act ? cls_ptr(ca.s_slot[k], ca.s_idx[k]) : nullptr, C, act ? ca.s_conf[k] : 0.f, sub,
                                        act, nullptr, &lab, &sc);
                if (act && sub == 0) ca.s_lab[k] = lab;
            }
            __syncthreads();
        }

        // ---------------- D: greedy suppression, tile by tile ----------------
        const float thr = p.iou_thre;
        const bool thr_pos = thr > 0.f;
        for (int base = 0; base < K; base += kTile) {
            const int tn = min(kTile, K - base);
            const int W = (tn + 31) >> 5;
            // (1) against the boxes kept in earlier tiles: one warp per candidate, the lanes split the kept boxes
            //     (a serial walk per thread was the longest phase of an image with several tiles)
            if (tid < kTileWords) rem0[tid] = 0u;
            // the tile's ranked boxes, areas and labels -> the idle shared lists (read K/2 times each below)
            float4* const t_box = reinterpret_cast<float4*>(smem_raw + p.stage_bytes);
            float* const t_area = reinterpret_cast<float*>(t_box + kTile);
            int32_t* const t_lab = reinterpret_cast<int32_t*>(t_area + kTile);
            const bool tiled = !FAST && overflow;  // (the shared lists are idle: everything lives in the workspace)
            if (tiled) {
                __syncthreads();  // (the sort keys / the previous tile's copy are no longer read)
                for (int j = tid; j < tn; j += kThreads) {
                    t_box[j] = ca.s_box[base + j];
                    t_area[j] = ca.s_area[base + j];
                    if (use_lab) t_lab[j] = ca.s_lab[base + j];
                }
            }
            __syncthreads();
            if (base > 0) {
                const int kept = s_kept;
                for (int jp = warp; jp < tn; jp += kWarps) {
                    const float4 bj = tiled ? t_box[jp] : ca.s_box[base + jp];
                    const float aj = tiled ? t_area[jp] : ca.s_area[base + jp];
                    const int lj = use_lab ? (tiled ? t_lab[jp] : ca.s_lab[base + jp]) : 0;
                    bool hit = false;
                    for (int q0 = 0; q0 < kept && !hit; q0 += 32) {
                        const int q = q0 + lane;
                        bool bit = false;
                        if (q < kept) {
                            const int i = ca.keep[q];
                            bit = suppresses_dense(ca.s_box[i], ca.s_area[i], bj, aj, thr, thr_pos) && (!use_lab || ca.s_lab[i] == lj);
                        }
                        hit = __any_sync(0xffffffffu, bit);
                    }
                    if (hit && lane == 0) atomicOr(&rem0[jp >> 5], 1u << (jp & 31));
                }
            }
            __syncthreads();
            // (2) intra-tile mask, by column: word (j, w) = candidates i in [32w, 32w+32), ranked before j,
            //     that suppress j if kept; each warp takes columns j = warp, warp + kWarps, ... and walks
            //     the words up to j's own (two at a time: the tests are independent)
            if (base == 0) NT(5);
            for (int j = warp; j < tn; j += kWarps) {
                const float4 bj = tiled ? t_box[j] : ca.s_box[base + j];
                const float aj = tiled ? t_area[j] : ca.s_area[base + j];
                const int lj = use_lab ? (tiled ? t_lab[j] : ca.s_lab[base + j]) : 0;
    #pragma unroll 2
                for (int w = 0; w <= (j >> 5); ++w) {
                    const int i = w * 32 + lane;
                    bool bit = false;
                    if (i < j) {
                        bit = tiled ? suppresses_dense(t_box[i], t_area[i], bj, aj, thr, thr_pos)
                                    : suppresses_dense(ca.s_box[base + i], ca.s_area[base + i], bj, aj, thr, thr_pos);
                        if (use_lab) bit = bit && lj == (tiled ? t_lab[i] : ca.s_lab[base + i]);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, bit);
                    if (lane == 0) mask[j * kTileWords + w] = m;
                }
            }
            __syncthreads();
            if (base == 0) NT(7);
            // (3) one warp resolves the greedy order by fixed-point iteration instead of a serial walk:
            //     alive[j] = !dead0[j] && no alive i < j suppresses j.  Candidate j's value is final once
            //     all i < j are final, so after t sweeps the first t candidates are right; in practice the
            //     suppression chains are 2-3 deep and the sweep converges in as many steps (<= tn + 1).
            //     The same warp then appends the survivors, in rank order, to the keep list.
            if (warp == 0) {
                unsigned alive[kTileWords], dead0[kTileWords];
    #pragma unroll
                for (int w = 0; w < kTileWords; ++w) {
                    dead0[w] = w < W ? rem0[w] : 0xffffffffu;
                    if (w == W - 1 && (tn & 31)) dead0[w] |= ~0u << (tn & 31);  // bits past the tile end
                    alive[w] = ~dead0[w];
                }
                if (W <= 2) {
                    // common case (<= 64 candidates): the lane's two columns live in registers
                    const unsigned c00 = lane < tn ? mask[lane * kTileWords] : 0u;
                    const unsigned c10 = 32 + lane < tn ? mask[(32 + lane) * kTileWords] : 0u;
                    const unsigned c11 = 32 + lane < tn ? mask[(32 + lane) * kTileWords + 1] : 0u;
                    for (int sweep = 0; sweep <= tn; ++sweep) {
                        const unsigned n0 = __ballot_sync(0xffffffffu, (c00 & alive[0]) == 0u) & ~dead0[0];
                        const unsigned n1 = __ballot_sync(0xffffffffu, ((c10 & n0) | (c11 & alive[1])) == 0u) & ~dead0[1];
                        const bool same = n0 == alive[0] && n1 == alive[1];
                        alive[0] = n0;
                        alive[1] = n1;
                        if (same) break;
                    }
                } else {
                    for (int sweep = 0; sweep <= tn; ++sweep) {
                        bool changed = false;
    #pragma unroll
                        for (int m = 0; m < kTileWords; ++m) {
                            if (m < W) {  // (warp-uniform)
                                const int j = 32 * m + lane;
                                bool a = false;
                                if (j < tn) {
                                    unsigned hit = 0u;
    #pragma unroll
                                    for (int w = 0; w < kTileWords; ++w)
                                        if (w <= m) hit |= mask[j * kTileWords + w] & alive[w];
                                    a = hit == 0u;
                                }
                                const unsigned nw = __ballot_sync(0xffffffffu, a) & ~dead0[m];
                                changed = changed || nw != alive[m];
                                alive[m] = nw;  // (later words of this sweep already see it)
                            }
                        }
                        if (!changed) break;
                    }
                }
                int kept_n = s_kept;
    #pragma unroll
                for (int m = 0; m < kTileWords; ++m) {
                    if (m < W) {
                        if ((alive[m] >> lane) & 1u) ca.keep[kept_n + __popc(alive[m] & ((1u << lane) - 1u))] = base + 32 * m + lane;
                        kept_n += __popc(alive[m]);
                    }
                }
                if (lane == 0) s_kept = kept_n;
            }
            __syncthreads();
            if (base == 0) NT(8);
        }

        NT(12);
        // ---------------- E: emit ----------------
        const int kept = s_kept;
        if (tid == 0) p.keep_cnt[img] = kept;
        const int nout = min(kept, p.max_out);
        const bool want_cls = head && (p.out_cls_spec || p.out_label || p.out_score);
        for (int t0 = 0; t0 < nout; t0 += kThreads / kPick) {
            if (t0 + (32 / kPick) * warp >= nout) break;
            const int t = t0 + tid / kPick;
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
                group_class_pick<kPick>(act ? cls_ptr(ca.s_slot[i], idx) : nullptr, C, conf, sub, act,
                                        (act && p.out_cls_spec) ? p.out_cls_spec + o * C : nullptr, &lab, &sc);
                if (act && sub == 2) {
                    if (p.out_label) p.out_label[o] = lab;
                    if (p.out_score) p.out_score[o] = sc;
                }
            }
        }
        NT(13);
    };

    const bool fast = !IMG && !overflow && K <= p.stage_slots;  // (IMG: only images that overflowed get here)
    if (fast && !with_labels) rest(FastTag<true>{}, FastTag<false>{});
    else if (fast) rest(FastTag<true>{}, FastTag<true>{});
    else rest(FastTag<false>{}, FastTag<true>{});
    if (TRAIN) publish_sums();
    // (overlapped call) everything above ran next to the tails of the kernels in front of it on the stream; this
    // grid must not complete before they have, or work launched after it could overtake them.  ONE thread of the
    // grid waits: the grid is not complete until it exits, while every other CTA leaves and frees its SM for the
    // kernels behind (all CTAs waiting here held their shared memory idle until the slowest predecessor was done)
    if (p.late_wait && img == 0 && tid == 0) yh_grid_dependency_wait();
}

template <int TV, int TA, int TC, int MODE, int NTH, int TRAIN = 0>
int launch_variant(const NmsParams& p, size_t smem, void* stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > 32 * 1024 && smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_nms_kernel<TV, TA, TC, MODE, NTH, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(nms)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    return yh_check_cuda(yh_launch_pdl(yh_nms_kernel<TV, TA, TC, MODE, NTH, TRAIN>, dim3((unsigned)p.n),
                                       dim3(NTH), smem, (cudaStream_t)stream, p),
                         "yh_nms launch");
}

// whole-image staging applies: the image's aligned window (<= 3 floats of shift in front, padded to 16 bytes) fits in
// shared memory (next to a second CTA's up to ~96 KB); needs bulk copies (aligned y) and C >= 3 (then the floats past
// the last whole 16 bytes of the tensor, which arrive by plain loads, are never an objectness logit)
bool img_mode_applies(const NmsParams& p) {
    const int win_bytes = ((p.img_floats + 3 + 3) & ~3) * 4;
    return p.src == SRC_HEAD && p.use_tma && p.c >= 3 && win_bytes <= kImgBytesMax;
}

int launch(NmsParams& p, void* ws, size_t ws_bytes, void* stream, bool train = false) {
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
    const int win_bytes = ((p.img_floats + 3 + 3) & ~3) * 4;
    const bool img_mode = img_mode_applies(p);
    p.stage_bytes = img_mode ? win_bytes : p.stage_slots * p.slot_floats * 4;
    const size_t smem = (size_t)p.stage_bytes + cand_bytes(kSmemCand);
    // compile-time geometries for the shapes the reference uses (VOC: YOLOv2 5 anchors x 20 classes,
    // YOLOv1 B=2, C=20); anything else, and decoded-box input, runs the run-time-geometry variant
    constexpr int NTF = kThreadsFull;
    // Images whose stage leaves room for ONE CTA per SM only (19x19x5x25: 180.5 KB) run on 1024 threads instead of
    // 512: twice the warps for the dense pass and, in the fused step, 24 record warps instead of 8 next to the NMS
    // team (BASELINE config 5: fused step 90.8 -> 85.7 us stream-ordered, 72.6 -> 69.0 us chained; post-process
    // 35.9 -> 34.9 us).  Same code, same results: nothing in the kernel depends on the CTA size but its loop strides.
    constexpr int NTB = 1024;
    const bool one_cta = img_mode && 2 * (smem + 12 * 1024) > 227 * 1024;
    if (train) {  // the fused step (the caller checked img_mode_applies and version 2)
        // batches dense with ground truth (BASELINE config 5) run the variant whose record warps take four records at
        // a time (yh_record.cuh: same bits); the common, sparse case keeps the kernel without that code
        const bool dense_gt = (long long)p.m_local > (long long)kRec4Min * p.n;
        if (one_cta) {
            if (p.g.a == 5 && p.c == 20)
                return dense_gt ? launch_variant<2, 5, 20, MODE_IMG, NTB, 2>(p, smem, stream) : launch_variant<2, 5, 20, MODE_IMG, NTB, 1>(p, smem, stream);
            return dense_gt ? launch_variant<2, 0, 0, MODE_IMG, NTB, 2>(p, smem, stream) : launch_variant<2, 0, 0, MODE_IMG, NTB, 1>(p, smem, stream);
        }
        if (p.g.a == 5 && p.c == 20)
            return dense_gt ? launch_variant<2, 5, 20, MODE_IMG, NTF, 2>(p, smem, stream) : launch_variant<2, 5, 20, MODE_IMG, NTF, 1>(p, smem, stream);
        return dense_gt ? launch_variant<2, 0, 0, MODE_IMG, NTF, 2>(p, smem, stream) : launch_variant<2, 0, 0, MODE_IMG, NTF, 1>(p, smem, stream);
    }
    if (img_mode && one_cta) {
        if (p.g.version == 2 && p.g.a == 5 && p.c == 20) return launch_variant<2, 5, 20, MODE_IMG, NTB>(p, smem, stream);
        return launch_variant<0, 0, 0, MODE_IMG, NTB>(p, smem, stream);
    }
    if (img_mode) {
        if (p.g.version == 2 && p.g.a == 5 && p.c == 20) return launch_variant<2, 5, 20, MODE_IMG, NTF>(p, smem, stream);
        if (p.g.version == 1 && p.g.a == 2 && p.c == 20) return launch_variant<1, 2, 20, MODE_IMG, NTF>(p, smem, stream);
        return launch_variant<0, 0, 0, MODE_IMG, NTF>(p, smem, stream);
    }
    if (p.src == SRC_HEAD && p.g.version == 2 && p.g.a == 5 && p.c == 20) return launch_variant<2, 5, 20, MODE_GENERAL, NTF>(p, smem, stream);
    if (p.src == SRC_HEAD && p.g.version == 1 && p.g.a == 2 && p.c == 20) return launch_variant<1, 2, 20, MODE_GENERAL, NTF>(p, smem, stream);
    return launch_variant<0, 0, 0, MODE_GENERAL, NTF>(p, smem, stream);
}

// everything a head-tensor post-process needs but the launch
int fill_head_params(NmsParams& p, int version, const float* y, int n, int s_h, int s_w, int a, int c,
                     const float* anchors_wh_host, float img_h, float img_w, float conf_thre,
                     float iou_thre, int flags, int max_out, int32_t* keep_idx, int32_t* keep_cnt,
                     float* out_bbox, float* out_conf, float* out_cls_spec, int32_t* out_label,
                     float* out_score) {
    memset(&p, 0, sizeof(p));
    int rc = yh_make_geom(&p.g, version, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w);
    if (rc) return rc;
    YH_REQUIRE(y, YH_ERR_INVALID, "y is NULL");
    YH_REQUIRE(((uintptr_t)y & 3) == 0, YH_ERR_INVALID, "y must be 4-byte aligned");
    p.src = SRC_HEAD;
    p.y = y;
    p.n = n; p.p = p.g.preds; p.c = c;
    p.conf_thre = conf_thre; p.iou_thre = iou_thre;
    yh_conf_band(conf_thre, &p.to_reject, &p.to_accept);
    p.class_aware = (flags & YH_POST_CLASS_AWARE) != 0; p.max_out = max_out;
    p.late_wait = (flags & YH_POST_INPUT_READY) != 0;
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt;
    p.out_bbox = reinterpret_cast<float4*>(out_bbox);
    p.out_conf = out_conf; p.out_cls_spec = out_cls_spec; p.out_label = out_label; p.out_score = out_score;

    p.magic_sw = (unsigned)(0xFFFFFFFFull / (unsigned)s_w) + 1u;
    YH_REQUIRE(p.g.cells < 65536, YH_ERR_UNSUPPORTED, "more than 65535 grid cells per image");
    p.img_floats = p.g.cells * p.g.cell_floats;
    p.total_floats = (long long)n * p.img_floats;
    p.use_tma = ((uintptr_t)y & 15) == 0;
    // a staged row: [<= 3 floats of alignment shift | 5 box logits | C class logits], for v1 the box
    // logits (slot floats 0..7) and the cell's class logits (from float 8) are two separate windows
    p.slot_floats = 8 + ((c + 3 + 3) & ~3);
    int slots = kStageBytesMax / (p.slot_floats * 4);
    p.stage_slots = slots < kSmemCand ? slots : kSmemCand;
    return YH_OK;
}

int postprocess_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
                     const float* anchors_wh_host, float img_h, float img_w, float conf_thre,
                     float iou_thre, int class_aware, int max_out, int32_t* keep_idx, int32_t* keep_cnt,
                     float* out_bbox, float* out_conf, float* out_cls_spec, int32_t* out_label,
                     float* out_score, void* ws, size_t ws_bytes, void* stream) {
    NmsParams p;
    int rc = fill_head_params(p, version, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, conf_thre, iou_thre,
                              class_aware, max_out, keep_idx, keep_cnt, out_bbox, out_conf, out_cls_spec, out_label, out_score);
    if (rc) return rc;
    return launch(p, ws, ws_bytes, stream);
}

}  // namespace

void yh_conf_band(float conf_thre, float* to_reject, float* to_accept) {
    // sigmoid(t) >= thr needs t >= logit(thr) up to a few ulp; reject only with a wide margin
    // (the margins are in logit units: 1e-3 there moves the sigmoid by 1e-3 * conf * (1 - conf), >= 1e-5 for
    // thresholds in [0.01, 0.99] -- far beyond the few ulp of expf and the division; outside that range only a
    // wide reject margin is used and every other logit takes the sigmoid)
    *to_accept = INFINITY;
    if (!(conf_thre > 0.f)) *to_reject = -INFINITY;          // everything passes (or thr is NaN)
    else if (conf_thre > 1.f) *to_reject = INFINITY;         // nothing can pass
    else if (conf_thre >= 0.01f && conf_thre <= 0.99f) {
        const double t = log((double)conf_thre / (1.0 - (double)conf_thre));
        *to_reject = (float)(t - 1e-3);
        *to_accept = (float)(t + 1e-3);
    } else {
        const double t = conf_thre < 0.9999999 ? log((double)conf_thre / (1.0 - (double)conf_thre)) : 16.0;
        *to_reject = (float)((t < 16.0 ? t : 16.0) - 0.01);
    }
}

extern "C" {

size_t yh_train_post_workspace_bytes(int n, int s_h, int s_w, int a, int c) {
    if (n <= 0 || s_h <= 0 || s_w <= 0 || a <= 0 || c <= 0) return 0;
    return 256 + yh_postprocess_workspace_bytes(n, s_h * s_w * a);  // [loss sums | post-process workspace]
}

int yh_v2_train_post(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                     float img_h, float img_w, const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                     const float* lambdas_host, float* dy, float* terms, float* loss, int32_t* resp, float* iou_resp,
                     float conf_thre, float iou_thre, int flags, int max_out, int32_t* keep_idx, int32_t* keep_cnt,
                     float* out_bbox, float* out_conf, float* out_cls_spec, int32_t* out_label, float* out_score,
                     const YhExchange* xch_host, void* ws, size_t ws_bytes, void* stream) {
    YH_REQUIRE(n > 0 && s_h > 0 && s_w > 0 && a > 0 && c > 0, YH_ERR_INVALID, "n, grid, anchors and classes must be positive");
    const size_t need = yh_train_post_workspace_bytes(n, s_h, s_w, a, c);
    YH_REQUIRE(ws && ws_bytes >= need, YH_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
    YH_REQUIRE(((uintptr_t)ws & 255) == 0, YH_ERR_INVALID, "workspace must be 256-byte aligned");
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    const int overlapped = (flags & YH_STEP_OVERLAPPED) ? 1 : 0;
    const int post_flags = ((flags & YH_STEP_CLASS_AWARE) ? YH_POST_CLASS_AWARE : 0) | (overlapped ? YH_POST_INPUT_READY : 0);
    NmsParams p;
    int rc = fill_head_params(p, 2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, conf_thre, iou_thre, post_flags, max_out,
                              keep_idx, keep_cnt, out_bbox, out_conf, out_cls_spec, out_label, out_score);
    if (rc) return rc;
    const bool fusable = img_mode_applies(p) && ((uintptr_t)dy & 15) == 0 && 5 * a <= 32 && p.g.preds >= 2;
    if (!fusable) {
        // inputs the fused kernel does not cover (unaligned tensors, images beyond the shared-memory stage): the
        // kernels of the two separate calls, same results
        rc = yh_train_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, gt, gt_off, m_local, m_global,
                           lambdas_host, dy, terms, loss, resp, iou_resp, base, 256, stream, overlapped, xch_host);
        if (rc) return rc;
        return launch(p, base + 256, ws_bytes - 256, stream);
    }
    YH_REQUIRE(gt_off && terms && loss && lambdas_host, YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(m_local >= 0 && (m_local == 0 || gt), YH_ERR_INVALID, "bad ground-truth arguments");
    YH_REQUIRE(m_global > 0, YH_ERR_EMPTY, "no ground-truth boxes in the batch (m_global=%d)", m_global);
    YH_REQUIRE(m_local <= m_global, YH_ERR_INVALID, "m_local > m_global");
    YH_REQUIRE(((uintptr_t)gt & 15) == 0, YH_ERR_INVALID, "gt must be 16-byte aligned");
    YhLossCoef kc;
    YhFinalParams f;
    yh_loss_coefs(lambdas_host, m_global, p.g.preds, &kc, &f);
    f.acc = reinterpret_cast<unsigned long long*>(base);
    f.terms = terms;
    f.loss = loss;
    rc = yh_fill_exchange(&f, xch_host);
    if (rc) return rc;
    p.dy = dy; p.gt = gt; p.gt_off = gt_off; p.m_local = m_local;
    p.acc = f.acc;
    p.cxy = kc.cxy; p.cwh = kc.cwh; p.cconf = kc.cconf; p.cno = kc.cno; p.ccls = kc.ccls;
    p.resp = resp; p.iou_resp = iou_resp;
    rc = launch(p, base + 256, ws_bytes - 256, stream, true);
    if (rc) return rc;
    if (getenv("YH_DEBUG_SYNC")) {  // (debugging aid: attribute a device fault to the fused kernel; never set under capture)
        rc = yh_check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "fused kernel (debug sync)");
        if (rc) return rc;
    }
    return yh_launch_finalize(f, (cudaStream_t)stream);
}

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
