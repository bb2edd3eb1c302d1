// Batched true-/false-positive matching for evaluation.
//
// Replaces the per-detection inner loop of evaluate_model (reference models/utils.py:231-262):
// a detection of class c is a true positive at IoU level L iff some ground-truth box of class c in
// the same image has get_iou(gt, detection, numpy=True) >= L (no one-to-one assignment; a class with
// no ground truth in the image makes the detection a false positive at every level).  The reference
// evaluates that IoU in float64 (numpy arrays built from Python floats), so does this kernel, with one
// rounding per operation.  One thread per detection slot, one launch for the whole batch; the
// precision/recall/AP arithmetic that follows stays where the reference has it (numpy, host).
#include "yh_common.cuh"

namespace {

constexpr int kMaxLevels = 16;

struct MatchParams {
    const float4* det_bbox;   // [N,max_out] xyxy, float32 (the post-process output)
    const int32_t* det_label; // [N,max_out]
    const int32_t* keep_cnt;  // [N]
    int n, max_out;
    const double* gt_boxes;   // [M,4] xyxy, float64 as annotated
    const int32_t* gt_labels; // [M]
    const int32_t* gt_off;    // [N+1]
    int num_levels;
    double levels[kMaxLevels];
    double* best_iou;         // [N,max_out], -1 where no ground truth of the class exists (or slot unused)
    unsigned char* tp;        // [N,max_out,num_levels]
};

__global__ void __launch_bounds__(256) yh_match_kernel(const MatchParams p) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.n * p.max_out) return;
    const int img = slot / p.max_out, t = slot - img * p.max_out;
    double best = -1.0;
    const int cnt = min(p.keep_cnt[img], p.max_out);
    if (t < cnt) {
        const float4 b = p.det_bbox[slot];
        const double px1 = b.x, py1 = b.y, px2 = b.z, py2 = b.w;
        const int lab = p.det_label[slot];
        const double parea = __dmul_rn(__dsub_rn(px2, px1), __dsub_rn(py2, py1));
        for (int j = p.gt_off[img]; j < p.gt_off[img + 1]; ++j) {
            if (p.gt_labels[j] != lab) continue;
            // get_iou(coord_tgt, coord_pred, numpy=True), models/utils.py:30-63
            const double gx1 = p.gt_boxes[4 * j], gy1 = p.gt_boxes[4 * j + 1], gx2 = p.gt_boxes[4 * j + 2], gy2 = p.gt_boxes[4 * j + 3];
            const double iw = fmax(__dsub_rn(fmin(gx2, px2), fmax(gx1, px1)), 0.0);
            const double ih = fmax(__dsub_rn(fmin(gy2, py2), fmax(gy1, py1)), 0.0);
            const double inter = __dmul_rn(iw, ih);
            const double garea = __dmul_rn(__dsub_rn(gx2, gx1), __dsub_rn(gy2, gy1));
            const double uni = __dsub_rn(__dadd_rn(garea, parea), inter);
            const double iou = __ddiv_rn(inter, __dadd_rn(uni, 1e-6));
            best = fmax(best, iou);
        }
    }
    p.best_iou[slot] = best;
    for (int l = 0; l < p.num_levels; ++l) p.tp[(size_t)slot * p.num_levels + l] = (t < cnt && best >= p.levels[l]) ? 1 : 0;
}

}  // namespace

extern "C" {

int yh_match_detections(const float* det_bbox, const int32_t* det_label, const int32_t* keep_cnt, int n, int max_out,
                        const double* gt_boxes_xyxy, const int32_t* gt_labels, const int32_t* gt_off,
                        const double* levels_host, int num_levels, double* best_iou, unsigned char* tp, void* stream) {
    YH_REQUIRE(n > 0 && max_out > 0, YH_ERR_INVALID, "n and max_out must be positive");
    YH_REQUIRE(num_levels > 0 && num_levels <= kMaxLevels, YH_ERR_INVALID, "1..%d IoU levels (got %d)", kMaxLevels, num_levels);
    YH_REQUIRE(det_bbox && det_label && keep_cnt && gt_off && levels_host && best_iou && tp, YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(((uintptr_t)det_bbox & 15) == 0, YH_ERR_INVALID, "det_bbox must be 16-byte aligned");
    MatchParams p;
    p.det_bbox = reinterpret_cast<const float4*>(det_bbox);
    p.det_label = det_label; p.keep_cnt = keep_cnt; p.n = n; p.max_out = max_out;
    p.gt_boxes = gt_boxes_xyxy; p.gt_labels = gt_labels; p.gt_off = gt_off;
    p.num_levels = num_levels;
    for (int l = 0; l < kMaxLevels; ++l) p.levels[l] = l < num_levels ? levels_host[l] : 2.0;
    p.best_iou = best_iou; p.tp = tp;
    const long long slots = (long long)n * max_out;
    YH_REQUIRE(slots < (1ll << 31), YH_ERR_UNSUPPORTED, "too many detection slots");
    yh_match_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_match_detections launch");
}

}  // extern "C"
