/*
 * yolohead.h -- C ABI of libyolohead.so: the YOLOv1/YOLOv2 detection-head hot path
 * (grid decode, IoU target assignment, loss forward+backward, confidence-thresholded greedy
 * NMS) as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary for hcnoh/object-detection-collection-pytorch.  The reference
 * is pure Python/torch (no FFI of its own); each entry point below names the reference
 * interface (file:line under the reference root) whose arithmetic it replaces.  The Python
 * binding that sits on top (ctypes, no torch types cross this boundary) is
 * object-detection-collection-pytorch_b200/_lib.py; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns all memory (inputs, outputs, workspace); the library never allocates,
 *     frees or keeps a pointer after the call returns;
 *   - calls only enqueue work on `stream` (a cudaStream_t passed as void*): no host
 *     synchronisation, no allocation -> CUDA-graph capturable; one device per call;
 *   - return value: 0 on success, a negative YH_ERR_* code otherwise; the message is available
 *     from yh_last_error() (thread-local); nothing throws or exits;
 *   - head tensor layouts (fp32, contiguous):
 *         v2: y[N][S_h][S_w][A][5+C]   channel order (tx,ty,tw,th,to,cls_0..cls_{C-1})
 *                                      -- reference models/yolov2.py:338-362
 *         v1: y[N][S_h][S_w][5*B+C]    B blocks of (tx,ty,tw,th,to) then C class logits
 *                                      -- reference models/yolov1.py:150-163, 250-261
 *     a "predictor" is one (cell, anchor/box) pair; its flat index inside an image is
 *     (cy*S_w + cx)*A + a.
 */
#ifndef YOLOHEAD_H_
#define YOLOHEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define YH_API __attribute__((visibility("default")))
#else
#define YH_API
#endif

#define YH_ABI_VERSION 2
#define YH_MAX_ANCHORS 16
#define YH_MAX_RANKS 16

#define YH_OK 0
#define YH_ERR_INVALID (-1)   /* bad argument (null pointer, non-positive size, ...)        */
#define YH_ERR_EMPTY (-2)     /* no ground-truth boxes: the reference cannot form its means */
#define YH_ERR_WORKSPACE (-3) /* workspace too small                                        */
#define YH_ERR_CUDA (-4)      /* a CUDA runtime call or the launch failed                   */
#define YH_ERR_UNSUPPORTED (-5)

#define YH_POST_CLASS_AWARE 1
#define YH_POST_INPUT_READY 2

/* One ground-truth box.  The 12 scalars `collate_fn` scatters into its dense per-box grids
 * (reference models/yolov2.py:1466-1499, models/yolov1.py:1264-1299).  tw/th hold `bwbh`
 * (box size in grid units) for v2 and `sig_twth` (box size / S) for v1. */
typedef struct YhGt {
    int32_t img, cy, cx, cls;
    float stx, sty, tw, th;
    float x1, y1, x2, y2;
} YhGt;

YH_API int yh_abi_version(void);
YH_API const char* yh_last_error(void);

/* Bytes of scratch the train entry points need.  The buffer must be zero-filled ONCE when it
 * is allocated; a call leaves it zeroed again (it holds the batch's six fixed-point loss sums
 * between the train kernel and its one-warp finalize kernel -- a train call enqueues both).  Do not
 * share one workspace between calls that may run concurrently on different streams. */
YH_API size_t yh_train_workspace_bytes(void);

/* Fused YOLOv2 head for one training step: decode + responsible-anchor assignment + the five
 * loss terms + dL/dy in ONE pass over y.
 * Replaces YOLOv2.get_loss (reference models/yolov2.py:747-1140, which calls predict
 * :433-649 and models/utils.py:5-65 get_iou) and the autograd backward of that graph
 * (loss.backward(), models/yolov2.py:1271).
 *   anchors_wh_host[A][2]  anchor (w,h) in grid units (models/yolov2.py:49-55), HOST memory
 *   gt[m_local]            records sorted by image, gt_off[N+1] CSR offsets into gt
 *   m_global               number of boxes the means are taken over (== m_local on one GPU;
 *                          the all-rank total when the batch is sharded)
 *   lambdas_host[5]        xy, wh, conf, noobj, cls weights (reference config.py:28-32)
 *   dy                     [N,S_h,S_w,A,5+C] gradient of `loss` w.r.t. y, every element
 *                          written; may be NULL (loss only, e.g. validation)
 *   terms[5], loss[1]      the five means (over m_global) and their weighted sum
 *   resp[m_local], iou_resp[m_local]  responsible anchor per record and its IoU; may be NULL
 */
YH_API int yh_v2_train(const float* y, int n, int s_h, int s_w, int a, int c,
                const float* anchors_wh_host, float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream);

/* Same for YOLOv1: b boxes per cell, box size S*sigmoid(t), class softmax and class loss per
 * cell.  Replaces YOLOv1.get_loss (reference models/yolov1.py:556-931) + its backward. */
YH_API int yh_v1_train(const float* y, int n, int s_h, int s_w, int b, int c,
                float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream);

/* The same two calls for callers that can promise more (HostHeadPipeline and bench.py do).
 * THE OVERLAP CONTRACT (also YH_POST_INPUT_READY and YH_STEP_OVERLAPPED below): an overlapped call starts
 * while the kernels in front of it on `stream` are still running -- it is launched as a programmatic
 * dependent and only waits for them right before it completes (the train head: before it publishes its sums),
 * so stream order still holds for everything launched after it.  Those kernels may themselves be overlapped
 * calls, so the set of kernels running concurrently reaches back to the last launch on the stream that was
 * NOT overlapped (a stream-ordered entry point of this library or any foreign kernel: both wait at their start
 * for everything before them).  The promise therefore is: NONE of this call's buffers -- inputs, outputs AND
 * workspace -- is read or written by ANY kernel launched on `stream` since the last non-overlapped launch
 * (reading the same read-only input, e.g. a post-process on the y its train head reads, is fine).  In
 * practice: rotate R complete buffer sets (workspace included) and issue every R-th call stream-ordered.
 * A caller that cannot promise this uses the stream-ordered entry points.  Same results, bit for bit. */
YH_API int yh_v2_train_overlapped(const float* y, int n, int s_h, int s_w, int a, int c,
                const float* anchors_wh_host, float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream);
YH_API int yh_v1_train_overlapped(const float* y, int n, int s_h, int s_w, int b, int c,
                float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, void* ws, size_t ws_bytes, void* stream);

/* The six outputs of predict() (reference models/yolov2.py:433-649, models/yolov1.py:207-437).
 * Any output pointer may be NULL.  Shapes, with P = S_h*S_w*A predictors per image:
 *   sig_txty[N,P,2]  wh_act[N,P,2] (v2: exp(twth), v1: sigmoid(twth))  bbox[N,P,4] pixels xyxy
 *   conf[N,P]  cls_prob (v2: [N,P,C], v1: [N,S_h*S_w,C])  cls_spec[N,P,C] */
YH_API int yh_v2_decode(const float* y, int n, int s_h, int s_w, int a, int c,
                 const float* anchors_wh_host, float img_h, float img_w,
                 float* sig_txty, float* wh_act, float* bbox, float* conf,
                 float* cls_prob, float* cls_spec, void* stream);
YH_API int yh_v1_decode(const float* y, int n, int s_h, int s_w, int b, int c,
                 float img_h, float img_w,
                 float* sig_txty, float* wh_act, float* bbox, float* conf,
                 float* cls_prob, float* cls_spec, void* stream);

/* Dense per-box target grids (the reference's get_loss arguments, models/yolov2.py:749-756)
 * -> compact records sorted by image + CSR offsets.
 *   sig_txty[M,S,S,2] twth[M,S,S,2] coord[M,S,S,4] cls_tgt[M,S,S,C] fp32;
 *   obj_mask[M,S,S] fp64 if obj_is_f64 else fp32 (the reference produces fp64, SURVEY B-7);
 *   x_img_id[N], bbox_img_id[M] int64; image of a box = first n with x_img_id[n]==bbox_img_id
 *   (0 if none, like the reference's argmax, models/yolov2.py:828-834).
 *   status[2] (device): [0] rows whose obj_mask does not have exactly one cell set,
 *                       [1] rows whose class row is not one-hot.
 *   ws: yh_compact_workspace_bytes(m, n) bytes, no initialisation needed. */
YH_API size_t yh_compact_workspace_bytes(int m, int n);
YH_API int yh_compact_targets(const float* sig_txty, const float* twth, const float* coord,
                       const float* cls_tgt, const void* obj_mask, int obj_is_f64,
                       const int64_t* x_img_id, const int64_t* bbox_img_id,
                       int m, int n, int s_h, int s_w, int c,
                       YhGt* gt_out, int32_t* gt_off_out, int32_t* status,
                       void* ws, size_t ws_bytes, void* stream);

/* Device-side target builder: pixel boxes -> compact records + CSR offsets, the float64 arithmetic
 * of collate_fn (reference models/yolov2.py:1440-1512, models/yolov1.py:1238-1312) without its
 * dense per-box grids.
 *   boxes_xyxy[M,4] float64 pixels, labels[M], img_index[M] = position of the owning image in the
 *   batch, non-decreasing (collate_fn appends boxes image by image); img_h/img_w the network input
 *   size.  gt_out[M] in the same order, gt_off_out[N+1].
 *   status[2] (device): [0] boxes out of image order, [1] boxes whose image index or cell is out of
 *   range (such records are ignored by the train head). */
YH_API int yh_build_targets(const double* boxes_xyxy, const int32_t* labels, const int32_t* img_index,
                     int m, int n, int version, int s_h, int s_w, double img_h, double img_w,
                     YhGt* gt_out, int32_t* gt_off_out, int32_t* status, void* stream);

/* Inference post-process straight from the head tensor: decode + `conf >= conf_thre` +
 * descending-confidence greedy NMS per image + class pick.
 * Replaces the predict -> nms -> argmax chain of detect() (reference models/yolov2.py:694-731,
 * models/yolov1.py:491-534) with models/utils.py:68-164 nms applied PER IMAGE.
 *   flags: bit 0 (YH_POST_CLASS_AWARE) 0 reproduces the reference (class-agnostic), 1 only lets
 *   boxes with the same argmax class suppress each other; bit 1 (YH_POST_INPUT_READY): an overlapped call
 *   under THE OVERLAP CONTRACT above (y not written, outputs and workspace not touched by any kernel launched
 *   since the last non-overlapped launch on `stream`): the kernel starts while those kernels are still
 *   draining and only waits for them before it exits.
 *   Outputs per image, first keep_cnt[n] (<= max_out) entries valid, in descending confidence:
 *   keep_idx[N,max_out] predictor index; out_bbox[N,max_out,4]; out_conf[N,max_out];
 *   out_cls_spec[N,max_out,C] (may be NULL); out_label[N,max_out]; out_score[N,max_out].
 *   keep_cnt[n] is the TRUE number kept even when it exceeds max_out (then only max_out are
 *   stored).  Ties in confidence are ordered by ascending predictor index.
 *   ws: yh_postprocess_workspace_bytes(n, P) bytes, no initialisation needed. */
YH_API size_t yh_postprocess_workspace_bytes(int n, int preds_per_image);
YH_API int yh_v2_postprocess(const float* y, int n, int s_h, int s_w, int a, int c,
                      const float* anchors_wh_host, float img_h, float img_w,
                      float conf_thre, float iou_thre, int flags, int max_out,
                      int32_t* keep_idx, int32_t* keep_cnt, float* out_bbox, float* out_conf,
                      float* out_cls_spec, int32_t* out_label, float* out_score,
                      void* ws, size_t ws_bytes, void* stream);
YH_API int yh_v1_postprocess(const float* y, int n, int s_h, int s_w, int b, int c,
                      float img_h, float img_w,
                      float conf_thre, float iou_thre, int flags, int max_out,
                      int32_t* keep_idx, int32_t* keep_cnt, float* out_bbox, float* out_conf,
                      float* out_cls_spec, int32_t* out_label, float* out_score,
                      void* ws, size_t ws_bytes, void* stream);

/* ---- the fused step: train head + post-process of the SAME head tensor in ONE kernel, y read once ---------
 * yh_v2_train followed by yh_v2_postprocess reads y twice, in two kernels.  Here one CTA per image stages the
 * image in shared memory (as the post-process does) and does BOTH: as the pieces land their candidates are
 * listed and the train head's dense pass runs over them (no-object term, dL/dy written from registers); eight
 * warps resolve the image's NMS (rank, decode, pair tests, greedy order, class pick, emit) as soon as the list
 * is complete while the others process the image's ground-truth records on top of the complete dense gradient;
 * the one-warp finalize kernel follows.  Decisions, dL/dy and detections are bit-identical to the two separate
 * calls; loss and terms agree to float rounding (the partial sums are grouped by image instead of by tile).
 * Inputs the fused kernel does not cover (unaligned y or dy, images larger than the 200 KB shared-memory
 * stage, fewer than 3 classes) run the kernels of the two separate calls instead -- same results.
 *   flags: YH_STEP_CLASS_AWARE (as YH_POST_CLASS_AWARE), YH_STEP_OVERLAPPED (THE OVERLAP CONTRACT above, for
 *   all buffers of this call including ws).
 *   xch_host: NULL on one GPU.  Sharded batches: see YhExchange -- the loss terms of all ranks are summed
 *   inside the finalize kernel over peer memory; every rank's terms/loss then hold the values of the whole batch.
 *   ws: yh_train_post_workspace_bytes() bytes, 256-byte aligned, zero-filled once when allocated (calls leave it
 *   reusable). */
#define YH_STEP_CLASS_AWARE 1
#define YH_STEP_OVERLAPPED 2

/* Peer-memory exchange of the loss sums between the ranks of a sharded batch (one process per GPU).  Every
 * rank allocates yh_exchange_bytes() bytes of device memory, zero-filled once, and maps the buffers of all
 * ranks into its own address space (CUDA IPC / peer access; odcp_b200.dist.PeerExchange does it with torch);
 * slots[q] is rank q's buffer as seen from THIS process, slots[rank] the local one.  Every rank must make the
 * same sequence of exchanging calls.  HOST struct, read during the call only. */
typedef struct YhExchange {
    int32_t rank, world;
    uint64_t* slots[YH_MAX_RANKS];
} YhExchange;
YH_API size_t yh_exchange_bytes(void);

YH_API size_t yh_train_post_workspace_bytes(int n, int s_h, int s_w, int a, int c);
YH_API int yh_v2_train_post(const float* y, int n, int s_h, int s_w, int a, int c,
                     const float* anchors_wh_host, float img_h, float img_w,
                     const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                     const float* lambdas_host, float* dy, float* terms, float* loss,
                     int32_t* resp, float* iou_resp,
                     float conf_thre, float iou_thre, int flags, int max_out,
                     int32_t* keep_idx, int32_t* keep_cnt, float* out_bbox, float* out_conf,
                     float* out_cls_spec, int32_t* out_label, float* out_score,
                     const YhExchange* xch_host, void* ws, size_t ws_bytes, void* stream);

/* yh_v2_train / yh_v1_train for a sharded batch: the same call (flags: YH_STEP_OVERLAPPED or 0) whose
 * finalize step sums the loss terms of all ranks over peer memory (xch_host as above; NULL: one GPU).
 * Replaces the all-reduce of six floats SURVEY 8(e) assigns to NCCL with 56-byte peer stores inside the
 * kernel that already ends the call. */
YH_API int yh_v2_train_sharded(const float* y, int n, int s_h, int s_w, int a, int c,
                const float* anchors_wh_host, float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, int flags, const YhExchange* xch_host,
                void* ws, size_t ws_bytes, void* stream);
YH_API int yh_v1_train_sharded(const float* y, int n, int s_h, int s_w, int b, int c,
                float img_h, float img_w,
                const YhGt* gt, const int32_t* gt_off, int m_local, int m_global,
                const float* lambdas_host, float* dy, float* terms, float* loss,
                int32_t* resp, float* iou_resp, int flags, const YhExchange* xch_host,
                void* ws, size_t ws_bytes, void* stream);

/* Batched true-positive matching for evaluation: the per-detection loop of evaluate_model
 * (reference models/utils.py:231-262).  A detection is a true positive at level L iff a ground-truth
 * box of its class in its image has get_iou(gt, det, numpy=True) >= L (float64, no one-to-one
 * assignment).
 *   det_bbox[N,max_out,4], det_label[N,max_out], keep_cnt[N]: the post-process outputs;
 *   gt_boxes_xyxy[M,4] float64, gt_labels[M], gt_off[N+1]: annotations grouped by image;
 *   levels_host[num_levels] (HOST, <= 16) IoU levels;
 *   best_iou[N,max_out] float64 (-1: no ground truth of the class, or unused slot);
 *   tp[N,max_out,num_levels] uint8. */
YH_API int yh_match_detections(const float* det_bbox, const int32_t* det_label, const int32_t* keep_cnt,
                        int n, int max_out, const double* gt_boxes_xyxy, const int32_t* gt_labels,
                        const int32_t* gt_off, const double* levels_host, int num_levels,
                        double* best_iou, unsigned char* tp, void* stream);

/* Greedy NMS on already-decoded boxes, per image: models/utils.py:68-164.
 *   bbox[N,P,4], conf[N,P]; labels[N,P] or NULL (NULL = class-agnostic, the reference).
 *   keep_idx[N,max_out], keep_cnt[N] as above.  ws: yh_postprocess_workspace_bytes(n, p). */
YH_API int yh_nms(const float* bbox, const float* conf, const int32_t* labels, int n, int p,
           float conf_thre, float iou_thre, int max_out,
           int32_t* keep_idx, int32_t* keep_cnt, void* ws, size_t ws_bytes, void* stream);

/* Elementwise IoU of `count` xyxy box pairs: models/utils.py:5-65 (torch branch). */
YH_API int yh_iou(const float* boxes1, const float* boxes2, int64_t count, float* iou, void* stream);

/* The same in float64: get_iou(..., numpy=True) keeps its inputs' dtype, and evaluate_model feeds it float64 boxes
 * (models/utils.py:30-38, 52-63, 250-252).  numpy's bits (one rounding per operation, NaNs propagate). */
YH_API int yh_iou_f64(const double* boxes1, const double* boxes2, int64_t count, double* iou, void* stream);

/* x[i] *= *scale_dev for i < count; returns immediately on the device when *scale_dev == 1
 * (the usual upstream gradient of a scalar loss). */
YH_API int yh_scale_inplace(float* x, int64_t count, const float* scale_dev, void* stream);

/* ---- fused multi-tensor SGD step ---------------------------------------
 * Replaces `opt = SGD(self.parameters(), lr, momentum=0.9, weight_decay=5e-4); ...; opt.step()` of
 * run_one_epoch (reference models/yolov2.py:1253-1272; models/yolov1.py likewise).  Per element
 * (torch.optim.SGD): d = g + weight_decay * p; buf = d on an optimizer's first step, else
 * momentum * buf + d; p -= lr * buf.
 *   p_ptrs_host / g_ptrs_host / buf_ptrs_host[n_tensors]: HOST arrays of DEVICE pointers to the fp32
 *   parameters, gradients and momentum buffers (buf_ptrs_host may be NULL, or hold NULLs: no buffer is
 *   kept for that tensor); sizes_host[n_tensors] element counts (< 2^31 each).  The list travels in the
 *   kernel's parameter space: nothing is uploaded or retained, one launch per 96 tensors.
 *   flags: YH_SGD_FRESH_MOMENTUM -- every step is a first step (buf = d), which is what the
 *   reference does by building a new optimizer in every iteration; without it buffers persist. */
#define YH_SGD_CHUNK 16384
#define YH_SGD_FRESH_MOMENTUM 1
YH_API int yh_sgd_step(const uint64_t* p_ptrs_host, const uint64_t* g_ptrs_host, const uint64_t* buf_ptrs_host,
                const int64_t* sizes_host, int n_tensors, float lr, float momentum, float weight_decay,
                int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOHEAD_H_ */
