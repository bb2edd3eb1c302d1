"""`get_iou` and `nms` with the reference's signatures (reference models/utils.py:5-164),
running on the CUDA kernels.  CUDA tensors in, CUDA tensors out; there is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("odcp_b200 needs a CUDA device: this path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def get_iou(bbox_coord1, bbox_coord2, numpy=False):
    """Elementwise IoU of broadcastable [...,4] xyxy boxes -> [...] (reference models/utils.py:5-65).

    `numpy=True` takes and returns ndarrays like the reference; they are staged through the
    device (the arithmetic still runs in the CUDA kernel)."""
    if numpy:
        dev = _device()
        b1 = torch.as_tensor(np.asarray(bbox_coord1, dtype=np.float32), device=dev)
        b2 = torch.as_tensor(np.asarray(bbox_coord2, dtype=np.float32), device=dev)
        return get_iou(b1, b2).cpu().numpy()
    b1, b2 = torch.broadcast_tensors(bbox_coord1, bbox_coord2)
    return ops.iou(b1.contiguous(), b2.contiguous())


def nms(bbox_coord_batch, conf_score_batch, cls_spec_conf_score_batch, conf_score_thre=0.9, iou_thre=0.5):
    """Confidence threshold + greedy class-agnostic NMS (reference models/utils.py:68-164).

    Like the reference, ALL leading dimensions are flattened into one candidate set (so a batch
    of several images suppresses across images -- its callers pass one image), and the three
    gathered tensors come back in descending-confidence order: ([K,4], [K], [K,num_cls])."""
    num_cls = cls_spec_conf_score_batch.shape[-1]
    bbox = bbox_coord_batch.reshape(1, -1, 4).contiguous()
    conf = conf_score_batch.reshape(1, -1).contiguous()
    spec = cls_spec_conf_score_batch.reshape(-1, num_cls)
    keep_idx, keep_cnt = ops.nms_indices(bbox, conf, conf_thre=conf_score_thre, iou_thre=iou_thre)
    k = int(keep_cnt.item())  # the reference's return shapes depend on the data as well
    sel = keep_idx[0, :k].long()
    return bbox[0].index_select(0, sel), conf[0].index_select(0, sel), spec.index_select(0, sel)


DEFAULT_LEVELS = (.50, .55, .60, .65, .70, .75, .80, .85, .90, .95)  # reference models/utils.py:177


def average_precision(tp, scores, num_gt, eps=1e-6):
    """AP per IoU level from true-positive flags `tp` [K,L] and class scores [K] of one class, the
    reference's way (models/utils.py:294-333): descending score order, cumulative precision/recall,
    precision envelope from the right, sum of envelope times recall increments."""
    tp = np.asarray(tp, dtype=np.int64).reshape(len(scores), -1)
    order = np.argsort(np.asarray(scores))[::-1]
    tp = tp[order]
    ctp = np.cumsum(tp, axis=0)
    cfp = np.cumsum(1 - tp, axis=0)
    prec = ctp / (ctp + cfp + eps)
    rec = ctp / (num_gt + eps)
    envelope = np.maximum.accumulate(prec[::-1], axis=0)[::-1]
    drec = rec - np.concatenate([np.zeros_like(rec[:1]), rec[:-1]], axis=0)
    return np.sum(envelope * drec, axis=0)


def evaluate_detections(post, gt_boxes_xyxy, gt_labels, gt_off, cls_list, level_list=DEFAULT_LEVELS):
    """Batched counterpart of evaluate_model's inner loops (reference models/utils.py:171-338): `post`
    is a batched postprocess()/detect result for N images, the ground truth comes as float64 boxes
    [M,4], class indices [M] and per-image offsets [N+1].  The true-positive matching runs on the device
    (yh_match_detections); the AP arithmetic is the reference's numpy, on the host.  Returns the
    reference's dict: {"level_list": levels, class name: AP per level}.  A class without detections gets
    zeros (the reference raises on it)."""
    levels = np.asarray(level_list, dtype=np.float64)
    _, tp = ops.match_detections(post, gt_boxes_xyxy, gt_labels, gt_off, levels)
    n, max_out = tp.shape[0], tp.shape[1]
    cnt = post["keep_cnt"].clamp(max=max_out).cpu().numpy()
    valid = np.arange(max_out)[None, :] < cnt[:, None]
    tp = tp.cpu().numpy()[valid]
    label = post["label"].cpu().numpy()[valid]
    score = post["score"].cpu().numpy()[valid]
    gt_labels = np.asarray(torch.as_tensor(gt_labels).cpu())
    result = {"level_list": levels}
    for ci, cls in enumerate(cls_list):
        sel = label == ci
        if not sel.any():
            result[cls] = np.zeros(len(levels))
            continue
        result[cls] = average_precision(tp[sel], score[sel], int(np.sum(gt_labels == ci)))
    return result
