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

    `numpy=True` takes and returns ndarrays like the reference and, like it, computes in the arrays' own
    precision: float64 boxes (what evaluate_model passes, models/utils.py:250-252) go through the float64
    kernel, whose bits are numpy's -- so true-positive decisions at the IoU levels do not move."""
    if numpy:
        dev = _device()
        a1, a2 = np.asarray(bbox_coord1), np.asarray(bbox_coord2)
        dt = np.float64 if (a1.dtype == np.float64 or a2.dtype == np.float64 or
                            not (a1.dtype == np.float32 and a2.dtype == np.float32)) else np.float32
        b1, b2 = np.broadcast_arrays(a1.astype(dt, copy=False), a2.astype(dt, copy=False))
        out = ops.iou(torch.as_tensor(np.ascontiguousarray(b1), device=dev), torch.as_tensor(np.ascontiguousarray(b2), device=dev))
        return out.cpu().numpy()
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


def _ap_table(tp, label, score, num_gt, cls_list, levels):
    """The reference's per-class AP arithmetic (models/utils.py:294-333) over all collected detections."""
    result = {"level_list": levels}
    for ci, cls in enumerate(cls_list):
        sel = label == ci
        if not sel.any():
            result[cls] = np.zeros(len(levels))
            continue
        result[cls] = average_precision(tp[sel], np.asarray(score[sel], dtype=np.float64), int(num_gt[ci]))
    return result


def evaluate_model(model, dataset, ckpt_path, conf_score_thre=0.9, iou_thre=0.5, level_list=DEFAULT_LEVELS,
                   batch_size=32):
    """Drop-in for the reference's evaluate_model (models/utils.py:171-338), same arguments and the same result
    dict {"level_list": levels, class name: AP per level}; `dataset` yields (id, image [H,W,3], annotation dict
    with "bbox_list" / "lbl_list") like data_loaders/voc.py.

    The reference walks the dataset one image at a time (detect: batch of one -> python NMS loop -> numpy
    matching).  Here consecutive images of equal size are batched: ONE post-process launch per batch
    (`model.postprocess`: decode + threshold + NMS + class pick for all images) and ONE matching launch
    (yh_match_detections: the float64 true-positive test of every detection against the ground truth of its
    class); the AP arithmetic stays the reference's numpy.  Models without the batched post-process (YOLOv1, whose
    detect rescales its boxes in float64 per image) take the reference's per-image route through `model.detect`
    and `get_iou(numpy=True)`.  A class without detections gets zeros (the reference raises on it).
    `ckpt_path` is accepted and ignored, as in the reference."""
    levels = np.asarray(level_list, dtype=np.float64)
    cls_list = list(model.cls_list)
    cls2idx = {c: i for i, c in enumerate(cls_list)}
    num_gt = np.zeros(len(cls_list), dtype=np.int64)
    tps, labels, scores = [], [], []
    batched = hasattr(model, "postprocess") and getattr(model, "_yh_version", 0) == 2
    if hasattr(model, "eval"):
        model.eval()

    def annotation(annot):
        boxes = np.asarray(annot["bbox_list"], dtype=np.float64).reshape(-1, 4)
        lbl = np.asarray([cls2idx[c] for c in annot["lbl_list"]], dtype=np.int32)
        np.add.at(num_gt, lbl, 1)
        return boxes, lbl

    def flush(batch):
        dev = _device()
        x = torch.as_tensor(np.stack([b[0] for b in batch])).to(dev)
        with torch.no_grad():
            post = model.postprocess(x, conf_score_thre, iou_thre)
        gt_boxes = np.concatenate([b[1] for b in batch], 0)
        gt_lbl = np.concatenate([b[2] for b in batch], 0)
        off = np.zeros(len(batch) + 1, dtype=np.int32)
        np.cumsum([len(b[2]) for b in batch], out=off[1:])
        _, tp = ops.match_detections(post, gt_boxes, gt_lbl, off, levels)
        max_out = tp.shape[1]
        cnt = post["keep_cnt"].clamp(max=max_out).cpu().numpy()
        valid = np.arange(max_out)[None, :] < cnt[:, None]
        tps.append(tp.cpu().numpy()[valid])
        labels.append(post["label"].cpu().numpy()[valid])
        scores.append(post["score"].cpu().numpy()[valid])

    batch = []
    for _, img, annot in dataset:
        boxes, lbl = annotation(annot)
        if batched:
            img = np.asarray(img)
            if batch and (img.shape != batch[0][0].shape or len(batch) >= batch_size):
                flush(batch)
                batch = []
            batch.append((img, boxes, lbl))
            continue
        # per-image route (the reference's loop, models/utils.py:220-262)
        pred = model.detect(img, conf_score_thre, iou_thre)
        pb = np.asarray(pred["bbox_list"], dtype=np.float64).reshape(-1, 4)
        pl = np.asarray([cls2idx[c] for c in pred["lbl_list"]], dtype=np.int32)
        ps = np.asarray(pred["cls_spec_conf_score_list"], dtype=np.float64)
        tp = np.zeros((len(pb), len(levels)), dtype=np.uint8)
        for k in range(len(pb)):
            tgt = boxes[lbl == pl[k]]
            if len(tgt):
                iou = get_iou(tgt, pb[k][None, :], numpy=True)
                tp[k] = 1 - ((iou[:, None] < levels).astype(int).prod(0) >= 1).astype(int)
        tps.append(tp)
        labels.append(pl)
        scores.append(ps)
    if batch:
        flush(batch)
    if not tps:
        return _ap_table(np.zeros((0, len(levels)), np.uint8), np.zeros(0, np.int32), np.zeros(0), num_gt, cls_list, levels)
    return _ap_table(np.concatenate(tps, 0), np.concatenate(labels, 0), np.concatenate(scores, 0), num_gt, cls_list, levels)
