"""Ground-truth target records for the YOLO head hot path (host side).

The reference hands `get_loss` one *dense* set of grids per ground-truth box
(`[M, S_h, S_w, .]`, built by `collate_fn`, reference models/yolov2.py:1440-1555 and
models/yolov1.py:1238-1355) in which exactly one cell per box is non-zero.  The
information content of a box is 12 scalars; this module defines that compact
48-byte record (`YhGt` in include/yolohead.h), the per-image CSR offsets that go
with it, and the converters between the two forms.

Record fields (all computed in float64 exactly as the reference does, then cast):
    img, cy, cx, cls            int32
    stx, sty                    sigmoid-space centre offsets inside the cell
    tw, th                      v2: box size in grid units (`bwbh`);
                                v1: box size / S (`sig_twth`)
    x1, y1, x2, y2              pixel corners
"""
from __future__ import annotations

import numpy as np
import torch

GT_DTYPE = np.dtype(
    [
        ("img", "<i4"), ("cy", "<i4"), ("cx", "<i4"), ("cls", "<i4"),
        ("stx", "<f4"), ("sty", "<f4"), ("tw", "<f4"), ("th", "<f4"),
        ("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4"),
    ]
)
assert GT_DTYPE.itemsize == 48


def boxes_to_records(boxes_xyxy, labels, img_index, height, width, s_h, s_w, version):
    """Turn pixel boxes into compact records.

    Mirrors the float64 arithmetic of the reference target builder
    (models/yolov2.py:1466-1482; v1 adds models/yolov1.py:1281-1282) so that the
    fp32 values are bit-identical to what `collate_fn` would have stored.

    boxes_xyxy: [M,4] float64 pixels, labels: [M] int, img_index: [M] int (position of the
    owning image inside the batch).  Returns a structured array sorted by image (stable).
    """
    boxes = np.asarray(boxes_xyxy, dtype=np.float64).reshape(-1, 4)
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    img_index = np.asarray(img_index, dtype=np.int64).reshape(-1)
    gh = height / s_h
    gw = width / s_w
    x1n = boxes[:, 0] / gw
    y1n = boxes[:, 1] / gh
    x2n = boxes[:, 2] / gw
    y2n = boxes[:, 3] / gh
    bx = (x1n + x2n) / 2
    by = (y1n + y2n) / 2
    bw = x2n - x1n
    bh = y2n - y1n
    cx = bx.astype(np.int64)  # int() truncation, boxes are non-negative
    cy = by.astype(np.int64)
    rec = np.zeros(len(boxes), dtype=GT_DTYPE)
    rec["img"] = img_index
    rec["cy"] = cy
    rec["cx"] = cx
    rec["cls"] = labels
    rec["stx"] = bx - cx
    rec["sty"] = by - cy
    if version == 1:
        rec["tw"] = bw / s_w
        rec["th"] = bh / s_h
    else:
        rec["tw"] = bw
        rec["th"] = bh
    rec["x1"] = boxes[:, 0]
    rec["y1"] = boxes[:, 1]
    rec["x2"] = boxes[:, 2]
    rec["y2"] = boxes[:, 3]
    order = np.argsort(rec["img"], kind="stable")
    return rec[order]


def csr_offsets(rec, num_images):
    """`gt_off[N+1]` (int32) for records sorted by image."""
    img = rec["img"].astype(np.int64)
    if len(img) > 1 and np.any(np.diff(img) < 0):
        raise ValueError("records must be sorted by image")
    if len(img) and (img.min() < 0 or img.max() >= num_images):
        raise ValueError("record image index out of range")
    counts = np.bincount(img, minlength=num_images)
    off = np.zeros(num_images + 1, dtype=np.int32)
    np.cumsum(counts, out=off[1:])
    return off


def records_to_tensor(rec, device=None, pin=False):
    """Structured records -> int32 tensor [M,12] (floats bit-cast), the layout the C ABI takes."""
    flat = np.ascontiguousarray(rec).view(np.int32).reshape(len(rec), 12)
    t = torch.from_numpy(flat.copy())
    if pin:
        t = t.pin_memory()
    if device is not None:
        t = t.to(device, non_blocking=pin)
    return t


def tensor_to_records(t):
    a = t.detach().cpu().contiguous().numpy().astype(np.int32, copy=False)
    return a.reshape(-1).view(GT_DTYPE).copy()


def records_to_dense(rec, num_images, s_h, s_w, num_cls, version=2, x_img_id=None):
    """Materialise the reference's dense `get_loss` inputs from compact records.

    Returns the 7 target tensors in the reference's order
    (sig_txty, bwbh|sig_twth, bbox_coord, cls_tgt, obj_mask[float64], x_img_id, bbox_img_id),
    with the dtypes `collate_fn` produces (models/yolov2.py:1501-1505, 1543-1544).
    """
    m = len(rec)
    sig_txty = np.zeros((m, s_h, s_w, 2), np.float32)
    twth = np.zeros((m, s_h, s_w, 2), np.float32)
    coord = np.zeros((m, s_h, s_w, 4), np.float32)
    cls_tgt = np.zeros((m, s_h, s_w, num_cls), np.float32)
    obj = np.zeros((m, s_h, s_w), np.float64)
    j = np.arange(m)
    cy, cx = rec["cy"], rec["cx"]
    sig_txty[j, cy, cx, 0] = rec["stx"]
    sig_txty[j, cy, cx, 1] = rec["sty"]
    twth[j, cy, cx, 0] = rec["tw"]
    twth[j, cy, cx, 1] = rec["th"]
    for c, k in enumerate(("x1", "y1", "x2", "y2")):
        coord[j, cy, cx, c] = rec[k]
    cls_tgt[j, cy, cx, rec["cls"]] = 1.0
    obj[j, cy, cx] = 1.0
    if x_img_id is None:
        x_img_id = np.arange(num_images, dtype=np.int64)
    x_img_id = np.asarray(x_img_id, dtype=np.int64)
    bbox_img_id = x_img_id[rec["img"]]
    return (
        torch.from_numpy(sig_txty),
        torch.from_numpy(twth),
        torch.from_numpy(coord),
        torch.from_numpy(cls_tgt),
        torch.from_numpy(obj),
        torch.from_numpy(x_img_id.copy()),
        torch.from_numpy(bbox_img_id.copy()),
    )


def shard_records(rec, gt_off, img_lo, img_hi):
    """Records and offsets of the image shard [img_lo, img_hi) with image indices rebased."""
    lo, hi = int(gt_off[img_lo]), int(gt_off[img_hi])
    sub = rec[lo:hi].copy()
    sub["img"] -= img_lo
    off = (gt_off[img_lo:img_hi + 1] - gt_off[img_lo]).astype(np.int32)
    return sub, off
