"""Tensor-level wrappers over the C ABI (include/yolohead.h).

torch is used here only for device memory and streams: every function takes CUDA tensors,
hands their raw pointers and the current stream to libyolohead, and returns CUDA tensors.
Nothing here computes on the CPU and nothing falls back when the library is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

LAMBDA_KEYS = ("lambda_xy", "lambda_wh", "lambda_conf", "lambda_noobj", "lambda_cls")

_ws_cache = {}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise ValueError("%s must be a CUDA tensor (this path has no CPU implementation)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _host_floats(values):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr


def _anchors_host(anchors):
    flat = [float(v) for wh in anchors for v in wh]
    return _host_floats(flat)


def _train_workspace(device):
    """Zero-initialised scratch, one per (device, stream): the kernels keep it zeroed."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = int(_lib.load().yh_train_workspace_bytes())
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def head_shape(y, version, a=None):
    """(n, s_h, s_w, a, c) from a head tensor."""
    if version == 2:
        if y.dim() != 5:
            raise ValueError("YOLOv2 head tensor must be [N,S_h,S_w,A,5+C], got %s" % (tuple(y.shape),))
        n, s_h, s_w, a_, d = y.shape
        return n, s_h, s_w, a_, d - 5
    if y.dim() != 4 or a is None:
        raise ValueError("YOLOv1 head tensor must be [N,S_h,S_w,5B+C] with B given, got %s" % (tuple(y.shape),))
    n, s_h, s_w, d = y.shape
    return n, s_h, s_w, a, d - 5 * a


def _exchange_arg(exchange):
    """ctypes pointer to the YhExchange of a dist.PeerExchange (None: single GPU)."""
    if exchange is None:
        return None
    return C.byref(exchange.struct)


def train_head(y, gt, gt_off, *, version, img_hw, lambdas, anchors=None, boxes_per_cell=None,
               m_global=None, want_grad=True, want_resp=False, out=None, input_ready=False, exchange=None):
    """Fused decode + assignment + loss (+ dL/dy) -- yh_v2_train / yh_v1_train.

    y       head tensor (CUDA fp32), gt int32 [M,12] records sorted by image (targets.py),
    gt_off  int32 [N+1] CSR offsets, img_hw (H, W) of the network input,
    lambdas dict with the five reference weights or a sequence of five floats.
    Returns dict(loss 0-dim, terms[5], dy or None, resp/iou_resp or None).
    `input_ready=True` is the overlap contract of include/yolohead.h (yh_v*_train_overlapped): NONE of this
    call's tensors -- y, gt, gt_off, the outputs in `out` and the workspace `out["_tws"]` -- is read or written
    by ANY kernel launched on the current stream since the last call that was not overlapped; the kernel then
    runs next to the tails of the kernels in front of it.  Overlapped calls need their own workspace per buffer
    set: pass the same `out` dict again for the same set (it then carries `_tws`).
    `exchange`: a dist.PeerExchange -- the loss terms of all ranks are summed inside the finalize kernel over
    peer memory (yh_v*_train_sharded); terms/loss then hold the values of the whole sharded batch.
    """
    y = _require_cuda_f32(y, "y")
    n, s_h, s_w, a, c = head_shape(y, version, boxes_per_cell)
    dev = y.device
    if gt.device != dev or gt_off.device != dev:
        raise ValueError("gt / gt_off must live on the same device as y")
    if gt.dtype != torch.int32 or gt.dim() != 2 or gt.shape[1] != 12 or not gt.is_contiguous():
        raise ValueError("gt must be a contiguous int32 [M,12] record tensor")
    if gt_off.dtype != torch.int32 or gt_off.numel() != n + 1 or not gt_off.is_contiguous():
        raise ValueError("gt_off must be a contiguous int32 [N+1] tensor")
    m_local = int(gt.shape[0])
    m_glob = m_local if m_global is None else int(m_global)
    lam = [lambdas[k] for k in LAMBDA_KEYS] if isinstance(lambdas, dict) else list(lambdas)
    out = out if out is not None else {}
    with torch.cuda.device(dev):
        if input_ready:  # an overlapped call must not share its workspace with the calls it overlaps
            ws = out.get("_tws")
            if ws is None:
                ws = out["_tws"] = torch.zeros(int(_lib.load().yh_train_workspace_bytes()), dtype=torch.uint8, device=dev)
        else:
            ws = _train_workspace(dev)
        dy = out.get("dy") if want_grad else None
        if want_grad and dy is None:
            dy = torch.empty_like(y)
        terms = out.get("terms")
        if terms is None:
            terms = torch.empty(5, dtype=torch.float32, device=dev)
        loss = out.get("loss")
        if loss is None:
            loss = torch.empty((), dtype=torch.float32, device=dev)
        resp = iou_resp = None
        if want_resp:
            resp = out.get("resp")
            if resp is None:
                resp = torch.empty(m_local, dtype=torch.int32, device=dev)
            iou_resp = out.get("iou_resp")
            if iou_resp is None:
                iou_resp = torch.empty(m_local, dtype=torch.float32, device=dev)
        lam_h = _host_floats(lam)
        if version == 2:
            if anchors is None or len(anchors) != a:
                raise ValueError("anchors must list %d (w,h) pairs" % a)
            if exchange is not None:
                _lib.call("yh_v2_train_sharded", _ptr(y), n, s_h, s_w, a, c, _anchors_host(anchors),
                          float(img_hw[0]), float(img_hw[1]), _ptr(gt), _ptr(gt_off), m_local, m_glob,
                          lam_h, _ptr(dy), _ptr(terms), _ptr(loss), _ptr(resp), _ptr(iou_resp),
                          STEP_OVERLAPPED if input_ready else 0, _exchange_arg(exchange), _ptr(ws), ws.numel(), _stream())
            else:
                _lib.call("yh_v2_train_overlapped" if input_ready else "yh_v2_train", _ptr(y), n, s_h, s_w, a, c, _anchors_host(anchors),
                          float(img_hw[0]), float(img_hw[1]), _ptr(gt), _ptr(gt_off), m_local, m_glob,
                          lam_h, _ptr(dy), _ptr(terms), _ptr(loss), _ptr(resp), _ptr(iou_resp),
                          _ptr(ws), ws.numel(), _stream())
        else:
            if exchange is not None:
                _lib.call("yh_v1_train_sharded", _ptr(y), n, s_h, s_w, a, c,
                          float(img_hw[0]), float(img_hw[1]), _ptr(gt), _ptr(gt_off), m_local, m_glob,
                          lam_h, _ptr(dy), _ptr(terms), _ptr(loss), _ptr(resp), _ptr(iou_resp),
                          STEP_OVERLAPPED if input_ready else 0, _exchange_arg(exchange), _ptr(ws), ws.numel(), _stream())
            else:
                _lib.call("yh_v1_train_overlapped" if input_ready else "yh_v1_train", _ptr(y), n, s_h, s_w, a, c,
                          float(img_hw[0]), float(img_hw[1]), _ptr(gt), _ptr(gt_off), m_local, m_glob,
                          lam_h, _ptr(dy), _ptr(terms), _ptr(loss), _ptr(resp), _ptr(iou_resp),
                          _ptr(ws), ws.numel(), _stream())
    res = dict(loss=loss, terms=terms, dy=dy, resp=resp, iou_resp=iou_resp)
    if input_ready:
        res["_tws"] = ws
    return res


STEP_CLASS_AWARE, STEP_OVERLAPPED = 1, 2


def train_post(y, gt, gt_off, *, img_hw, lambdas, anchors, conf_thre, iou_thre, m_global=None, want_grad=True,
               want_resp=False, class_aware=False, max_out=None, want_cls_spec=True, out=None, overlapped=False,
               exchange=None):
    """The fused step of a YOLOv2 head -- yh_v2_train_post: train head (loss, terms, dL/dy) AND post-process
    (threshold + per-image greedy NMS + class pick) of the same head tensor in ONE kernel that reads it once
    (one CTA per image holds the image in shared memory and does both).  Decisions, dL/dy and detections are
    bit-identical to train_head(...) followed by postprocess(...); loss/terms agree to float rounding.

    Returns dict(train=<train_head dict>, post=<postprocess dict>, _ws=workspace).  Pass the returned dict as
    `out` to reuse every buffer (pipelines, CUDA graphs).  `overlapped=True`: the overlap contract of
    include/yolohead.h for ALL of this call's tensors including the workspace -- rotate complete `out` sets.
    `exchange`: a dist.PeerExchange for a sharded batch (terms/loss of the whole batch on every rank).
    """
    y = _require_cuda_f32(y, "y")
    n, s_h, s_w, a, c = head_shape(y, 2)
    dev = y.device
    if gt.device != dev or gt_off.device != dev:
        raise ValueError("gt / gt_off must live on the same device as y")
    if gt.dtype != torch.int32 or gt.dim() != 2 or gt.shape[1] != 12 or not gt.is_contiguous():
        raise ValueError("gt must be a contiguous int32 [M,12] record tensor")
    if gt_off.dtype != torch.int32 or gt_off.numel() != n + 1 or not gt_off.is_contiguous():
        raise ValueError("gt_off must be a contiguous int32 [N+1] tensor")
    if anchors is None or len(anchors) != a:
        raise ValueError("anchors must list %d (w,h) pairs" % a)
    m_local = int(gt.shape[0])
    m_glob = m_local if m_global is None else int(m_global)
    lam = [lambdas[k] for k in LAMBDA_KEYS] if isinstance(lambdas, dict) else list(lambdas)
    p = s_h * s_w * a
    max_out = p if max_out is None else int(max_out)
    out = out if out is not None else {}
    tr, po = out.get("train") or {}, out.get("post")
    with torch.cuda.device(dev):
        ws = out.get("_ws")
        if ws is None:
            nbytes = int(_lib.load().yh_train_post_workspace_bytes(n, s_h, s_w, a, c))
            if nbytes == 0:
                raise ValueError("unsupported head geometry for the fused step")
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        dy = tr.get("dy") if want_grad else None
        if want_grad and dy is None:
            dy = torch.empty_like(y)
        terms = tr.get("terms") if tr.get("terms") is not None else torch.empty(5, dtype=torch.float32, device=dev)
        loss = tr.get("loss") if tr.get("loss") is not None else torch.empty((), dtype=torch.float32, device=dev)
        resp = iou_resp = None
        if want_resp:
            resp = tr.get("resp") if tr.get("resp") is not None else torch.empty(m_local, dtype=torch.int32, device=dev)
            iou_resp = tr.get("iou_resp") if tr.get("iou_resp") is not None else torch.empty(m_local, dtype=torch.float32, device=dev)
        if po is not None:
            if po["keep_idx"].shape != (n, max_out) or po["keep_idx"].device != dev:
                raise ValueError("`out` does not match this call's shapes")
        else:
            po = dict(keep_idx=torch.empty(n, max_out, dtype=torch.int32, device=dev),
                      keep_cnt=torch.empty(n, dtype=torch.int32, device=dev),
                      bbox=torch.empty(n, max_out, 4, dtype=torch.float32, device=dev),
                      conf=torch.empty(n, max_out, dtype=torch.float32, device=dev),
                      cls_spec=torch.empty(n, max_out, c, dtype=torch.float32, device=dev) if want_cls_spec else None,
                      label=torch.empty(n, max_out, dtype=torch.int32, device=dev),
                      score=torch.empty(n, max_out, dtype=torch.float32, device=dev))
        flags = (STEP_CLASS_AWARE if class_aware else 0) | (STEP_OVERLAPPED if overlapped else 0)
        _lib.call("yh_v2_train_post", _ptr(y), n, s_h, s_w, a, c, _anchors_host(anchors),
                  float(img_hw[0]), float(img_hw[1]), _ptr(gt), _ptr(gt_off), m_local, m_glob,
                  _host_floats(lam), _ptr(dy), _ptr(terms), _ptr(loss), _ptr(resp), _ptr(iou_resp),
                  float(conf_thre), float(iou_thre), flags, max_out, _ptr(po["keep_idx"]), _ptr(po["keep_cnt"]),
                  _ptr(po["bbox"]), _ptr(po["conf"]), _ptr(po["cls_spec"]), _ptr(po["label"]), _ptr(po["score"]),
                  _exchange_arg(exchange), _ptr(ws), ws.numel(), _stream())
    return dict(train=dict(loss=loss, terms=terms, dy=dy, resp=resp, iou_resp=iou_resp), post=po, _ws=ws)


def decode(y, *, version, img_hw, anchors=None, boxes_per_cell=None):
    """The six predict() outputs -- yh_v2_decode / yh_v1_decode."""
    y = _require_cuda_f32(y, "y")
    n, s_h, s_w, a, c = head_shape(y, version, boxes_per_cell)
    dev = y.device
    kw = dict(dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        sig_txty = torch.empty(n, s_h, s_w, a, 2, **kw)
        wh_act = torch.empty(n, s_h, s_w, a, 2, **kw)
        bbox = torch.empty(n, s_h, s_w, a, 4, **kw)
        conf = torch.empty(n, s_h, s_w, a, **kw)
        spec = torch.empty(n, s_h, s_w, a, c, **kw)
        if version == 2:
            prob = torch.empty(n, s_h, s_w, a, c, **kw)
            _lib.call("yh_v2_decode", _ptr(y), n, s_h, s_w, a, c, _anchors_host(anchors),
                      float(img_hw[0]), float(img_hw[1]), _ptr(sig_txty), _ptr(wh_act), _ptr(bbox),
                      _ptr(conf), _ptr(prob), _ptr(spec), _stream())
        else:
            prob = torch.empty(n, s_h, s_w, c, **kw)
            _lib.call("yh_v1_decode", _ptr(y), n, s_h, s_w, a, c,
                      float(img_hw[0]), float(img_hw[1]), _ptr(sig_txty), _ptr(wh_act), _ptr(bbox),
                      _ptr(conf), _ptr(prob), _ptr(spec), _stream())
    return sig_txty, wh_act, bbox, conf, prob, spec


def compact_targets(sig_txty, twth, coord, cls_tgt, obj_mask, x_img_id, bbox_img_id):
    """Dense reference targets -> (gt [M,12] int32, gt_off [N+1] int32, status [2] int32)."""
    dev = sig_txty.device
    m, s_h, s_w, _ = sig_txty.shape
    n = int(x_img_id.numel())
    c = int(cls_tgt.shape[-1])
    sig_txty = _require_cuda_f32(sig_txty, "sig_txty_tgt")
    twth = _require_cuda_f32(twth, "wh_tgt")
    coord = _require_cuda_f32(coord, "bbox_coord_tgt")
    cls_tgt = _require_cuda_f32(cls_tgt, "cls_tgt")
    if obj_mask.dtype not in (torch.float64, torch.float32):
        obj_mask = obj_mask.to(torch.float32)
    obj_mask = obj_mask.contiguous()
    x_img_id = x_img_id.to(device=dev, dtype=torch.int64).contiguous()
    bbox_img_id = bbox_img_id.to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        gt = torch.empty(m, 12, dtype=torch.int32, device=dev)
        gt_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
        status = torch.empty(2, dtype=torch.int32, device=dev)
        nbytes = int(_lib.load().yh_compact_workspace_bytes(m, n))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _lib.call("yh_compact_targets", _ptr(sig_txty), _ptr(twth), _ptr(coord), _ptr(cls_tgt),
                  _ptr(obj_mask), 1 if obj_mask.dtype == torch.float64 else 0, _ptr(x_img_id),
                  _ptr(bbox_img_id), m, n, s_h, s_w, c, _ptr(gt), _ptr(gt_off), _ptr(status),
                  _ptr(ws), ws.numel(), _stream())
    return gt, gt_off, status


def build_targets(boxes_xyxy, labels, img_index, *, num_images, version, img_hw, grid):
    """Pixel boxes -> (gt [M,12] int32 records, gt_off [N+1] int32, status [2] int32) on the device --
    yh_build_targets.  boxes_xyxy float64 [M,4], labels / img_index integer [M] with img_index
    non-decreasing (boxes grouped by image, the order collate_fn produces)."""
    if not (isinstance(boxes_xyxy, torch.Tensor) and boxes_xyxy.is_cuda):
        raise ValueError("boxes_xyxy must be a CUDA tensor (this path has no CPU implementation)")
    dev = boxes_xyxy.device
    boxes = boxes_xyxy.to(torch.float64).reshape(-1, 4).contiguous()
    m = int(boxes.shape[0])
    labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    img_index = img_index.to(device=dev, dtype=torch.int32).contiguous()
    if labels.numel() != m or img_index.numel() != m:
        raise ValueError("labels / img_index must have one entry per box")
    with torch.cuda.device(dev):
        gt = torch.empty(m, 12, dtype=torch.int32, device=dev)
        gt_off = torch.empty(int(num_images) + 1, dtype=torch.int32, device=dev)
        status = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.call("yh_build_targets", _ptr(boxes), _ptr(labels), _ptr(img_index), m, int(num_images),
                  int(version), int(grid[0]), int(grid[1]), float(img_hw[0]), float(img_hw[1]),
                  _ptr(gt), _ptr(gt_off), _ptr(status), _stream())
    return gt, gt_off, status


def postprocess(y, *, version, img_hw, conf_thre, iou_thre, anchors=None, boxes_per_cell=None,
                class_aware=False, max_out=None, want_cls_spec=True, out=None, input_ready=False):
    """Decode + threshold + per-image greedy NMS + class pick -- yh_v{1,2}_postprocess.

    Returns dict(keep_idx [N,max_out] int32, keep_cnt [N] int32, bbox, conf, cls_spec, label, score).
    `out` may be a dict returned by an earlier call with the same shapes: its tensors are reused
    (no allocation, e.g. inside a pipeline or a CUDA graph).  `input_ready=True` promises that y was
    not written by the kernel launched just before this call on the current stream (e.g. it is the
    train head, which only reads y): the kernel then overlaps that kernel's tail (YH_POST_INPUT_READY).
    """
    y = _require_cuda_f32(y, "y")
    n, s_h, s_w, a, c = head_shape(y, version, boxes_per_cell)
    p = s_h * s_w * a
    max_out = p if max_out is None else int(max_out)
    dev = y.device
    with torch.cuda.device(dev):
        if out is not None:
            keep_idx, keep_cnt, bbox, conf = out["keep_idx"], out["keep_cnt"], out["bbox"], out["conf"]
            spec, label, score, ws = out["cls_spec"], out["label"], out["score"], out["_ws"]
            if keep_idx.shape != (n, max_out) or keep_idx.device != dev:
                raise ValueError("`out` does not match this call's shapes")
        else:
            keep_idx = torch.empty(n, max_out, dtype=torch.int32, device=dev)
            keep_cnt = torch.empty(n, dtype=torch.int32, device=dev)
            bbox = torch.empty(n, max_out, 4, dtype=torch.float32, device=dev)
            conf = torch.empty(n, max_out, dtype=torch.float32, device=dev)
            spec = torch.empty(n, max_out, c, dtype=torch.float32, device=dev) if want_cls_spec else None
            label = torch.empty(n, max_out, dtype=torch.int32, device=dev)
            score = torch.empty(n, max_out, dtype=torch.float32, device=dev)
            nbytes = int(_lib.load().yh_postprocess_workspace_bytes(n, p))
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        if version == 2:
            _lib.call("yh_v2_postprocess", _ptr(y), n, s_h, s_w, a, c, _anchors_host(anchors),
                      float(img_hw[0]), float(img_hw[1]), float(conf_thre), float(iou_thre),
                      int(bool(class_aware)) | (2 if input_ready else 0), max_out, _ptr(keep_idx), _ptr(keep_cnt), _ptr(bbox),
                      _ptr(conf), _ptr(spec), _ptr(label), _ptr(score), _ptr(ws), ws.numel(), _stream())
        else:
            _lib.call("yh_v1_postprocess", _ptr(y), n, s_h, s_w, a, c,
                      float(img_hw[0]), float(img_hw[1]), float(conf_thre), float(iou_thre),
                      int(bool(class_aware)) | (2 if input_ready else 0), max_out, _ptr(keep_idx), _ptr(keep_cnt), _ptr(bbox),
                      _ptr(conf), _ptr(spec), _ptr(label), _ptr(score), _ptr(ws), ws.numel(), _stream())
    return dict(keep_idx=keep_idx, keep_cnt=keep_cnt, bbox=bbox, conf=conf, cls_spec=spec,
                label=label, score=score, _ws=ws)


def match_detections(post, gt_boxes_xyxy, gt_labels, gt_off, levels):
    """True-positive flags of the detections in `post` (a postprocess() result) against float64
    ground-truth boxes grouped by image -- yh_match_detections.  Returns (best_iou [N,max_out] float64,
    tp [N,max_out,L] uint8)."""
    bbox, label, cnt = post["bbox"], post["label"], post["keep_cnt"]
    dev = bbox.device
    n, max_out = int(bbox.shape[0]), int(bbox.shape[1])
    gt_boxes = torch.as_tensor(gt_boxes_xyxy, dtype=torch.float64).to(dev).reshape(-1, 4).contiguous()
    gt_labels = torch.as_tensor(gt_labels).to(device=dev, dtype=torch.int32).contiguous()
    gt_off = torch.as_tensor(gt_off).to(device=dev, dtype=torch.int32).contiguous()
    if gt_off.numel() != n + 1:
        raise ValueError("gt_off must have N+1 entries")
    lv = [float(v) for v in levels]
    lv_h = (C.c_double * len(lv))(*lv)
    with torch.cuda.device(dev):
        best = torch.empty(n, max_out, dtype=torch.float64, device=dev)
        tp = torch.empty(n, max_out, len(lv), dtype=torch.uint8, device=dev)
        _lib.call("yh_match_detections", _ptr(bbox), _ptr(label), _ptr(cnt), n, max_out, _ptr(gt_boxes),
                  _ptr(gt_labels), _ptr(gt_off), lv_h, len(lv), _ptr(best), _ptr(tp), _stream())
    return best, tp


def nms_indices(bbox, conf, *, conf_thre, iou_thre, labels=None, max_out=None):
    """Per-image greedy NMS on decoded boxes: bbox [N,P,4], conf [N,P] -> (keep_idx, keep_cnt)."""
    bbox = _require_cuda_f32(bbox, "bbox")
    conf = _require_cuda_f32(conf, "conf")
    n, p = conf.shape
    max_out = p if max_out is None else int(max_out)
    dev = bbox.device
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        keep_idx = torch.empty(n, max(max_out, 1), dtype=torch.int32, device=dev)
        keep_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        nbytes = int(_lib.load().yh_postprocess_workspace_bytes(n, p))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _lib.call("yh_nms", _ptr(bbox), _ptr(conf), _ptr(labels), n, p, float(conf_thre),
                  float(iou_thre), max_out, _ptr(keep_idx), _ptr(keep_cnt), _ptr(ws), ws.numel(),
                  _stream())
    return keep_idx, keep_cnt


def iou(boxes1, boxes2):
    """Elementwise IoU of two [K,4] CUDA tensors -- yh_iou (float32) / yh_iou_f64 (float64): the result keeps
    the inputs' dtype, like the reference's get_iou."""
    if not (isinstance(boxes1, torch.Tensor) and boxes1.is_cuda and isinstance(boxes2, torch.Tensor) and boxes2.is_cuda):
        raise ValueError("boxes must be CUDA tensors (this path has no CPU implementation)")
    if boxes1.shape != boxes2.shape or boxes1.shape[-1] != 4:
        raise ValueError("boxes must have identical [...,4] shapes")
    f64 = boxes1.dtype == torch.float64 or boxes2.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    boxes1, boxes2 = boxes1.to(dt).contiguous(), boxes2.to(dt).contiguous()
    count = boxes1.numel() // 4
    with torch.cuda.device(boxes1.device):
        out = torch.empty(boxes1.shape[:-1], dtype=dt, device=boxes1.device)
        _lib.call("yh_iou_f64" if f64 else "yh_iou", _ptr(boxes1), _ptr(boxes2), count, _ptr(out), _stream())
    return out


def scale_inplace(x, scale):
    """x *= scale (0-dim CUDA tensor); a device-side no-op when scale == 1."""
    with torch.cuda.device(x.device):
        _lib.call("yh_scale_inplace", _ptr(x), x.numel(), _ptr(scale), _stream())
    return x
