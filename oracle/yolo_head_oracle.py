"""CPU oracle for the YOLO detection-head hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference computes for
decode -> responsible-predictor assignment -> five-term loss (+ gradient) -> greedy NMS.
Only `tests/`, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference`
legs may import it; the product package never does (the product path raises when the CUDA
library is missing).

Parity status: PINNED.  The reference ships no golden vectors (it has no tests at all), so the
oracle is pinned by running the reference's own functions in the build container
(tests/golden/make_golden.py imports /root/reference with an albumentations stub) and
committing their outputs under tests/golden/*.npz; tests/test_oracle_golden.py checks both
tiers of this oracle against those files.

Two tiers:
  * "dense"   -- torch-CPU ops in the same order the reference applies them, including its
                 per-GT-box replication of the predictions.  Bit-identical to the reference on
                 the same torch build; gradients come from autograd, like the reference.
                 Memory grows as M*S*S*A, so it is only usable up to a few thousand boxes.
                 This tier is also what bench.py times as the reference's CPU cost ("port").
  * "compact" -- per-GT-box closed form in numpy (float32 op-by-op for the decisions, float64
                 for the sums), memory-light, for the large configurations.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

F32 = np.float32


# --------------------------------------------------------------------------------------
# dense tier (torch, autograd)
# --------------------------------------------------------------------------------------
def iou_torch(b1, b2):
    """Elementwise xyxy IoU with broadcasting -- models/utils.py:25-63 (torch branch)."""
    ax1, ay1, ax2, ay2 = b1[..., 0], b1[..., 1], b1[..., 2], b1[..., 3]
    bx1, by1, bx2, by2 = b2[..., 0], b2[..., 1], b2[..., 2], b2[..., 3]
    iw = torch.clamp(torch.minimum(ax2, bx2) - torch.maximum(ax1, bx1), min=0)
    ih = torch.clamp(torch.minimum(ay2, by2) - torch.maximum(ay1, by1), min=0)
    inter = iw * ih
    union = (ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter
    return inter / (union + 1e-6)


def decode_torch(y, height, width, version, anchors=None):
    """The six `predict` outputs from a head tensor.

    v2: models/yolov2.py:469-649 (y is [N,Sh,Sw,A,5+C]);
    v1: models/yolov1.py:250-437 (y is [N,Sh,Sw,5B+C], `a` boxes per cell given by `anchors`
        being an int).
    Returns (sig_txty, wh_act, bbox_px, conf, cls_prob, cls_spec) where wh_act is exp(twth)
    for v2 and sigmoid(twth) for v1.
    """
    if version == 2:
        n, sh, sw, a, _ = y.shape
        box = y
        cls_logits = y[..., 5:]
        pw = torch.tensor([p[0] for p in anchors])[None, None, None]
        ph = torch.tensor([p[1] for p in anchors])[None, None, None]
    else:
        n, sh, sw, _ = y.shape
        a = int(anchors)
        box = y[..., : a * 5].reshape(n, sh, sw, a, 5)
        c = y.shape[-1] - 5 * a
        cls_logits = y[..., -c:]
    sx = torch.sigmoid(box[..., 0])
    sy = torch.sigmoid(box[..., 1])
    if version == 2:
        aw = torch.exp(box[..., 2])
        ah = torch.exp(box[..., 3])
        bw = pw * aw
        bh = ph * ah
    else:
        aw = torch.sigmoid(box[..., 2])
        ah = torch.sigmoid(box[..., 3])
        bw = sw * aw
        bh = sh * ah
    cy = torch.arange(sh)[None, :, None, None]
    cx = torch.arange(sw)[None, None, :, None]
    bx = sx + cx
    by = sy + cy
    x1 = bx - (bw / 2)
    y1 = by - (bh / 2)
    x2 = bx + (bw / 2)
    y2 = by + (bh / 2)
    gh = height / sh
    gw = width / sw
    bbox = torch.stack([x1 * gw, y1 * gh, x2 * gw, y2 * gh], dim=-1)
    conf = torch.sigmoid(box[..., 4])
    prob = torch.softmax(cls_logits, dim=-1)
    if version == 2:
        spec = prob * conf.unsqueeze(-1)
    else:
        spec = prob.unsqueeze(-2) * conf.unsqueeze(-1)
    return (torch.stack([sx, sy], -1), torch.stack([aw, ah], -1), bbox, conf, prob, spec)


def loss_dense_torch(y, height, width, version, anchors, sig_txty_t, twth_t, coord_t, cls_t,
                     obj_mask, x_img_id, bbox_img_id, lambdas):
    """`get_loss` on dense per-box targets -- models/yolov2.py:820-1140, models/yolov1.py:629-931.

    Returns (loss, terms[5], iou[M,Sh,Sw,A], responsible_bool[M,Sh,Sw,A]); `loss` carries the
    autograd graph back to `y`.
    """
    mapper = (bbox_img_id[:, None] == x_img_id[None, :]).long().argmax(-1)
    sig_txty, wh_act, bbox, conf, prob, _ = decode_torch(y, height, width, version, anchors)
    sig_txty, wh_act, bbox, conf, prob = (t[mapper] for t in (sig_txty, wh_act, bbox, conf, prob))
    a = conf.shape[-1]
    sqrt_wh = torch.sqrt(wh_act)
    txty_t = sig_txty_t.unsqueeze(-2)
    if version == 2:
        pwph = torch.tensor([[p[0], p[1]] for p in anchors])[None, None, None]
        sqrt_wh_t = torch.sqrt(twth_t.unsqueeze(-2) / pwph)
    else:
        sqrt_wh_t = torch.sqrt(twth_t.unsqueeze(-2))
    iou = iou_torch(bbox, coord_t.unsqueeze(-2)).detach()
    resp = torch.nn.functional.one_hot(iou.max(dim=-1)[1], a) * obj_mask.unsqueeze(-1)
    not_resp = (resp != 1)
    resp = resp.bool()
    sq = lambda t, p: (p - t) ** 2  # mse, reduction none
    l_xy = torch.masked_select(sq(txty_t, sig_txty), resp.unsqueeze(-1)).mean()
    l_wh = torch.masked_select(sq(sqrt_wh_t, sqrt_wh), resp.unsqueeze(-1)).mean()
    l_conf = torch.masked_select(sq(iou, conf), resp).mean()
    l_noobj = torch.masked_select(conf ** 2, not_resp).mean()
    if version == 2:
        l_cls = torch.masked_select(sq(cls_t.unsqueeze(-2), prob).sum(-1), resp).mean()
    else:
        l_cls = torch.masked_select(sq(cls_t, prob).sum(-1), obj_mask.bool()).mean()
    terms = [l_xy, l_wh, l_conf, l_noobj, l_cls]
    lam = [lambdas[k] for k in ("lambda_xy", "lambda_wh", "lambda_conf", "lambda_noobj", "lambda_cls")]
    loss = lam[0] * l_xy + lam[1] * l_wh + lam[2] * l_conf + lam[3] * l_noobj + lam[4] * l_cls
    return loss, terms, iou, resp


def train_head_dense(case, lambdas, chunk_images=None):
    """Loss, five terms, dL/dy, responsible index and its IoU per GT record, dense tier.

    With `chunk_images` the batch is evaluated in image chunks and the means are recombined
    exactly (every term is sum/count with known counts -- SURVEY 8(c)); this is how the
    replicated formulation is made to fit in memory for larger batches.
    """
    n, m = case.n, case.m
    y = case.y.clone().requires_grad_(True)
    anchors = case.anchors if case.version == 2 else case.a
    preds = case.s_h * case.s_w * case.a
    denom = np.array([2 * m, 2 * m, m, m * (preds - 1), m], dtype=np.float64)
    step = n if not chunk_images else chunk_images
    total_terms = np.zeros(5, dtype=np.float64)
    resp_idx = np.zeros(m, dtype=np.int32)
    resp_iou = np.zeros(m, dtype=np.float32)
    lam = [lambdas[k] for k in ("lambda_xy", "lambda_wh", "lambda_conf", "lambda_noobj", "lambda_cls")]
    loss32 = 0.0
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        rec, off = shard_records(case.rec, case.gt_off, lo, hi)
        mc = len(rec)
        if mc == 0:
            continue
        dense = records_to_dense(rec, hi - lo, case.s_h, case.s_w, case.c)
        ysub = y[lo:hi]
        _, terms, iou, resp = loss_dense_torch(ysub, case.height, case.width, case.version,
                                               anchors, *dense, lambdas)
        cnt = np.array([2 * mc, 2 * mc, mc, mc * (preds - 1), mc], dtype=np.float64)
        part = sum(lam[i] * terms[i] * float(cnt[i] / denom[i]) for i in range(5))
        part.backward()
        loss32 = loss32 + float(part.detach())
        total_terms += np.array([float(t.detach()) for t in terms]) * cnt / denom
        j = np.arange(mc)
        r = resp[j, rec["cy"], rec["cx"]].int().argmax(-1).numpy()
        g0 = int(case.gt_off[lo])
        resp_idx[g0:g0 + mc] = r
        resp_iou[g0:g0 + mc] = iou[j, rec["cy"], rec["cx"], r].numpy()
    # unchunked: `loss32` is the reference's own fp32 scalar, bit for bit
    loss = loss32 if step >= n else float(sum(lam[i] * total_terms[i] for i in range(5)))
    return dict(loss=loss, terms=total_terms, dy=y.grad.detach().numpy(), resp=resp_idx,
                iou_resp=resp_iou)


def shard_records(rec, gt_off, img_lo, img_hi):
    """Records and CSR offsets of the image range [img_lo, img_hi), image indices rebased (the oracle's own helper:
    test infrastructure does not lean on the product package for its inputs)."""
    lo, hi = int(gt_off[img_lo]), int(gt_off[img_hi])
    sub = rec[lo:hi].copy()
    sub["img"] -= img_lo
    return sub, (np.asarray(gt_off[img_lo:img_hi + 1]) - gt_off[img_lo]).astype(np.int32)


def records_to_dense(rec, num_images, s_h, s_w, num_cls):
    """The reference's dense get_loss inputs from compact records: what collate_fn builds per ground-truth box
    (models/yolov2.py:1484-1505, 1537-1555; models/yolov1.py:1284-1312): all-zero [S,S,.] grids with the box's
    scalars scattered at its cell, obj_mask float64, x_img_id = arange(N), bbox_img_id = owning image."""
    m = len(rec)
    j, cy, cx = np.arange(m), rec["cy"], rec["cx"]
    sig_txty = np.zeros((m, s_h, s_w, 2), np.float32)
    twth = np.zeros((m, s_h, s_w, 2), np.float32)
    coord = np.zeros((m, s_h, s_w, 4), np.float32)
    cls_tgt = np.zeros((m, s_h, s_w, num_cls), np.float32)
    obj = np.zeros((m, s_h, s_w), np.float64)
    sig_txty[j, cy, cx, 0], sig_txty[j, cy, cx, 1] = rec["stx"], rec["sty"]
    twth[j, cy, cx, 0], twth[j, cy, cx, 1] = rec["tw"], rec["th"]
    for c, k in enumerate(("x1", "y1", "x2", "y2")):
        coord[j, cy, cx, c] = rec[k]
    cls_tgt[j, cy, cx, rec["cls"]] = 1.0
    obj[j, cy, cx] = 1.0
    x_img_id = np.arange(num_images, dtype=np.int64)
    return tuple(torch.from_numpy(a) for a in (sig_txty, twth, coord, cls_tgt, obj, x_img_id, x_img_id[rec["img"]].copy()))


# --------------------------------------------------------------------------------------
# compact tier (numpy closed form, per GT record)
# --------------------------------------------------------------------------------------
def _sigmoid32(x):
    x = x.astype(F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def decode_boxes_np(y, height, width, version, anchors):
    """fp32 op-by-op box decode; returns dict of [N,Sh,Sw,A] arrays + bbox [N,Sh,Sw,A,4].

    Same operation order as decode_torch (models/yolov2.py:488-595, models/yolov1.py:275-382);
    exp/sigmoid come from numpy's libm so they may differ from torch by an ulp.
    """
    y = np.asarray(y, dtype=F32)
    if version == 2:
        n, sh, sw, a, _ = y.shape
        box = y[..., :5]
        pw = np.array([p[0] for p in anchors], dtype=F32)
        ph = np.array([p[1] for p in anchors], dtype=F32)
    else:
        n, sh, sw, _ = y.shape
        a = int(anchors)
        box = y[..., : 5 * a].reshape(n, sh, sw, a, 5)
    sx = _sigmoid32(box[..., 0])
    sy = _sigmoid32(box[..., 1])
    if version == 2:
        aw = np.exp(box[..., 2], dtype=F32)
        ah = np.exp(box[..., 3], dtype=F32)
        bw = (pw * aw).astype(F32)
        bh = (ph * ah).astype(F32)
    else:
        aw = _sigmoid32(box[..., 2])
        ah = _sigmoid32(box[..., 3])
        bw = (F32(sw) * aw).astype(F32)
        bh = (F32(sh) * ah).astype(F32)
    cx = np.arange(sw, dtype=F32)[None, None, :, None]
    cy = np.arange(sh, dtype=F32)[None, :, None, None]
    bx = (sx + cx).astype(F32)
    by = (sy + cy).astype(F32)
    gw = F32(width / sw)
    gh = F32(height / sh)
    hw = (bw / F32(2)).astype(F32)
    hh = (bh / F32(2)).astype(F32)
    bbox = np.stack([(bx - hw) * gw, (by - hh) * gh, (bx + hw) * gw, (by + hh) * gh], -1).astype(F32)
    conf = _sigmoid32(box[..., 4])
    return dict(sx=sx, sy=sy, aw=aw, ah=ah, bbox=bbox, conf=conf)


def iou_np(b1, b2):
    """fp32 IoU, same rounding sequence as models/utils.py:47-63."""
    b1 = np.asarray(b1, dtype=F32)
    b2 = np.asarray(b2, dtype=F32)
    iw = np.maximum(np.minimum(b1[..., 2], b2[..., 2]) - np.maximum(b1[..., 0], b2[..., 0]), F32(0))
    ih = np.maximum(np.minimum(b1[..., 3], b2[..., 3]) - np.maximum(b1[..., 1], b2[..., 1]), F32(0))
    inter = (iw * ih).astype(F32)
    a1 = ((b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])).astype(F32)
    a2 = ((b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])).astype(F32)
    union = ((a1 + a2).astype(F32) - inter).astype(F32)
    return (inter / (union + F32(1e-6)).astype(F32)).astype(F32)


def train_head_compact(case, lambdas, m_global=None):
    """Closed-form loss / gradient per GT record, numpy.

    Decisions (IoU, argmax) in fp32 op-by-op; sums and gradients in float64 and cast at the
    end, so this tier is the more *accurate* statement of the same math; the tests allow it
    the north-star tolerance (1e-5 relative) against the dense tier and the CUDA path.
    """
    v = case.version
    y = case.y.numpy()
    n, sh, sw, a, c = case.n, case.s_h, case.s_w, case.a, case.c
    rec = case.rec
    m = len(rec)
    mg = m if m_global is None else int(m_global)
    anchors = case.anchors if v == 2 else a
    d = decode_boxes_np(y, case.height, case.width, v, anchors)
    lam = [float(lambdas[k]) for k in ("lambda_xy", "lambda_wh", "lambda_conf", "lambda_noobj", "lambda_cls")]
    preds = sh * sw * a
    k_n = np.diff(case.gt_off).astype(np.float64)

    img, cy, cx = rec["img"], rec["cy"], rec["cx"]
    gt_box = np.stack([rec["x1"], rec["y1"], rec["x2"], rec["y2"]], -1).astype(F32)
    pred_box = d["bbox"][img, cy, cx]                       # [M,A,4]
    iou = iou_np(pred_box, gt_box[:, None, :])               # [M,A]
    r = iou.argmax(-1)                                       # first max, like torch.max
    j = np.arange(m)
    iou_r = iou[j, r]

    sx = d["sx"][img, cy, cx, r].astype(np.float64)
    sy = d["sy"][img, cy, cx, r].astype(np.float64)
    aw = d["aw"][img, cy, cx, r]
    ah = d["ah"][img, cy, cx, r]
    conf_all = d["conf"].astype(np.float64)
    conf_r = conf_all[img, cy, cx, r]
    if v == 2:
        pw = np.array([p[0] for p in case.anchors], dtype=F32)[r]
        ph = np.array([p[1] for p in case.anchors], dtype=F32)[r]
        tw = np.sqrt((rec["tw"] / pw).astype(F32)).astype(np.float64)
        th = np.sqrt((rec["th"] / ph).astype(F32)).astype(np.float64)
    else:
        tw = np.sqrt(rec["tw"]).astype(np.float64)
        th = np.sqrt(rec["th"]).astype(np.float64)
    sqw = np.sqrt(aw).astype(np.float64)
    sqh = np.sqrt(ah).astype(np.float64)

    if v == 2:
        logits = y[img, cy, cx, r, 5:].astype(np.float64)    # [M,C]
    else:
        logits = y[img, cy, cx, 5 * a:].astype(np.float64)
    logits = logits - logits.max(-1, keepdims=True)
    p = np.exp(logits)
    p /= p.sum(-1, keepdims=True)
    onehot = np.zeros((m, c))
    onehot[j, rec["cls"]] = 1.0
    g = p - onehot

    s_xy = ((sx - rec["stx"]) ** 2 + (sy - rec["sty"]) ** 2).sum()
    s_wh = ((sqw - tw) ** 2 + (sqh - th) ** 2).sum()
    s_conf = ((iou_r.astype(np.float64) - conf_r) ** 2).sum()
    s_noobj = (k_n * (conf_all ** 2).reshape(n, -1).sum(-1)).sum() - (conf_r ** 2).sum()
    s_cls = (g ** 2).sum()
    den = np.array([2 * mg, 2 * mg, mg, mg * (preds - 1), mg], dtype=np.float64)
    terms = np.array([s_xy, s_wh, s_conf, s_noobj, s_cls]) / den
    loss = float((np.array(lam) * terms).sum())

    dy = np.zeros(y.shape, dtype=np.float64)
    if v == 2:
        dyb = dy                                             # [N,Sh,Sw,A,D]
    else:
        dyb = dy[..., : 5 * a].reshape(n, sh, sw, a, 5)      # view into dy
    # dense no-object part on the `to` channel
    dsig = conf_all * (1.0 - conf_all)
    dyb[..., 4] += lam[3] * 2.0 * conf_all * k_n[:, None, None, None] / den[3] * dsig
    # per-record sparse part (np.add.at: records may collide on a predictor)
    idx = (img, cy, cx, r)
    np.add.at(dyb[..., 0], idx, lam[0] * (sx - rec["stx"]) / mg * sx * (1 - sx))
    np.add.at(dyb[..., 1], idx, lam[0] * (sy - rec["sty"]) / mg * sy * (1 - sy))
    if v == 2:
        np.add.at(dyb[..., 2], idx, lam[1] * (sqw - tw) * sqw / (2.0 * mg))
        np.add.at(dyb[..., 3], idx, lam[1] * (sqh - th) * sqh / (2.0 * mg))
    else:
        awd, ahd = aw.astype(np.float64), ah.astype(np.float64)
        np.add.at(dyb[..., 2], idx, lam[1] * (sqw - tw) * sqw * (1 - awd) / (2.0 * mg))
        np.add.at(dyb[..., 3], idx, lam[1] * (sqh - th) * sqh * (1 - ahd) / (2.0 * mg))
    dconf = conf_r * (1 - conf_r)
    np.add.at(dyb[..., 4], idx,
              (lam[2] * 2.0 * (conf_r - iou_r) / mg - lam[3] * 2.0 * conf_r / den[3]) * dconf)
    gcls = lam[4] * (2.0 / mg) * p * (g - (g * p).sum(-1, keepdims=True))   # [M,C]
    if v == 2:
        for cc in range(c):
            np.add.at(dy[..., 5 + cc], idx, gcls[:, cc])
    else:
        for cc in range(c):
            np.add.at(dy[..., 5 * a + cc], (img, cy, cx), gcls[:, cc])
    return dict(loss=loss, terms=terms, dy=dy.astype(F32), resp=r.astype(np.int32),
                iou_resp=iou_r.astype(F32), iou_all=iou)


# --------------------------------------------------------------------------------------
# NMS
# --------------------------------------------------------------------------------------
def nms_image_np(bbox, conf, conf_thre, iou_thre, labels=None):
    """Greedy NMS of ONE image; returns kept predictor indices in descending-confidence order.

    models/utils.py:92-158: keep `conf >= conf_thre`, sort descending, a later box survives an
    earlier kept box iff `iou < iou_thre`; class-agnostic unless `labels` is given (the
    class-aware extra, in which only boxes with equal label suppress each other).
    Ties in confidence are ordered by ascending index (torch's sort leaves them unspecified).
    """
    bbox = np.asarray(bbox, dtype=F32).reshape(-1, 4)
    conf = np.asarray(conf, dtype=F32).reshape(-1)
    cand = np.nonzero(conf >= F32(conf_thre))[0]
    order = cand[np.argsort(-conf[cand], kind="stable")]
    alive = np.ones(len(order), dtype=bool)
    keep = []
    thr = F32(iou_thre)
    for i in range(len(order)):
        if not alive[i]:
            continue
        keep.append(int(order[i]))
        rest = order[i + 1:]
        if len(rest) == 0:
            break
        iou = iou_np(bbox[order[i]][None, :], bbox[rest])
        sup = ~(iou < thr)  # (utils.py:133 keeps j iff iou < thre: a NaN IoU removes it)
        if labels is not None:
            sup &= (labels[rest] == labels[order[i]])
        alive[i + 1:] &= ~sup
    return np.asarray(keep, dtype=np.int32)


def postprocess_np(y, height, width, version, anchors, conf_thre, iou_thre, class_aware=False):
    """Per-image decode + threshold + NMS + class pick (detect, models/yolov2.py:694-743).

    Returns a list (one entry per image) of dicts: idx, bbox, conf, label, score.
    Uses torch for sigmoid/softmax so values match the reference's CPU path bit for bit.
    """
    yt = torch.as_tensor(y)
    _, _, bbox, conf, _, spec = decode_torch(yt, height, width, version, anchors)
    n = yt.shape[0]
    c = spec.shape[-1]
    bbox = bbox.reshape(n, -1, 4).numpy()
    conf = conf.reshape(n, -1).numpy()
    spec = spec.reshape(n, -1, c).numpy()
    out = []
    for i in range(n):
        labels_all = spec[i].argmax(-1)
        keep = nms_image_np(bbox[i], conf[i], conf_thre, iou_thre,
                            labels=labels_all if class_aware else None)
        out.append(dict(idx=keep, bbox=bbox[i][keep], conf=conf[i][keep],
                        cls_spec=spec[i][keep],
                        label=labels_all[keep].astype(np.int32), score=spec[i][keep].max(-1) if len(keep) else np.zeros(0, F32)))
    return out


# --------------------------------------------------------------------------------------
# torch-CPU port of the reference's own NMS loop (cost model for the CPU baseline)
# --------------------------------------------------------------------------------------
def nms_torch(bbox, conf, cls_spec, conf_thre=0.9, iou_thre=0.5):
    """Shrinking-list greedy NMS the way the reference executes it on torch CPU.

    models/utils.py:89-164: threshold with `>=`, flatten, `torch.sort(descending=True)`, then
    for position i: IoU of row i against all later rows, keep the first i+1 rows and every
    later row with `iou < iou_thre`, re-gather all three tensors, advance.  Same tensor ops per
    iteration as the reference, so its wall-clock is representative of the reference's cost;
    bench.py times this as the CPU baseline ("port").  Returns the three gathered tensors in
    descending-confidence order, like the reference.
    """
    num_cls = cls_spec.shape[-1]
    sel = (conf >= conf_thre).reshape(-1)
    bbox = bbox.reshape(-1, 4)[sel]
    conf = conf.reshape(-1)[sel]
    cls_spec = cls_spec.reshape(-1, num_cls)[sel]
    conf, order = torch.sort(conf, descending=True)
    bbox = bbox[order]
    cls_spec = cls_spec[order]
    i = 0
    while i < conf.numel() - 1:
        survive = iou_torch(bbox[i:i + 1], bbox[i + 1:]) < iou_thre
        keep = torch.cat([torch.ones(i + 1, dtype=torch.bool), survive])
        bbox, conf, cls_spec = bbox[keep], conf[keep], cls_spec[keep]
        i += 1
    return bbox, conf, cls_spec


def postprocess_torch(y, height, width, version, anchors, conf_thre, iou_thre):
    """detect() for every image of a batch on torch CPU: decode once, then per image
    nms_torch + class pick (models/yolov2.py:694-743).  The flat predictor index rides along as
    an extra class column (exact in fp32 below 2^24) so that kept indices can be read back."""
    _, _, bbox, conf, _, spec = decode_torch(torch.as_tensor(y), height, width, version, anchors)
    n = bbox.shape[0]
    c = spec.shape[-1]
    out = []
    for i in range(n):
        s = spec[i].reshape(-1, c)
        idx = torch.arange(s.shape[0], dtype=torch.float32)[:, None]
        kb, kc, ks = nms_torch(bbox[i], conf[i], torch.cat([s, idx], -1), conf_thre, iou_thre)
        sp = ks[:, :-1]
        out.append(dict(idx=ks[:, -1].numpy().astype(np.int32), bbox=kb.numpy(), conf=kc.numpy(),
                        cls_spec=sp.numpy(),
                        label=sp.argmax(-1).numpy().astype(np.int32) if len(kc) else np.zeros(0, np.int32),
                        score=sp.max(-1)[0].numpy() if len(kc) else np.zeros(0, F32)))
    return out


# --------------------------------------------------------------------------------------
# evaluation: true-positive matching -- models/utils.py:231-262
# --------------------------------------------------------------------------------------
def match_detections_np(det_bbox, det_label, gt_boxes, gt_labels, levels):
    """One image: tp[K,L] for K detections against the image's ground truth, float64 like
    get_iou(..., numpy=True): a detection is a true positive at a level iff some ground-truth box of
    its class reaches that IoU (models/utils.py:241-257)."""
    levels = np.asarray(levels, dtype=np.float64)
    det_bbox = np.asarray(det_bbox, dtype=np.float64).reshape(-1, 4)
    gt_boxes = np.asarray(gt_boxes, dtype=np.float64).reshape(-1, 4)
    tp = np.zeros((len(det_bbox), len(levels)), dtype=np.int64)
    for k, (b, c) in enumerate(zip(det_bbox, det_label)):
        g = gt_boxes[np.asarray(gt_labels) == c]
        ix1, iy1 = np.maximum(g[:, 0], b[0]), np.maximum(g[:, 1], b[1])
        ix2, iy2 = np.minimum(g[:, 2], b[2]), np.minimum(g[:, 3], b[3])
        inter = np.clip(ix2 - ix1, 0, None) * np.clip(iy2 - iy1, 0, None)
        union = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
        iou = inter / (union + 1e-6)
        fp = ((iou[:, None] < levels).astype(int).prod(0) >= 1).astype(int)
        tp[k] = 1 - fp
    return tp


# --------------------------------------------------------------------------------------
# optimizer step -- models/yolov2.py:1253-1272 (torch.optim.SGD, re-created every iteration)
# --------------------------------------------------------------------------------------
def sgd_step_np(params, grads, lr, momentum=0.9, weight_decay=5e-4, bufs=None):
    """One torch.optim.SGD step per tensor, float32 with one rounding per torch op:
    d = g + weight_decay * p; buf = d on an optimizer's first step (bufs is None: what the reference
    gets by building a new optimizer in every iteration, models/yolov2.py:1254-1268), else
    momentum * buf + d; p <- p - lr * buf.  Returns (new params, new bufs)."""
    f32 = np.float32
    new_p, new_b = [], []
    for i, (p, g) in enumerate(zip(params, grads)):
        p = np.asarray(p, dtype=f32)
        d = np.asarray(g, dtype=f32) + f32(weight_decay) * p if weight_decay != 0 else np.asarray(g, dtype=f32)
        if momentum != 0:
            if bufs is not None and bufs[i] is not None:
                d = f32(momentum) * np.asarray(bufs[i], dtype=f32) + d
            new_b.append(d.astype(f32))
        else:
            new_b.append(None)
        new_p.append((p - f32(lr) * d).astype(f32))
    return new_p, new_b
