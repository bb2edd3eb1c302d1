"""Seeded synthetic VOC-shaped inputs for the YOLO head hot path.

There is no dataset offline, so tests and bench.py drive the path with synthetic head
tensors and ground-truth boxes whose *shape* follows the reference's data pipeline:
head tensor `[N,S_h,S_w,A,5+C]` (v2, reference models/yolov2.py:354-362) or
`[N,S_h,S_w,5B+C]` (v1, models/yolov1.py:150-163); boxes inside the image with the
target scalars derived exactly as `collate_fn` does (see targets.boxes_to_records).

Everything is generated on the CPU from explicit seeds so that the CUDA path and the CPU
oracle see identical bits.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .targets import boxes_to_records, csr_offsets

# YOLOv2 anchor priors in grid units, (w, h) -- reference models/yolov2.py:49-55
YOLOV2_ANCHORS = (
    (1.3221, 1.73145),
    (3.19275, 4.00944),
    (5.05587, 8.09892),
    (9.47112, 4.84053),
    (11.2364, 10.0071),
)

# default loss weights -- reference config.py:28-32, 47-51
DEFAULT_LAMBDAS = dict(lambda_xy=5.0, lambda_wh=5.0, lambda_conf=1.0, lambda_noobj=0.5, lambda_cls=1.0)


@dataclass
class HeadCase:
    """One synthetic workload: head tensor + compact GT."""
    name: str
    version: int          # 1 or 2
    n: int
    s_h: int
    s_w: int
    a: int                # anchors (v2) / boxes per cell (v1)
    c: int
    height: int
    width: int
    y: torch.Tensor       # fp32 CPU
    rec: np.ndarray       # GT_DTYPE, sorted by image
    gt_off: np.ndarray    # int32 [N+1]
    anchors: tuple = field(default=YOLOV2_ANCHORS)

    @property
    def m(self):
        return len(self.rec)

    @property
    def cell_floats(self):
        return self.a * (5 + self.c) if self.version == 2 else 5 * self.a + self.c

    @property
    def image_bytes(self):
        """P of SURVEY 8(d): bytes of one image's head tensor."""
        return self.s_h * self.s_w * self.cell_floats * 4


def make_boxes(rng, n, height, width, k_lo, k_hi, num_cls, wh_lo=0.05, wh_hi=0.6):
    """k_n ~ U{k_lo..k_hi} boxes per image, float64 xyxy pixels fully inside the image."""
    k = rng.integers(k_lo, k_hi + 1, size=n)
    m = int(k.sum())
    img = np.repeat(np.arange(n), k)
    w = rng.uniform(wh_lo, wh_hi, size=m) * width
    h = rng.uniform(wh_lo, wh_hi, size=m) * height
    cxp = rng.uniform(w / 2, width - w / 2)
    cyp = rng.uniform(h / 2, height - h / 2)
    boxes = np.stack([cxp - w / 2, cyp - h / 2, cxp + w / 2, cyp + h / 2], axis=1)
    boxes = np.clip(boxes, 0.0, [width, height, width, height])
    labels = rng.integers(0, num_cls, size=m)
    return boxes, labels, img


def make_case(name, version, n, s_h, s_w, a, c, height, width, seed, k_lo=1, k_hi=5,
              y_scale=1.0, to_shift=0.0, anchors=YOLOV2_ANCHORS):
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    if version == 2:
        y = torch.randn(n, s_h, s_w, a, 5 + c, generator=gen, dtype=torch.float32)
        if y_scale != 1.0:
            y = y * y_scale
        if to_shift != 0.0:
            y[..., 4] += to_shift
    else:
        y = torch.randn(n, s_h, s_w, 5 * a + c, generator=gen, dtype=torch.float32)
        if y_scale != 1.0:
            y = y * y_scale
        if to_shift != 0.0:
            y[..., 4:5 * a:5] += to_shift
    boxes, labels, img = make_boxes(rng, n, height, width, k_lo, k_hi, c)
    rec = boxes_to_records(boxes, labels, img, height, width, s_h, s_w, version)
    off = csr_offsets(rec, n)
    return HeadCase(name, version, n, s_h, s_w, a, c, height, width, y.contiguous(), rec, off,
                    anchors=tuple(anchors))


# The BASELINE.json configurations; `n` can be overridden for samples.
def cfg1(n=8):
    """YOLOv1 448x448, S=7, B=2, C=20, batch 8."""
    return make_case("cfg1_v1_7x7", 1, n, 7, 7, 2, 20, 448, 448, seed=101)


def cfg2(n=64, k_hi=5):
    """YOLOv2 416x416, 13x13, 5 anchors, C=20, batch 64."""
    return make_case("cfg2_v2_13x13", 2, n, 13, 13, 5, 20, 416, 416, seed=102, k_hi=k_hi)


def cfg3(n=256, conf_thre=0.5):
    """YOLOv2 inference post-process, batch 256, ~50 of 845 predictors above the threshold."""
    # P(sigmoid(N(mu,1)) >= thr) = 50/845  ->  mu = logit(thr) - 1.563
    mu = float(np.log(conf_thre / (1.0 - conf_thre)) - 1.563)
    return make_case("cfg3_v2_nms", 2, n, 13, 13, 5, 20, 416, 416, seed=103, to_shift=mu)


def cfg5(n=512):
    """YOLOv2 608x608, 19x19, batch 512, 50..100 GT boxes per image."""
    return make_case("cfg5_v2_19x19_dense", 2, n, 19, 19, 5, 20, 608, 608, seed=105,
                     k_lo=50, k_hi=100, to_shift=float(np.log(0.5 / 0.5) - 1.59))


def headline(n=256):
    """The configuration the north-star target is quoted on: YOLOv2 13x13x5, C=20, batch 256,
    VOC-like 1..5 boxes per image, `to` shifted so ~50 predictors/image pass conf 0.5."""
    mu = -1.563
    return make_case("v2_13x13x5_c20_b%d" % n, 2, n, 13, 13, 5, 20, 416, 416, seed=102,
                     to_shift=mu)


def distinct_scores(y, version, a):
    """True if all objectness logits within every image are distinct (NMS order is then unique)."""
    if version == 2:
        to = y[..., 4].reshape(y.shape[0], -1)
    else:
        to = y[..., 4:5 * a:5].reshape(y.shape[0], -1)
    s, _ = torch.sort(to, dim=1)
    return bool((s[:, 1:] != s[:, :-1]).all())


def with_collisions(case, extra, seed=0):
    """Copy of `case` with `extra` additional boxes that land in cells already occupied:
    every other one an exact duplicate with another class (same responsible predictor),
    the rest jittered by a pixel or two (same cell, possibly another predictor).
    The reference treats every box independently, so contributions must add up."""
    rng = np.random.default_rng(seed)
    rec = case.rec
    pick = rng.integers(0, len(rec), size=extra)
    boxes = np.stack([rec["x1"], rec["y1"], rec["x2"], rec["y2"]], 1).astype(np.float64)
    new_boxes = boxes[pick].copy()
    jitter = rng.uniform(-1.5, 1.5, size=new_boxes.shape)
    jitter[::2] = 0.0
    new_boxes = np.clip(new_boxes + jitter, 0.0, [case.width, case.height, case.width, case.height])
    new_lbl = (rec["cls"][pick] + 1 + rng.integers(0, case.c - 1, size=extra)) % case.c
    all_boxes = np.concatenate([boxes, new_boxes])
    all_lbl = np.concatenate([rec["cls"], new_lbl])
    all_img = np.concatenate([rec["img"], rec["img"][pick]])
    rec2 = boxes_to_records(all_boxes, all_lbl, all_img, case.height, case.width, case.s_h,
                            case.s_w, case.version)
    return HeadCase(case.name + "_coll", case.version, case.n, case.s_h, case.s_w, case.a, case.c,
                    case.height, case.width, case.y, rec2, csr_offsets(rec2, case.n),
                    anchors=case.anchors)
