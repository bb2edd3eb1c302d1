"""Host-buffer entry point of the head path: the call a host-side caller (the reference's
`run_one_epoch` with CPU tensors, or any non-torch binding of the C ABI) makes when its head
tensor and ground truth live in HOST memory.

    pipe = HostHeadPipeline(n, s_h, s_w, a, c, ...)
    t = pipe.submit(y_host, gt_host, gt_off_host)       # enqueue H2D -> kernels -> D2H
    res = pipe.result(t)                                 # loss, dL/dy, kept boxes in host memory

Each submit stages its inputs through pinned buffers, runs the fused train head and the
post-process kernels (libyolohead, C ABI) and brings loss, dL/dy and the detections back.  Three
streams (H2D, compute, D2H) and `depth` slots let consecutive steps overlap on the two PCIe
directions; nothing is computed on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class HostHeadPipeline:
    def __init__(self, n, s_h, s_w, a, c, *, version=2, img_hw, anchors=None, lambdas, conf_thre=0.5,
                 iou_thre=0.45, max_out=128, max_boxes=None, depth=3, device=None, return_dy=True,
                 class_aware=False):
        if not torch.cuda.is_available():
            raise RuntimeError("HostHeadPipeline needs a CUDA device (no CPU path)")
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.version, self.n, self.a, self.c = version, n, a, c
        self.img_hw, self.anchors, self.lambdas = img_hw, anchors, lambdas
        self.conf_thre, self.iou_thre, self.max_out = conf_thre, iou_thre, max_out
        self.class_aware = class_aware
        self.return_dy = return_dy
        self.depth = depth
        self.max_boxes = int(max_boxes if max_boxes is not None else 128 * n)
        shape = (n, s_h, s_w, a, 5 + c) if version == 2 else (n, s_h, s_w, 5 * a + c)
        self.shape = shape
        d, f32, i32 = self.dev, torch.float32, torch.int32
        self.s_h2d, self.s_run, self.s_d2h = (torch.cuda.Stream(d) for _ in range(3))
        self.slots = []
        for _ in range(depth):
            s = dict(
                y=torch.empty(shape, dtype=f32, device=d), dy=torch.empty(shape, dtype=f32, device=d),
                gt=torch.empty(self.max_boxes, 12, dtype=i32, device=d), off=torch.empty(n + 1, dtype=i32, device=d),
                loss=torch.empty((), dtype=f32, device=d), terms=torch.empty(5, dtype=f32, device=d),
                h_y=torch.empty(shape, dtype=f32).pin_memory(),
                h_gt=torch.empty(self.max_boxes, 12, dtype=i32).pin_memory(),
                h_off=torch.empty(n + 1, dtype=i32).pin_memory(),
                h_dy=torch.empty(shape, dtype=f32).pin_memory() if return_dy else None,
                h_scal=torch.empty(6, dtype=f32).pin_memory(),
                h_cnt=torch.empty(n, dtype=i32).pin_memory(),
                h_idx=torch.empty(n, max_out, dtype=i32).pin_memory(),
                h_box=torch.empty(n, max_out, 4, dtype=f32).pin_memory(),
                h_conf=torch.empty(n, max_out, dtype=f32).pin_memory(),
                h_label=torch.empty(n, max_out, dtype=i32).pin_memory(),
                h_score=torch.empty(n, max_out, dtype=f32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_run=torch.cuda.Event(), ev_out=torch.cuda.Event(),
                busy=False, post=None,
            )
            self.slots.append(s)
        self._ticket = 0

    # bytes that cross PCIe per step (for reporting)
    def h2d_bytes(self, m):
        return int(np.prod(self.shape)) * 4 + m * 48 + (self.n + 1) * 4

    def d2h_bytes(self):
        per = self.max_out * (4 + 16 + 4 + 4 + 4) + 4
        return (int(np.prod(self.shape)) * 4 if self.return_dy else 0) + 6 * 4 + self.n * per

    def submit(self, y_host, gt_host, gt_off_host, m_global=None, staged=False):
        """Enqueue one step.  y_host fp32 [shape]; gt_host int32 [M,12] sorted by image; gt_off_host
        int32 [N+1].  With staged=True the arguments are already this slot's pinned buffers
        (see `pinned_inputs`).  Returns a ticket for result()."""
        t = self._ticket
        s = self.slots[t % self.depth]
        if s["busy"]:
            raise RuntimeError("slot still in flight: call result(%d) first" % (t - self.depth))
        m = int(gt_host.shape[0])
        if m > self.max_boxes:
            raise ValueError("more boxes (%d) than max_boxes (%d)" % (m, self.max_boxes))
        if not staged:
            s["h_y"].copy_(y_host)
            s["h_gt"][:m].copy_(gt_host)
            s["h_off"].copy_(gt_off_host)
        with torch.cuda.stream(self.s_h2d):
            s["y"].copy_(s["h_y"], non_blocking=True)
            s["gt"][:m].copy_(s["h_gt"][:m], non_blocking=True)
            s["off"].copy_(s["h_off"], non_blocking=True)
            s["ev_in"].record(self.s_h2d)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(s["ev_in"])
            ops.train_head(s["y"], s["gt"][:m], s["off"], version=self.version, img_hw=self.img_hw,
                           lambdas=self.lambdas, anchors=self.anchors, boxes_per_cell=self.a,
                           m_global=m_global, want_grad=self.return_dy,
                           out=dict(dy=s["dy"], loss=s["loss"], terms=s["terms"]))
            s["post"] = ops.postprocess(s["y"], version=self.version, img_hw=self.img_hw,
                                        conf_thre=self.conf_thre, iou_thre=self.iou_thre,
                                        anchors=self.anchors, boxes_per_cell=self.a,
                                        class_aware=self.class_aware, max_out=self.max_out,
                                        want_cls_spec=False, out=s["post"])
            s["ev_run"].record(self.s_run)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(s["ev_run"])
            p = s["post"]
            if self.return_dy:
                s["h_dy"].copy_(s["dy"], non_blocking=True)
            s["h_scal"][:5].copy_(s["terms"], non_blocking=True)
            s["h_scal"][5:].copy_(s["loss"].reshape(1), non_blocking=True)
            s["h_cnt"].copy_(p["keep_cnt"], non_blocking=True)
            s["h_idx"].copy_(p["keep_idx"], non_blocking=True)
            s["h_box"].copy_(p["bbox"], non_blocking=True)
            s["h_conf"].copy_(p["conf"], non_blocking=True)
            s["h_label"].copy_(p["label"], non_blocking=True)
            s["h_score"].copy_(p["score"], non_blocking=True)
            s["ev_out"].record(self.s_d2h)
        s["busy"] = True
        self._ticket += 1
        return t

    def pinned_inputs(self, ticket=None):
        """The pinned staging buffers of the slot the next submit (or `ticket`) uses: a producer
        can write the head tensor there directly and pass staged=True."""
        s = self.slots[(self._ticket if ticket is None else ticket) % self.depth]
        return s["h_y"], s["h_gt"], s["h_off"]

    def result(self, ticket):
        """Block until step `ticket` is back in host memory; returns views of the pinned result
        buffers (valid until the slot is reused, `depth` submits later)."""
        s = self.slots[ticket % self.depth]
        s["ev_out"].synchronize()
        s["busy"] = False
        return dict(loss=s["h_scal"][5], terms=s["h_scal"][:5], dy=s["h_dy"], keep_cnt=s["h_cnt"],
                    keep_idx=s["h_idx"], bbox=s["h_box"], conf=s["h_conf"], label=s["h_label"],
                    score=s["h_score"])
