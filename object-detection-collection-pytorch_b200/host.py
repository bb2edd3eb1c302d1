"""Host-buffer entry point of the head path: the call a host-side caller (the reference's
`run_one_epoch` with CPU tensors, or any non-torch binding of the C ABI) makes when its head
tensor and ground truth live in HOST memory.

    pipe = HostHeadPipeline(n, s_h, s_w, a, c, ...)
    t = pipe.submit(y_host, gt_host, gt_off_host)       # enqueue H2D -> kernels -> D2H
    res = pipe.result(t)                                 # loss, dL/dy, kept boxes in host memory

Each submit stages its inputs through pinned buffers, runs the fused step (train head + post-process,
libyolohead's yh_v2_train_post through the C ABI; YOLOv1: the two separate kernels) and brings the loss,
its terms, the detections and -- on request -- dL/dy back.  Four streams (two H2D, kernels, D2H) and `depth`
slots let consecutive steps overlap on the two PCIe directions; nothing is computed on the CPU.  The head tensor
goes up in two halves on two streams: a copy carries ~45 us of set-up and completion latency that queued copies
on ONE stream do not overlap (measured on the B200 boxes, scratch/h2d_probe.py: 21.7 MB back to back 48.6 GB/s
on one stream, 53.9 GB/s as two halves on two streams -- the link's rate for large copies is 54.5 GB/s).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _write_combined_pinned(nbytes):
    """`nbytes` of write-combined pinned host memory as a uint8 tensor (cudaHostAlloc, WRITE_COMBINED | PORTABLE).
    The device reads such memory over PCIe without snooping the CPU caches -- what helps when several GPUs pull
    from one host at once -- and the CPU must only ever WRITE to it (reads are uncached).  Freed at process exit."""
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    ptr = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(nbytes), C.c_uint(0x04 | 0x01))
    if rc != 0 or not ptr.value:
        raise RuntimeError("cudaHostAlloc(write-combined, %d bytes) failed: %d" % (nbytes, rc))
    buf = (C.c_uint8 * nbytes).from_address(ptr.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


class HostHeadPipeline:
    def __init__(self, n, s_h, s_w, a, c, *, version=2, img_hw, anchors=None, lambdas, conf_thre=0.5,
                 iou_thre=0.45, max_out=128, max_boxes=None, depth=3, device=None, return_dy=True,
                 compute_dy=True, class_aware=False, write_combined=None):
        """`compute_dy`: the train head also produces dL/dy (the backward of get_loss).  `return_dy`: that
        gradient is copied back to host memory as well; with return_dy=False it stays in the slot's device
        buffer (`device_dy(ticket)`), which is where a caller whose backbone runs on the GPU consumes it --
        the step's host-side result is then the loss, the five terms and the detections."""
        if not torch.cuda.is_available():
            raise RuntimeError("HostHeadPipeline needs a CUDA device (no CPU path)")
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.version, self.n, self.a, self.c = version, n, a, c
        self.img_hw, self.anchors, self.lambdas = img_hw, anchors, lambdas
        self.conf_thre, self.iou_thre, self.max_out = conf_thre, iou_thre, max_out
        self.class_aware = class_aware
        self.compute_dy = bool(compute_dy or return_dy)
        # the head tensor's staging buffers in write-combined pinned memory (opt-in: YH_PINNED_WC=1 or the argument)
        import os
        self.write_combined = bool(int(os.environ.get("YH_PINNED_WC", "0"))) if write_combined is None else bool(write_combined)
        self.return_dy = return_dy
        self.depth = depth
        self.max_boxes = int(max_boxes if max_boxes is not None else 128 * n)
        shape = (n, s_h, s_w, a, 5 + c) if version == 2 else (n, s_h, s_w, 5 * a + c)
        self.shape = shape
        d, f32, i32 = self.dev, torch.float32, torch.int32
        self.s_h2d, self.s_h2d2, self.s_run, self.s_d2h = (torch.cuda.Stream(d) for _ in range(4))
        self.slots = []
        # Small inputs (records + offsets) and small outputs (terms, loss, detections) are packed into
        # one buffer each, so a step is two host-to-device and two device-to-host copies: every
        # separate copy costs ~10 us of launch/DMA set-up, which adds up against a 0.4 ms step.
        def carve(buf, spec):
            views, o = {}, 0
            for name, shp, dt in spec:
                o = (o + 15) & ~15
                nb = int(np.prod(shp)) * torch.empty((), dtype=dt).element_size()
                views[name] = buf[o:o + nb].view(dt).view(*shp) if shp else buf[o:o + nb].view(dt).view(())
                o += nb
            return views

        def layout(spec):
            o = 0
            for _, shp, dt in spec:
                o = ((o + 15) & ~15) + int(np.prod(shp)) * torch.empty((), dtype=dt).element_size()
            return (o + 15) & ~15

        in_spec = [("gt", (self.max_boxes, 12), i32), ("off", (n + 1,), i32)]
        out_spec = [("terms", (5,), f32), ("loss", (), f32), ("keep_cnt", (n,), i32), ("keep_idx", (n, max_out), i32),
                    ("bbox", (n, max_out, 4), f32), ("conf", (n, max_out), f32), ("label", (n, max_out), i32),
                    ("score", (n, max_out), f32)]
        self._in_bytes, self._out_bytes = layout(in_spec), layout(out_spec)
        for _ in range(depth):
            d_in = torch.empty(self._in_bytes, dtype=torch.uint8, device=d)
            d_out = torch.empty(self._out_bytes, dtype=torch.uint8, device=d)
            h_in = torch.empty(self._in_bytes, dtype=torch.uint8).pin_memory()
            h_out = torch.empty(self._out_bytes, dtype=torch.uint8).pin_memory()
            di, do, hi, ho = carve(d_in, in_spec), carve(d_out, out_spec), carve(h_in, in_spec), carve(h_out, out_spec)
            nbytes = int(ops._lib.load().yh_postprocess_workspace_bytes(n, s_h * s_w * a))
            post = dict(keep_idx=do["keep_idx"], keep_cnt=do["keep_cnt"], bbox=do["bbox"], conf=do["conf"], cls_spec=None,
                        label=do["label"], score=do["score"],
                        _ws=torch.empty(max(nbytes, 16), dtype=torch.uint8, device=d))
            s = dict(
                y=torch.empty(shape, dtype=f32, device=d), dy=torch.empty(shape, dtype=f32, device=d),
                d_in=d_in, d_out=d_out, h_in=h_in, h_out=h_out,
                gt=di["gt"], off=di["off"], loss=do["loss"], terms=do["terms"],
                h_y=(_write_combined_pinned(4 * int(np.prod(shape))).view(f32).view(shape) if self.write_combined
                     else torch.empty(shape, dtype=f32).pin_memory()),
                h_gt=hi["gt"], h_off=hi["off"],
                h_dy=torch.empty(shape, dtype=f32).pin_memory() if return_dy else None,
                h=ho,
                ev_in=torch.cuda.Event(), ev_in2=torch.cuda.Event(), ev_run=torch.cuda.Event(), ev_post=torch.cuda.Event(), ev_out=torch.cuda.Event(),
                busy=False, post=post,
            )
            s["res"] = dict(train=dict(dy=s["dy"] if self.compute_dy else None, loss=s["loss"], terms=s["terms"]),
                            post={k: v for k, v in post.items() if k != "_ws"}, _ws=None)
            self.slots.append(s)
        self._ticket = 0

    # bytes that cross PCIe per step (for reporting)
    def h2d_bytes(self, m):
        nb = (m * 48 + 15) & ~15
        small = self._in_bytes if nb + (self.n + 1) * 4 + 64 >= self._in_bytes // 2 else m * 48 + (self.n + 1) * 4
        return int(np.prod(self.shape)) * 4 + small

    def d2h_bytes(self):
        return (int(np.prod(self.shape)) * 4 if self.return_dy else 0) + self._out_bytes

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
        yd, yh = s["y"].view(-1), s["h_y"].view(-1)
        half = (yd.numel() // 2) & ~63
        with torch.cuda.stream(self.s_h2d2):
            yd[half:].copy_(yh[half:], non_blocking=True)
            s["ev_in2"].record(self.s_h2d2)
        with torch.cuda.stream(self.s_h2d):
            yd[:half].copy_(yh[:half], non_blocking=True)
            nb = (m * 48 + 15) & ~15  # the records in use; the offsets sit behind the full record area
            if nb + (self.n + 1) * 4 + 64 >= self._in_bytes // 2:
                s["d_in"].copy_(s["h_in"], non_blocking=True)
            else:
                s["gt"][:m].copy_(s["h_gt"][:m], non_blocking=True)
                s["off"].copy_(s["h_off"], non_blocking=True)
            s["ev_in"].record(self.s_h2d)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(s["ev_in"])
            self.s_run.wait_event(s["ev_in2"])
            if self.version == 2:
                # the fused step: train head + post-process, the head tensor read once (yh_v2_train_post)
                s["res"] = ops.train_post(s["y"], s["gt"][:m], s["off"], img_hw=self.img_hw, lambdas=self.lambdas,
                                          anchors=self.anchors, conf_thre=self.conf_thre, iou_thre=self.iou_thre,
                                          m_global=m_global, want_grad=self.compute_dy, class_aware=self.class_aware,
                                          max_out=self.max_out, want_cls_spec=False, out=s["res"])
            else:
                ops.train_head(s["y"], s["gt"][:m], s["off"], version=self.version, img_hw=self.img_hw,
                               lambdas=self.lambdas, anchors=self.anchors, boxes_per_cell=self.a,
                               m_global=m_global, want_grad=self.compute_dy,
                               out=dict(dy=s["dy"], loss=s["loss"], terms=s["terms"]))
                s["post"] = ops.postprocess(s["y"], version=self.version, img_hw=self.img_hw,
                                            conf_thre=self.conf_thre, iou_thre=self.iou_thre,
                                            anchors=self.anchors, boxes_per_cell=self.a,
                                            class_aware=self.class_aware, max_out=self.max_out,
                                            want_cls_spec=False, out=s["post"])
            s["ev_run"].record(self.s_run)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(s["ev_run"])
            if self.return_dy:
                s["h_dy"].copy_(s["dy"], non_blocking=True)
            s["h_out"].copy_(s["d_out"], non_blocking=True)  # terms, loss and the detections in one copy
            s["ev_out"].record(self.s_d2h)
        s["busy"] = True
        self._ticket += 1
        return t

    def pinned_inputs(self, ticket=None):
        """The pinned staging buffers of the slot the next submit (or `ticket`) uses: a producer
        can write the head tensor there directly and pass staged=True."""
        s = self.slots[(self._ticket if ticket is None else ticket) % self.depth]
        return s["h_y"], s["h_gt"], s["h_off"]

    def device_dy(self, ticket):
        """The device tensor holding dL/dy of step `ticket` (valid until the slot is reused)."""
        return self.slots[ticket % self.depth]["dy"] if self.compute_dy else None

    def result(self, ticket):
        """Block until step `ticket` is back in host memory; returns views of the pinned result
        buffers (valid until the slot is reused, `depth` submits later)."""
        s = self.slots[ticket % self.depth]
        s["ev_out"].synchronize()
        s["busy"] = False
        h = s["h"]
        return dict(loss=h["loss"], terms=h["terms"], dy=s["h_dy"], keep_cnt=h["keep_cnt"], keep_idx=h["keep_idx"],
                    bbox=h["bbox"], conf=h["conf"], label=h["label"], score=h["score"])
