"""Shared implementation of predict / get_loss / detect on the CUDA kernels."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

LAMBDA_KEYS = ops.LAMBDA_KEYS


class HeadLoss(torch.autograd.Function):
    """loss = fused train head(y, compact targets); backward hands out the dL/dy the same
    kernel already produced, scaled by the upstream gradient on the device (no host sync)."""

    @staticmethod
    def forward(ctx, y, gt, gt_off, cfg):
        want_grad = bool(ctx.needs_input_grad[0])
        kw = {k: cfg[k] for k in ("version", "img_hw", "lambdas", "anchors", "boxes_per_cell", "m_global")}
        r = ops.train_head(y.detach(), gt, gt_off, want_grad=want_grad, want_resp=cfg.get("want_resp", False), **kw)
        ctx.dy = r["dy"]
        ctx.kw = kw
        ctx.save_for_backward(y, gt, gt_off)  # only read if backward runs a second time (retain_graph)
        cfg["last"] = r
        ctx.mark_non_differentiable(r["terms"])
        return r["loss"], r["terms"]

    @staticmethod
    def backward(ctx, grad_loss, _grad_terms):
        dy, ctx.dy = ctx.dy, None  # the kernel's gradient buffer is handed out once: later backwards never touch it
        if dy is None:
            # a repeated backward (retain_graph=True): the first one gave the buffer away, possibly scaled -- run the
            # kernel again instead of un-scaling (exact, and an upstream gradient of 0 loses nothing); without
            # retain_graph reading the saved tensors raises torch's usual "backward through the graph a second time"
            y, gt, gt_off = ctx.saved_tensors
            dy = ops.train_head(y.detach(), gt, gt_off, want_grad=True, **ctx.kw)["dy"]
        g = grad_loss.to(device=dy.device, dtype=torch.float32).reshape(())
        return ops.scale_inplace(dy, g.contiguous()), None, None, None  # (a no-op on the device for g == 1)


class HeadOps:
    """Mixin: needs `self(x_batch)` -> head tensor, `self._yh_version`, `self.num_anchor_box`,
    `self.num_cls`, `self.cls_list` and (v2) `self.anchor_box_size_list`."""

    _yh_version = 2

    # -- helpers -----------------------------------------------------------------------------
    def _yh_anchors(self):
        return tuple(self.anchor_box_size_list) if self._yh_version == 2 else None

    def _yh_kwargs(self, x_batch):
        h, w = int(x_batch.shape[1]), int(x_batch.shape[2])
        return dict(version=self._yh_version, img_hw=(h, w), anchors=self._yh_anchors(),
                    boxes_per_cell=int(self.num_anchor_box))

    # -- predict -----------------------------------------------------------------------------
    def predict(self, x_batch):
        """The reference's six outputs (models/yolov2.py:433-649, models/yolov1.py:207-437):
        (sig_txty, exp_twth | sig_twth, bbox_xyxy_px, conf, cls_prob, cls_spec_conf).
        They carry no autograd graph: training goes through get_loss, whose backward is fused."""
        y = self(x_batch)
        return ops.decode(y.detach(), **self._yh_kwargs(x_batch))

    # -- loss --------------------------------------------------------------------------------
    def get_loss(self, x_batch, sig_txty_tgt_batch, wh_tgt_batch, bbox_coord_tgt_batch, cls_tgt_batch,
                 obj_mask_batch, x_img_id_batch, bbox_img_id_batch, lambda_xy, lambda_wh, lambda_conf,
                 lambda_noobj, lambda_cls):
        """Reference signature (models/yolov2.py:747-762, models/yolov1.py:556-571): dense per-box
        target grids in, 0-dim fp32 loss with a working .backward() out."""
        y = self(x_batch)
        dev = y.device
        gt, gt_off, status = ops.compact_targets(
            sig_txty_tgt_batch.to(dev), wh_tgt_batch.to(dev), bbox_coord_tgt_batch.to(dev),
            cls_tgt_batch.to(dev), obj_mask_batch.to(dev), x_img_id_batch, bbox_img_id_batch)
        self._yh_target_status = status  # device int32[2]: malformed obj_mask rows, non-one-hot class rows
        lam = dict(zip(LAMBDA_KEYS, (lambda_xy, lambda_wh, lambda_conf, lambda_noobj, lambda_cls)))
        return self.get_loss_compact(x_batch, gt, gt_off, y=y, **lam)

    def get_loss_compact(self, x_batch, gt, gt_off, *, y=None, m_global=None, want_resp=False,
                         lambda_xy=5.0, lambda_wh=5.0, lambda_conf=1.0, lambda_noobj=0.5, lambda_cls=1.0):
        """Fast path: compact ground-truth records (targets.py) instead of dense grids.
        `m_global` is the box count the means are taken over when the batch is sharded."""
        if y is None:
            y = self(x_batch)
        cfg = self._yh_kwargs(x_batch)
        cfg.update(lambdas=(lambda_xy, lambda_wh, lambda_conf, lambda_noobj, lambda_cls),
                   m_global=m_global, want_resp=want_resp)
        if not y.is_contiguous():
            y = y.contiguous()
        loss, terms = HeadLoss.apply(y, gt, gt_off, cfg)
        self._yh_last = dict(terms=terms, **{k: cfg["last"][k] for k in ("resp", "iou_resp")})
        return loss

    def get_loss_from_boxes(self, x_batch, boxes_xyxy, labels, img_index, *, m_global=None, lambda_xy=5.0,
                            lambda_wh=5.0, lambda_conf=1.0, lambda_noobj=0.5, lambda_cls=1.0):
        """Fast path from raw annotations: float64 pixel boxes [M,4], class indices [M] and the batch
        position of each box's image [M] (grouped by image, the order collate_fn walks them).  The
        targets are built on the device with collate_fn's float64 arithmetic (yh_build_targets), so
        neither the dense grids nor their M*5 host-to-device copies exist."""
        y = self(x_batch)
        dev = y.device
        s_h, s_w = int(y.shape[1]), int(y.shape[2])
        gt, gt_off, status = ops.build_targets(
            torch.as_tensor(boxes_xyxy, dtype=torch.float64).to(dev), torch.as_tensor(labels), torch.as_tensor(img_index),
            num_images=int(y.shape[0]), version=self._yh_version,
            img_hw=(int(x_batch.shape[1]), int(x_batch.shape[2])), grid=(s_h, s_w))
        self._yh_target_status = status  # device int32[2]: boxes out of image order, out-of-range boxes
        return self.get_loss_compact(x_batch, gt, gt_off, y=y, m_global=m_global, lambda_xy=lambda_xy,
                                     lambda_wh=lambda_wh, lambda_conf=lambda_conf, lambda_noobj=lambda_noobj,
                                     lambda_cls=lambda_cls)

    # -- detect ------------------------------------------------------------------------------
    def _yh_image_batch(self, img):
        dev = next(self.parameters(), torch.empty(0, device="cuda")).device
        if dev.type != "cuda":
            raise RuntimeError("the model must live on a CUDA device (no CPU path)")
        return torch.as_tensor(np.asarray([img]), device=dev)

    def postprocess(self, x_batch, conf_score_thre=0.9, iou_thre=0.5, class_aware=False, max_out=None):
        """Batched decode + threshold + per-image NMS + class pick, everything on the device."""
        y = self(x_batch)
        return ops.postprocess(y.detach(), conf_thre=conf_score_thre, iou_thre=iou_thre,
                               class_aware=class_aware, max_out=max_out, **self._yh_kwargs(x_batch))

    def _yh_detect_host(self, x_batch, conf_score_thre, iou_thre, class_aware=False, max_out=None):
        """Post-process of a batch with the results in HOST memory: the outputs the annotation dicts need
        (counts, boxes, confidences, labels, scores) live in one packed device buffer and come back with ONE
        device-to-host copy -- the only synchronisation of detect / detect_batch."""
        y = self(x_batch).detach()
        kw = self._yh_kwargs(x_batch)
        n, s_h, s_w, a, c = ops.head_shape(y, kw["version"], kw["boxes_per_cell"])
        p = s_h * s_w * a
        max_out = p if max_out is None else int(max_out)
        i32, f32 = torch.int32, torch.float32
        spec = [("keep_cnt", (n,), i32), ("keep_idx", (n, max_out), i32), ("bbox", (n, max_out, 4), f32),
                ("conf", (n, max_out), f32), ("label", (n, max_out), i32), ("score", (n, max_out), f32)]
        offs, o = {}, 0
        for name, shp, dt in spec:
            o = (o + 15) & ~15
            offs[name] = (o, int(np.prod(shp)) * 4, shp, dt)
            o += offs[name][1]

        def carve(buf):
            return {k: buf[lo:lo + nb].view(dt).view(*shp) for k, (lo, nb, shp, dt) in offs.items()}

        with torch.cuda.device(y.device):
            buf = torch.empty((o + 15) & ~15, dtype=torch.uint8, device=y.device)
            ws = torch.empty(max(int(ops._lib.load().yh_postprocess_workspace_bytes(n, p)), 16), dtype=torch.uint8,
                             device=y.device)
            ops.postprocess(y, conf_thre=conf_score_thre, iou_thre=iou_thre, class_aware=class_aware, max_out=max_out,
                            want_cls_spec=False, out=dict(carve(buf), cls_spec=None, _ws=ws), **kw)
            host = buf.cpu()
        return {k: v.numpy() for k, v in carve(host).items()}

    def _yh_annot(self, h, n, box_fn=None):
        k = min(int(h["keep_cnt"][n]), h["keep_idx"].shape[1])
        bbox = h["bbox"][n, :k]
        if box_fn is not None:
            bbox = box_fn(bbox)
        return {
            "bbox_list": bbox.tolist(),
            "lbl_list": [self.cls_list[i] for i in h["label"][n, :k].tolist()],
            "conf_score_list": h["conf"][n, :k].tolist(),
            "cls_spec_conf_score_list": h["score"][n, :k].tolist(),
        }

    def detect(self, img, conf_score_thre=0.9, iou_thre=0.5):
        """One image -> dict of python lists (reference models/yolov2.py:651-745)."""
        self.eval()
        with torch.no_grad():
            h = self._yh_detect_host(self._yh_image_batch(img), conf_score_thre, iou_thre)
        return self._yh_annot(h, 0)

    def detect_batch(self, x_batch, conf_score_thre=0.9, iou_thre=0.5, class_aware=False):
        """Batched detect: one launch and one device-to-host copy for all images, list of per-image dicts."""
        self.eval()
        with torch.no_grad():
            h = self._yh_detect_host(x_batch, conf_score_thre, iou_thre, class_aware=class_aware)
        return [self._yh_annot(h, n) for n in range(x_batch.shape[0])]


class InjectedHead(torch.nn.Module):
    """A module whose forward returns a head tensor handed in beforehand: the conv backbone is
    out of scope of this package, tests and the bench drive the head path with synthetic head
    tensors through this class."""

    def __init__(self):
        super().__init__()
        self._y = None

    def set_head_output(self, y):
        self._y = y
        return self

    def forward(self, x_batch):
        if self._y is None:
            raise RuntimeError("set_head_output(y) first")
        return self._y
