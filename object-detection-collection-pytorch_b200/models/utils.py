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
