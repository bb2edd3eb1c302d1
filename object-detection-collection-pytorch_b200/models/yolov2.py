"""YOLOv2 head path (reference models/yolov2.py:433-1140) on the CUDA kernels."""
from __future__ import annotations

from ..synthetic import YOLOV2_ANCHORS
from ._head import HeadOps, InjectedHead


class YOLOv2HeadOps(HeadOps):
    """predict / get_loss / detect of the reference's YOLOv2; mix into (or patch onto) any module
    whose forward returns the [N,S_h,S_w,A,5+C] head tensor (reference models/yolov2.py:338-362)."""
    _yh_version = 2


class YOLOv2Head(YOLOv2HeadOps, InjectedHead):
    """Head-only YOLOv2: same attributes as the reference constructor sets
    (models/yolov2.py:49-70) minus the backbone; no parameters or buffers."""

    def __init__(self, cls_list=None, cls2idx=None, num_cls=None, anchors=YOLOV2_ANCHORS):
        InjectedHead.__init__(self)
        if cls_list is None:
            cls_list = [str(i) for i in range(int(num_cls))]
        self.cls_list = list(cls_list)
        self.cls2idx = cls2idx if cls2idx is not None else {c: i for i, c in enumerate(self.cls_list)}
        self.num_cls = len(self.cls_list)
        self.anchor_box_size_list = [tuple(a) for a in anchors]
        self.num_anchor_box = len(self.anchor_box_size_list)
        self.head_output_dim = self.num_anchor_box * (5 + self.num_cls)
