"""YOLOv1 head path (reference models/yolov1.py:207-931) on the CUDA kernels."""
from __future__ import annotations

import numpy as np
import torch

from ._head import HeadOps, InjectedHead

V1_INPUT = 224  # the reference's YOLOv1 pipeline runs at 224x224 (models/yolov1.py:40-47)


class YOLOv1HeadOps(HeadOps):
    """predict / get_loss / detect of the reference's YOLOv1; forward must return the
    [N,S_h,S_w,5B+C] head tensor (reference models/yolov1.py:150-163)."""
    _yh_version = 1

    def detect(self, img, conf_score_thre=0.9, iou_thre=0.5):
        """Resize to 224x224 -> post-process -> clip to [0,223] -> rescale boxes to the original
        size (reference models/yolov1.py:439-554; the resizes there go through albumentations,
        here through cv2 with the same bilinear interpolation and a plain box rescale)."""
        import cv2
        self.eval()
        height, width = img.shape[:2]
        small = cv2.resize(np.asarray(img), (V1_INPUT, V1_INPUT), interpolation=cv2.INTER_LINEAR)
        with torch.no_grad():
            h = self._yh_detect_host(self._yh_image_batch(small), conf_score_thre, iou_thre)

        def to_original(b):
            # the reference clips in float32 (models/yolov1.py:517-523), hands python floats to an albumentations
            # Resize(height, width) with pascal_voc boxes (:536-543, :1357-1367), which normalises by the 224x224
            # image and scales to the original size in float64: (x / 224) * width
            b = np.clip(b, np.float32(0.0), np.float32(V1_INPUT - 1.0)).astype(np.float64)
            b[:, 0::2] = b[:, 0::2] / V1_INPUT * width
            b[:, 1::2] = b[:, 1::2] / V1_INPUT * height
            return b

        return self._yh_annot(h, 0, to_original)


class YOLOv1Head(YOLOv1HeadOps, InjectedHead):
    """Head-only YOLOv1 with the reference constructor's arguments (models/yolov1.py:51-68)."""

    def __init__(self, num_grid_cell_in_height=7, num_grid_cell_in_width=7, num_anchor_box=2,
                 cls_list=None, cls2idx=None, num_cls=None):
        InjectedHead.__init__(self)
        if cls_list is None:
            cls_list = [str(i) for i in range(int(num_cls))]
        self.num_grid_cell_in_height = num_grid_cell_in_height
        self.num_grid_cell_in_width = num_grid_cell_in_width
        self.num_anchor_box = num_anchor_box
        self.cls_list = list(cls_list)
        self.cls2idx = cls2idx if cls2idx is not None else {c: i for i, c in enumerate(self.cls_list)}
        self.num_cls = len(self.cls_list)
