"""Drop-in surface of the reference's detector classes for the head hot path.

    models.utils.get_iou / nms          <- reference models/utils.py:5-164
    models.yolov2.YOLOv2HeadOps         <- predict / get_loss / detect of reference models/yolov2.py
    models.yolov1.YOLOv1HeadOps         <- predict / get_loss / detect of reference models/yolov1.py
    models.patch_reference(...)         install the three methods on the reference's own classes

The conv backbone is not part of this package (it stays on stock cuDNN layers); the classes
here only need `self(x_batch)` to return the head tensor.
"""
from .patch import patch_reference  # noqa: F401
