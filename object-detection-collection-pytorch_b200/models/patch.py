"""Install the CUDA head path on the reference's own classes.

    import models.yolov1, models.yolov2, models.utils      # the reference, unmodified
    from odcp_b200.models import patch_reference
    patch_reference(models.yolov1, models.yolov2, models.utils)

After this, `train.py` constructs `YOLOv1/YOLOv2` exactly as before (same constructor, same
parameters and state_dict keys) and `run_one_epoch` / `evaluate_model` reach the kernels through
the unchanged `get_loss` / `detect` / `nms` / `get_iou` calls.
"""
from __future__ import annotations

_METHODS = ("predict", "get_loss", "get_loss_compact", "get_loss_from_boxes", "detect", "detect_batch", "postprocess",
            "_yh_anchors", "_yh_kwargs", "_yh_image_batch", "_yh_annot", "_yh_detect_host")


def _channels_last_init(cls):
    """Wrap `cls.__init__` so that every Conv2d of a new model is put into channels_last memory format (values and
    state_dict keys unchanged): the head conv then writes NHWC and the reference's `permute(0, 2, 3, 1)` + `reshape`
    in head() (models/yolov2.py:338-362) is a view instead of a copy of the head tensor -- and of its gradient."""
    from .layout import use_channels_last_head
    orig = cls.__init__
    if getattr(orig, "_yh_channels_last", False):
        return

    def __init__(self, *args, **kwargs):
        orig(self, *args, **kwargs)
        use_channels_last_head(self)

    __init__._yh_channels_last = True
    __init__.__doc__ = orig.__doc__
    cls.__init__ = __init__


def patch_reference(ref_yolov1=None, ref_yolov2=None, ref_utils=None, fused_sgd=False, channels_last=False):
    """`fused_sgd=True` also replaces the `SGD` name the reference's run_one_epoch resolves
    (models/yolov2.py:7, :1253-1268) with odcp_b200.optim.SGD: the same update (a new optimizer per
    iteration, so every step is a first step) in one kernel launch.
    `channels_last=True` makes YOLOv2 models constructed afterwards run their convolutions in channels_last, which
    turns the layout change in front of the head path into a view (models/layout.py)."""
    from . import utils as u
    from .yolov1 import YOLOv1HeadOps
    from .yolov2 import YOLOv2HeadOps
    done = []
    for mod, cls_name, ops_cls in ((ref_yolov1, "YOLOv1", YOLOv1HeadOps), (ref_yolov2, "YOLOv2", YOLOv2HeadOps)):
        if mod is None:
            continue
        cls = getattr(mod, cls_name)
        for name in _METHODS:
            for klass in ops_cls.__mro__:
                if name in klass.__dict__:
                    setattr(cls, name, klass.__dict__[name])
                    break
        cls._yh_version = ops_cls._yh_version
        # the module-level names the reference's methods resolve at call time
        mod.nms = u.nms
        mod.get_iou = u.get_iou
        if fused_sgd:
            from ..optim import SGD
            mod.SGD = SGD
        if channels_last and cls_name == "YOLOv2":  # (YOLOv1's head is a Linear layer: nothing to lay out)
            _channels_last_init(cls)
        done.append(cls_name)
    if ref_utils is not None:
        ref_utils.nms = u.nms
        ref_utils.get_iou = u.get_iou
        ref_utils.evaluate_model = u.evaluate_model  # batched drop-in (same arguments, same result dict)
        done.append("utils")
    return done
