"""Import the UNMODIFIED reference (TEST INFRASTRUCTURE: tests/, bench.py's reference legs, golden generators).

    ref = load_reference(device="cpu")         # or "cuda": the reference's global DEVICE string
    ref.yolov2.YOLOv2, ref.utils.nms, ref.head_only(case), ...

The reference is imported from `/root/reference` when it exists (the build container) and from the staged copy
`oracle/_ref/` otherwise (the GPU box; see oracle/stage_reference.py), with the two shims of SURVEY App. C and
no edits: an `albumentations` stub (not installed offline; the reference only builds module-level transform
objects with it) and a `config` module carrying the DEVICE string the reference reads at import time
(config.py:2).  `available()` says whether either location exists.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LOCATIONS = ("/root/reference", os.path.join(HERE, "_ref"))


def location():
    for loc in LOCATIONS:
        if os.path.exists(os.path.join(loc, "models", "yolov2.py")):
            return loc
    return None


def available():
    return location() is not None


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):  # (inspect / importlib probe modules for __file__, __path__, ...)
            raise AttributeError(name)
        return lambda *a, **k: None


class Reference:
    pass


_loaded = {}


def load_reference(device="cpu"):
    """Import (once per device string) config, models.utils, models.yolov1, models.yolov2 of the reference."""
    if device in _loaded:
        return _loaded[device]
    loc = location()
    if loc is None:
        raise RuntimeError("the reference is neither at /root/reference nor staged in oracle/_ref "
                           "(run `python oracle/stage_reference.py` in the build container)")
    for name in ("albumentations", "albumentations.pytorch"):
        sys.modules.setdefault(name, _Stub(name))
    # a fresh import per device string: the reference binds DEVICE into module globals at import time
    for name in [n for n in sys.modules if n == "config" or n == "models" or n.startswith("models.")]:
        del sys.modules[name]
    sys.path.insert(0, loc)
    try:
        cfg = importlib.import_module("config")
        cfg.DEVICE = device
        ref = Reference()
        ref.config = cfg
        ref.utils = importlib.import_module("models.utils")
        ref.yolov2 = importlib.import_module("models.yolov2")
        ref.yolov1 = importlib.import_module("models.yolov1")
        ref.location, ref.device = loc, device
    finally:
        sys.path.remove(loc)
    # keep the modules importable by name for the reference's own lazy imports, but drop them from sys.modules on
    # the next load_reference(other device)
    _add_head_only(ref)
    _loaded[device] = ref
    return ref


def _add_head_only(ref):
    """Head-only subclasses: the reference's predict / get_loss / detect on an injected head tensor (the conv
    backbone is out of scope and cannot even be constructed offline for YOLOv1: torch.hub, googlenet.py:12-14)."""
    import torch

    class HeadOnlyV2(ref.yolov2.YOLOv2):
        def __init__(self, num_cls, anchors):
            torch.nn.Module.__init__(self)
            self.anchor_box_size_list = [tuple(a) for a in anchors]
            self.num_anchor_box = len(self.anchor_box_size_list)
            self.anchor_box_width_list = torch.tensor([b[0] for b in self.anchor_box_size_list]).to(ref.device)
            self.anchor_box_height_list = torch.tensor([b[1] for b in self.anchor_box_size_list]).to(ref.device)
            self.cls_list = [str(i) for i in range(num_cls)]
            self.cls2idx = {c: i for i, c in enumerate(self.cls_list)}
            self.num_cls = num_cls

        def forward(self, x):
            return self._y

    class HeadOnlyV1(ref.yolov1.YOLOv1):
        def __init__(self, s_h, s_w, b, num_cls):
            torch.nn.Module.__init__(self)
            self.num_grid_cell_in_height = s_h
            self.num_grid_cell_in_width = s_w
            self.num_anchor_box = b
            self.cls_list = [str(i) for i in range(num_cls)]
            self.cls2idx = {c: i for i, c in enumerate(self.cls_list)}
            self.num_cls = num_cls

        def forward(self, x):
            return self._y

    ref.HeadOnlyV2, ref.HeadOnlyV1 = HeadOnlyV2, HeadOnlyV1

    def head_only(case):
        if case.version == 2:
            return HeadOnlyV2(case.c, case.anchors)
        return HeadOnlyV1(case.s_h, case.s_w, case.a, case.c)

    ref.head_only = head_only


def reference_step(ref, case, dense_targets, lambdas, conf_thre, iou_thre):
    """One pass of the reference's own path over `case` on ref.device: get_loss + backward (models/yolov2.py:747-1140,
    autograd) and the predict -> per-image nms loop of detect (models/yolov2.py:694-731, models/utils.py:68-164).
    Returns (loss, dL/dy, list of kept (bbox, conf, cls_spec) per image)."""
    import torch
    m = ref.head_only(case)
    y = case.y.clone().to(ref.device).requires_grad_(True)
    m._y = y
    x = torch.zeros(case.n, case.height, case.width, 3, device=ref.device)
    loss = m.get_loss(x, *[t.to(ref.device) for t in dense_targets], **lambdas)
    loss.backward()
    kept = []
    with torch.no_grad():
        _, _, bbox, conf, _, spec = m.predict(x)
        bbox, conf, spec = bbox.cpu(), conf.cpu(), spec.cpu()  # the reference's nms is CPU-only (models/utils.py:136-138)
        for i in range(case.n):
            kept.append(ref.utils.nms(bbox[i], conf[i], spec[i], conf_thre, iou_thre))
    return loss.detach(), y.grad, kept
