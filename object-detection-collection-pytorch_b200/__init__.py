"""B200-native YOLO detection-head hot path (decode, IoU target assignment, loss forward and
backward, greedy NMS) behind the call signatures of hcnoh/object-detection-collection-pytorch.

Layout:
    csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/yolohead.h)
    _lib.py      ctypes loader for libyolohead.so (raises if the library is missing)
    ops.py       tensor-level wrappers over the C ABI
    models/      drop-in `YOLOv1` / `YOLOv2` / `utils.get_iou` / `utils.nms`
    targets.py   compact ground-truth records <-> the reference's dense target grids
    synthetic.py seeded VOC-shaped inputs
    dist.py      batch sharding across ranks
"""
__version__ = "0.1.0"
