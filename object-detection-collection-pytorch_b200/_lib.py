"""ctypes binding of libyolohead.so (include/yolohead.h).

There is deliberately no fallback: if the library has not been built, or a call fails, this
module raises.  Build with `python -m odcp_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YH_LIB_PATH") or os.path.join(PKG_DIR, "libyolohead.so")  # (override: experimental builds)

ABI_VERSION = 2
MAX_RANKS = 16

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t
_i64 = C.c_int64

class YhExchange(C.Structure):
    """include/yolohead.h YhExchange: peer-mapped exchange buffers of all ranks (host struct)."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("slots", C.c_void_p * MAX_RANKS)]


# name -> (restype, argtypes); mirrors include/yolohead.h one to one
SIGNATURES = {
    "yh_abi_version": (_i, []),
    "yh_last_error": (C.c_char_p, []),
    "yh_train_workspace_bytes": (_sz, []),
    "yh_v2_train": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_v1_train": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_v2_train_overlapped": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_v1_train_overlapped": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_exchange_bytes": (_sz, []),
    "yh_train_post_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "yh_v2_train_post": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p,
                              _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_v2_train_sharded": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _sz, _p]),
    "yh_v1_train_sharded": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _sz, _p]),
    "yh_v2_decode": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "yh_v1_decode": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "yh_compact_workspace_bytes": (_sz, [_i, _i]),
    "yh_compact_targets": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "yh_build_targets": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, C.c_double, C.c_double, _p, _p, _p, _p]),
    "yh_postprocess_workspace_bytes": (_sz, [_i, _i]),
    "yh_v2_postprocess": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_v1_postprocess": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "yh_match_detections": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p]),
    "yh_nms": (_i, [_p, _p, _p, _i, _i, _f, _f, _i, _p, _p, _p, _sz, _p]),
    "yh_iou": (_i, [_p, _p, _i64, _p, _p]),
    "yh_iou_f64": (_i, [_p, _p, _i64, _p, _p]),
    "yh_scale_inplace": (_i, [_p, _i64, _p, _p]),
    "yh_sgd_step": (_i, [_p, _p, _p, _p, _i, _f, _f, _f, _i, _p]),
}

_lock = threading.Lock()
_lib = None


class YoloHeadError(RuntimeError):
    """A libyolohead call returned a negative status."""

    def __init__(self, func, code, message):
        super().__init__("%s failed (%d): %s" % (func, code, message))
        self.code = code


def load():
    """Load (once) and return the ctypes library handle.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libyolohead.so is not built (%s). There is no CPU fallback: run "
                "`python -m odcp_b200.build` or `__graft_entry__.build()` first." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        ver = lib.yh_abi_version()
        if ver != ABI_VERSION:
            raise RuntimeError("libyolohead ABI %d != binding ABI %d: rebuild" % (ver, ABI_VERSION))
        _lib = lib
    return _lib


def check(func, code):
    if code != 0:
        msg = load().yh_last_error()
        raise YoloHeadError(func, code, msg.decode("utf-8", "replace") if msg else "")


def call(name, *args):
    lib = load()
    check(name, getattr(lib, name)(*args))
