"""Fused SGD step behind torch.optim.SGD's call signature, as the reference uses it.

The reference builds a NEW optimizer in every iteration (models/yolov2.py:1253-1272,
models/yolov1.py run_one_epoch likewise):

    opt = SGD(self.parameters(), lr=..., momentum=0.9, weight_decay=5e-4)
    opt.zero_grad(); loss.backward(); opt.step()

so every step is the optimizer's first one and the momentum buffer never accumulates: the update is
p -= lr * (g + weight_decay * p).  `SGD` here is a drop-in for that name (the maintainer's change is
`from odcp_b200.optim import SGD` instead of `from torch.optim import SGD`): same constructor, same
zero_grad()/step(), the same update -- all parameter tensors in ONE kernel launch (yh_sgd_step)
instead of three foreach passes.  `persistent_momentum=True` keeps the momentum buffers across
re-created optimizers (keyed by parameter), i.e. real SGD with momentum: a change of training semantics,
hence an explicit flag.

No CPU path: parameters and gradients must be CUDA fp32 tensors.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import _lib

_BUFFERS = {}    # id(parameter) -> (weak reference to it, momentum buffer)   (persistent_momentum)


def _dense(t):
    """Non-overlapping and dense in SOME dimension order (contiguous, channels_last, ...): the elements fill
    numel() consecutive floats, which is all an elementwise update needs."""
    dims = sorted((st, n) for st, n in zip(t.stride(), t.shape) if n > 1)
    want = 1
    for st, n in dims:
        if st != want:
            return False
        want *= n
    return True


def _same_layout(a, b):
    """Same element order in memory (strides of size-1 dimensions are arbitrary and ignored)."""
    return a.shape == b.shape and all(x == y for x, y, n in zip(a.stride(), b.stride(), a.shape) if n > 1)


class SGD:
    def __init__(self, params, lr, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False, *,
                 persistent_momentum=False):
        if dampening != 0 or nesterov:
            raise NotImplementedError("the reference uses plain SGD (dampening=0, nesterov=False)")
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("negative hyper-parameter")
        self.params = [p for p in params]
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.persistent_momentum = bool(persistent_momentum)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_()
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        gs = [p.grad for p in ps]
        for p, g in zip(ps, gs):
            if not (p.is_cuda and g.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32):
                raise RuntimeError("odcp_b200.optim.SGD needs CUDA fp32 parameters and gradients (no CPU path)")
            # elementwise: any dense layout will do (channels_last conv weights included) as long as the
            # parameter and its gradient share it
            if not (_dense(p) and _same_layout(p, g)):
                raise RuntimeError("a parameter and its gradient must be dense and share one memory layout")
        bufs, flags = None, 1  # YH_SGD_FRESH_MOMENTUM: the reference's new-optimizer-per-iteration update
        if self.persistent_momentum and self.momentum != 0.0:
            flags = 0
            bufs = []
            for p, g in zip(ps, gs):
                ent = _BUFFERS.get(id(p))
                b = ent[1] if ent is not None and ent[0]() is p else None
                if b is None or b.shape != p.shape or b.device != p.device or not _same_layout(b, p):
                    # torch: the first step's buffer is the (decayed) gradient; momentum * 0 + d gives the same
                    b = torch.zeros_like(p, memory_format=torch.preserve_format)
                    key = id(p)
                    _BUFFERS[key] = (weakref.ref(p, lambda _r, key=key: _BUFFERS.pop(key, None)), b)
                bufs.append(b)
        dev = ps[0].device
        n = len(ps)
        arr = np.empty((4, n), dtype=np.uint64)   # device pointers of p, g, buf and the element counts
        arr[0] = [t.data_ptr() for t in ps]
        arr[1] = [t.data_ptr() for t in gs]
        arr[2] = [t.data_ptr() for t in bufs] if bufs is not None else 0
        arr[3] = [t.numel() for t in ps]
        with torch.cuda.device(dev):
            _lib.check("yh_sgd_step", _lib.load().yh_sgd_step(
                arr[0].ctypes.data, arr[1].ctypes.data, arr[2].ctypes.data if bufs is not None else None,
                arr[3].ctypes.data, n, self.lr, self.momentum, self.weight_decay, flags,
                torch.cuda.current_stream().cuda_stream))


def reset_state():
    """Forget the persistent momentum buffers (tests)."""
    _BUFFERS.clear()
