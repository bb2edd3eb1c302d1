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

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib

_PLANS = {}      # (device, flags-relevant key, pointer tuple) -> (device table, n_chunks)
_BUFFERS = {}    # id(parameter) -> (weak reference to it, momentum buffer)   (persistent_momentum)
_MAX_PLANS = 8


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

    def _plan(self, ps, gs, bufs):
        dev = ps[0].device
        key = (dev.index, tuple(t.data_ptr() for t in ps), tuple(t.data_ptr() for t in gs),
               None if bufs is None else tuple(t.data_ptr() for t in bufs), tuple(t.numel() for t in ps))
        hit = _PLANS.get(key)
        if hit is not None:
            return hit
        lib = _lib.load()
        n = len(ps)
        sizes = np.array([t.numel() for t in ps], dtype=np.int64)
        pp = np.array(key[1], dtype=np.uint64)
        gp = np.array(key[2], dtype=np.uint64)
        bp = None if bufs is None else np.array(key[3], dtype=np.uint64)
        n_chunks = int(lib.yh_sgd_chunk_count(sizes.ctypes.data, n))
        table = np.zeros((max(n_chunks, 1), 4), dtype=np.uint64)  # 32-byte YhSgdChunk records
        _lib.check("yh_sgd_plan", lib.yh_sgd_plan(pp.ctypes.data, gp.ctypes.data, None if bp is None else bp.ctypes.data,
                                                  sizes.ctypes.data, n, table.ctypes.data, n_chunks))
        dtab = torch.from_numpy(table.view(np.int64)).to(dev)
        if len(_PLANS) >= _MAX_PLANS:
            _PLANS.pop(next(iter(_PLANS)))
        _PLANS[key] = (dtab, n_chunks)
        return _PLANS[key]

    @torch.no_grad()
    def step(self):
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        gs = [p.grad for p in ps]
        for p, g in zip(ps, gs):
            if not (p.is_cuda and g.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32):
                raise RuntimeError("odcp_b200.optim.SGD needs CUDA fp32 parameters and gradients (no CPU path)")
            if not (p.is_contiguous() and g.is_contiguous()):
                raise RuntimeError("parameters and gradients must be contiguous")
        bufs, flags = None, 1  # YH_SGD_FRESH_MOMENTUM: the reference's new-optimizer-per-iteration update
        if self.persistent_momentum and self.momentum != 0.0:
            flags = 0
            bufs = []
            for p, g in zip(ps, gs):
                ent = _BUFFERS.get(id(p))
                b = ent[1] if ent is not None and ent[0]() is p else None
                if b is None or b.shape != p.shape or b.device != p.device:
                    # torch: the first step's buffer is the (decayed) gradient; momentum * 0 + d gives the same
                    b = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    key = id(p)
                    _BUFFERS[key] = (weakref.ref(p, lambda _r, key=key: _BUFFERS.pop(key, None)), b)
                bufs.append(b)
        dev = ps[0].device
        with torch.cuda.device(dev):
            dtab, n_chunks = self._plan(ps, gs, bufs)
            _lib.check("yh_sgd_step", _lib.load().yh_sgd_step(dtab.data_ptr(), n_chunks, self.lr, self.momentum,
                                                              self.weight_decay, flags,
                                                              torch.cuda.current_stream().cuda_stream))


def reset_state():
    """Forget cached chunk tables and persistent momentum buffers (tests)."""
    _PLANS.clear()
    _BUFFERS.clear()
