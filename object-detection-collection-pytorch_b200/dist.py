"""Batch sharding of the head path across ranks (one process per GPU).

The path shards by image: every ground-truth box, predictor and NMS decision
belongs to one image; the only cross-image coupling is the denominators of the five loss
means, which are global box counts.  So each rank runs the train head on its contiguous image
shard with `m_local` = its boxes and `m_global` = the all-rank box count, and the only
data-path exchange is one all-reduce of six floats (five terms + loss).  The post-process needs
no communication.  The reference has no distributed code; this module is host-side plumbing
over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch

from . import targets


def image_shard(num_images, rank, world):
    """Contiguous image range [lo, hi) of `rank`: the first `num_images % world` ranks get one more."""
    base, extra = divmod(int(num_images), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_case(rec, gt_off, num_images, rank, world):
    """(records, offsets, (lo, hi)) of this rank's image shard, image indices rebased to the shard."""
    lo, hi = image_shard(num_images, rank, world)
    sub, off = targets.shard_records(rec, np.asarray(gt_off), lo, hi)
    return sub, off, (lo, hi)


def global_box_count(m_local, group=None, device=None):
    """Sum of the per-rank box counts: the `m_global` every rank passes to the train head."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(m_local)
    t = torch.tensor([int(m_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


def reduce_terms(terms, loss, group=None):
    """All-reduce (sum) of the per-rank partial means.  Each rank's kernel divides its partial
    sums by the GLOBAL denominators, so the per-rank terms simply add up to the terms of the
    unsharded batch.  Returns (terms[5], loss) tensors holding the global values on every rank."""
    import torch.distributed as dist
    buf = torch.cat([terms.reshape(5), loss.reshape(1)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, group=group)
    return buf[:5], buf[5]


class PeerExchange:
    """Peer-mapped exchange buffers for the in-kernel reduction of the loss terms (include/yolohead.h
    YhExchange, csrc/yh_finalize.cuh): every rank allocates yh_exchange_bytes() of zeroed device memory, the
    CUDA IPC handles travel through the process group once (`all_gather_object`), every rank opens its peers'
    buffers (NVLink peer access on one node), and from then on a sharded train call sums the six loss sums of
    all ranks with 56-byte peer stores inside the kernel that ends the call -- no NCCL call per step.

        xch = PeerExchange(group)            # collective: every rank of the group must call it
        ops.train_head(..., m_global=M, exchange=xch)   /   ops.train_post(..., exchange=xch)

    Every rank must make the same sequence of exchanging calls (as with any collective).  One node only
    (CUDA IPC); `reduce_terms` below is the NCCL/gloo form of the same reduction."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_RANKS:
            raise ValueError("at most %d ranks" % _lib.MAX_RANKS)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(_lib.load().yh_exchange_bytes())
        # a private allocation (not a slice of a cached block another tensor may share) whose storage is exported
        self.local = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        handle = self.local.untyped_storage()._share_cuda_()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self._peers = []  # keep the opened storages alive
        ptrs = []
        for q, h in enumerate(handles):
            if q == self.rank:
                ptrs.append(self.local.data_ptr())
                continue
            # Open the peer's allocation with THIS rank's device current (first field: the device the handle is
            # opened on): cudaIpcOpenMemHandle then maps it for kernels of this device and enables NVLink peer
            # access lazily.  (Opened under the owner's device index -- what torch.multiprocessing does for
            # tensors it hands to another process -- the pointer would only be valid for kernels on that GPU.)
            with torch.cuda.device(dev):
                st = torch.UntypedStorage._new_shared_cuda(dev.index, *h[1:])
            self._peers.append(st)
            ptrs.append(st.data_ptr())
        self.struct = _lib.YhExchange()
        self.struct.rank, self.struct.world = self.rank, self.world
        for q, ptr in enumerate(ptrs):
            self.struct.slots[q] = ptr
        dist.barrier(group=group)  # every rank has opened every buffer before anyone uses them


def ddp_gradient_scale(world):
    """DDP averages parameter gradients over ranks; the shards' dL/dy are already normalised by
    the global box count, so multiplying the local loss by `world` makes the averaged gradient
    equal the single-process gradient of the whole batch (SURVEY 8e normalisation caveat)."""
    return float(world)
