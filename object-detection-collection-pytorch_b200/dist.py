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


def ddp_gradient_scale(world):
    """DDP averages parameter gradients over ranks; the shards' dL/dy are already normalised by
    the global box count, so multiplying the local loss by `world` makes the averaged gradient
    equal the single-process gradient of the whole batch (SURVEY 8e normalisation caveat)."""
    return float(world)
