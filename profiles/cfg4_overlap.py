"""How much of DDP's gradient all-reduce hides behind the backward of BASELINE config 4 (run under torchrun).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/cfg4_overlap.py [--bucket-mb B]

Traces a few steps of odcp_b200.train_step.ShardedTrainStep with torch.profiler (CUPTI kernel records; there is no
nsys in the image) and prints one JSON line on rank 0: per step, the device time of the NCCL kernels, the part of it
during which a compute kernel (cuDNN / aten / libyolohead) runs on another stream at the same time, the exposed rest,
and the names of the repo's own kernels seen in the step.  Profiler overhead inflates host-side gaps: the plain
step time comes from `bench.py --workload cfg4`, this script only apportions the all-reduce.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import odcp_b200  # noqa: F401
from odcp_b200 import dist as yh_dist, synthetic, targets
from odcp_b200.models.layout import use_channels_last_head
from odcp_b200.train_step import ShardedTrainStep, YOLOv2Net


def union_length(iv):
    iv = sorted(iv)
    total, cur_lo, cur_hi = 0.0, None, None
    for lo, hi in iv:
        if cur_hi is None or lo > cur_hi:
            if cur_hi is not None:
                total += cur_hi - cur_lo
            cur_lo, cur_hi = lo, hi
        else:
            cur_hi = max(cur_hi, hi)
    if cur_hi is not None:
        total += cur_hi - cur_lo
    return total


def overlap_length(a, b):
    """Length of (union of a) intersected with (union of b)."""
    return union_length(a) + union_length(b) - union_length(a + b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--bucket-mb", type=int, default=0)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    case = synthetic.make_case("cfg4", 2, args.batch, 13, 13, 5, 20, 416, 416, seed=104 + rank)
    torch.manual_seed(11)
    model = use_channels_last_head(YOLOv2Net()).to(dev)
    gt, off = targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
    m_global = yh_dist.global_box_count(case.m, device=dev)
    trainer = ShardedTrainStep(model, bucket_cap_mb=args.bucket_mb or None, static_box_count=m_global)
    x = torch.rand(args.batch, 416, 416, 3, device=dev) * 255.0
    for _ in range(4):
        trainer.step(x, gt, off, case.m)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            trainer.step(x, gt, off, case.m)
        torch.cuda.synchronize()
    nccl, compute, ours = [], [], {}
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        if "Memcpy" in name or "Memset" in name:
            continue
        lo = ev.time_range.start
        hi = ev.time_range.end
        if "nccl" in name.lower():
            nccl.append((lo, hi))
        else:
            compute.append((lo, hi))
            if "yh_" in name:
                key = name[name.index("yh_"):].split("<")[0].split("(")[0][:48]
                ours[key] = ours.get(key, 0) + 1
    span = max(h for _, h in nccl + compute) - min(l for l, _ in nccl + compute)
    n = args.steps
    out = dict(what="cfg4 step under torch.profiler: NCCL all-reduce time vs concurrent compute, per step",
               world=world, batch_per_gpu=args.batch, bucket_mb=args.bucket_mb or 25, steps=n,
               allreduce_bytes_per_step=trainer.allreduce_bytes() if world > 1 else 0,
               traced_ms_per_step=span / n / 1e3,
               nccl_kernels_per_step=len(nccl) / n,
               nccl_ms_per_step=union_length(nccl) / n / 1e3,
               nccl_overlapped_with_compute_ms_per_step=overlap_length(nccl, compute) / n / 1e3,
               nccl_exposed_ms_per_step=(union_length(nccl) - overlap_length(nccl, compute)) / n / 1e3,
               compute_busy_ms_per_step=union_length(compute) / n / 1e3,
               own_kernels_per_step={k: v / n for k, v in sorted(ours.items())})
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
