#!/usr/bin/env python
"""Benchmark of the YOLO detection-head hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): YOLO head images/sec for decode + target assignment + loss + backward
(one fused launch) + confidence threshold + NMS (one launch), on the configuration the
north-star target is quoted on: YOLOv2 13x13 grid, 5 anchors, C=20, batch 256 per GPU, synthetic
VOC-shaped data.  One "step" = one train-head call + one post-process call over one batch.

  value      device-timed throughput, inputs resident in HBM, CUDA-graph replay over rotating
             buffer sets larger than L2, max over ranks
  e2e        same step through the host-buffer entry point (pinned host memory -> H2D -> kernels
             -> D2H of loss, dL/dy and detections), wall clock between device synchronisations
  roofline   the fused train-head kernel alone: algorithmic bytes / launch duration vs measured HBM
  cpu_baseline  the CPU port of the reference's torch path (oracle/), bounded sample, rank 0
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "yolo_head_images_per_sec"
UNIT = "images/s"
CONF_THRE, IOU_THRE = 0.5, 0.45
MAX_OUT = 128
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="0: 4000 (headline, cfg5) / 30 (cfg4)")
    ap.add_argument("--warmup", type=int, default=-1, help="-1: 50 (headline, cfg5) / 5 (cfg4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the workload's own: 256 / 512)")
    ap.add_argument("--sets", type=int, default=8, help="rotating buffer sets (total must exceed L2)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="images per step of the CPU legs (0: the full batch for the headline workload, 8 for cfg5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS),
                    help="headline: the configuration the north-star target is quoted on (YOLOv2 13x13x5, C=20, 1..5 "
                         "boxes/image); cfg5: BASELINE config 5 (19x19, batch 512 unless --batch, 50..100 boxes/image)")
    ap.add_argument("--unfused", action="store_true",
                    help="a step = the two separate calls (yh_v2_train + yh_v2_postprocess: 3 launches, y read twice) "
                         "instead of the fused step (yh_v2_train_post: one kernel per image batch + finalize, y read once)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the workload's batch is the GLOBAL batch, every rank takes batch / N images "
                         "(default: weak scaling, the batch per GPU is fixed)")
    ap.add_argument("--no-collective", action="store_true",
                    help="N > 1: leave the in-kernel peer-memory reduction of the loss terms out of the steps")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 400)")
    ap.add_argument("--e2e-return-dy", action="store_true",
                    help="e2e: also copy dL/dy (21.6 MB) back to the host every step.  Default: the gradient stays on "
                         "the device (where the backbone's backward consumes it); the step's host-side result is the "
                         "loss, its five terms and the kept boxes")
    ap.add_argument("--repeats", type=int, default=0,
                    help="repetitions of the timed K-step region (median reported); 0 = enough for ~25 ms of "
                         "device time, between 3 and 15")
    ap.add_argument("--gate-us", type=float, default=150.0,
                    help="device-side gate in front of every timed region: the GPU spins this long while the host "
                         "enqueues the start event, the K steps and the stop event, so the region holds no launch gap")
    ap.add_argument("--torch-sgd", action="store_true", help="cfg4: torch.optim.SGD instead of the fused SGD step")
    ap.add_argument("--bucket-mb", type=int, default=0, help="cfg4: DDP bucket size (0: DDP's default, 25 MB)")
    args = ap.parse_args()
    heavy = args.workload == "cfg4"
    if args.steps <= 0:
        args.steps = 30 if heavy else 4000
    if args.warmup < 0:
        args.warmup = 5 if heavy else 50
    return args


def _headline_case(n):
    from odcp_b200 import synthetic
    return synthetic.headline(n=n)


def _cfg5_case(n):
    from odcp_b200 import synthetic
    return synthetic.cfg5(n=n)


WORKLOADS = {
    "headline": dict(case=_headline_case, batch=256, grid=[13, 13], image=[416, 416], boxes="U{1..5}",
                     cands="~50 of 845", name="yolov2_head_13x13x5_c20"),
    "cfg5": dict(case=_cfg5_case, batch=512, grid=[19, 19], image=[608, 608], boxes="U{50..100}",
                 cands="~100 of 1805", name="yolov2_head_19x19x5_c20_dense_gt"),
    # BASELINE config 4: the full training step (run_cfg4 below; its own metric)
    "cfg4": dict(case=None, batch=64, grid=[13, 13], image=[416, 416], boxes="U{1..5}", cands="-",
                 name="yolov2_full_training_step"),
}


def workload_config(workload, batch):
    """`config` of the JSON line: the workload, identical in both arms (ours and --impl reference)."""
    w = WORKLOADS[workload]
    return {
        "workload": "%s_b%d_train(decode+assign+loss+bwd)+postprocess(conf%.2f,nms_iou%.2f)"
                    % (w["name"], batch, CONF_THRE, IOU_THRE),
        "batch_per_gpu": batch, "grid": w["grid"], "anchors": 5, "classes": 20, "image": w["image"],
        "gt_boxes_per_image": w["boxes"], "candidates_per_image": w["cands"],
    }


def run_description(sets, set_bytes, fused, collective):
    """How our arm ran the workload (top-level `run` of the line; not part of `config`)."""
    run = {
        "l2": "%d rotating buffer sets per GPU, inputs + outputs + workspace distinct per set (%.0f MB > 126 MB L2)" % (
            sets, sets * set_bytes / 1e6),
        "step": ("fused step: ONE call yh_v2_train_post = one kernel, one CTA per image (dense loss/gradient pass, the "
                 "image's records, threshold, NMS, class pick) + the one-warp finalize kernel; y read once" if fused else
                 "two separate calls: yh_v2_train (+ finalize kernel) + yh_v2_postprocess; y read twice"),
        "streams": ("one stream, programmatic dependent launches; a graph replay is 4 passes over the %d buffer sets; "
                    "every call but the first of a pass is an overlapped call (include/yolohead.h, THE OVERLAP "
                    "CONTRACT) and runs next to the tails of the kernels in front of it; the first call of every "
                    "pass waits for everything before it" % sets),
        "timing": ("the timed K-step region is rehearsed twice (its own graphs), starts behind a device-side gate "
                   "kernel (N > 1: and a one-float all-reduce that lines the ranks' streams up, both before the start "
                   "event) and is repeated; value = the median region, max over ranks per region"),
    }
    if collective is not None:
        run["collective"] = collective
    return run


# ------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/) -- the baseline arm
# ------------------------------------------------------------------------------------------
def make_cpu_step(workload, sample):
    """The reference's CPU implementation of the path over a `sample`-image batch of the workload, as a callable.
    With the staged reference (oracle/_ref, or /root/reference in the build container) it is the UNMODIFIED
    reference: HeadOnly.get_loss + loss.backward() + predict + the per-image nms loop of detect
    (oracle/refharness.py) -- kind "reference"; otherwise the oracle's port of the same op sequence (pinned
    bit-exact against the reference's outputs, tests/golden/) -- kind "port"."""
    from odcp_b200 import synthetic, targets
    from oracle import refharness as RH
    torch.set_num_threads(os.cpu_count() or 1)
    case = WORKLOADS[workload]["case"](sample)
    lam = synthetic.DEFAULT_LAMBDAS
    if RH.available():
        ref = RH.load_reference("cpu")
        dense = targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, case.version)
        import warnings
        warnings.filterwarnings("ignore", message="Using a target size")  # the reference's own broadcasting mse_loss
        return (lambda: RH.reference_step(ref, case, dense, lam, CONF_THRE, IOU_THRE)), "reference", \
            "the unmodified reference (staged copy, %s): get_loss + backward + predict + per-image nms" % ref.location
    from oracle import yolo_head_oracle as O

    def port():
        O.train_head_dense(case, lam)
        O.postprocess_torch(case.y, case.height, case.width, case.version, case.anchors, CONF_THRE, IOU_THRE)
    return port, "port", "oracle/ torch-CPU port of get_loss + backward and the per-image nms loop (reference not staged)"


def cpu_sample_size(workload, batch, requested):
    """Images per CPU step: the full batch where the reference can hold it (its per-box replication needs
    49*S*S*A floats per box, several times over: headline batch 256 ~3 GB), a bounded sample otherwise."""
    if requested:
        return min(requested, batch)
    return batch if workload == "headline" else 8


def time_cpu(workload, sample, reps, warm):
    fn, kind, what = make_cpu_step(workload, sample)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts, kind, what


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads,
    on our arm's workload (the full batch for the headline; for cfg5 the reference cannot hold the full batch --
    SURVEY 8(a-4): 18 GB of replicated predictions -- and an 8-image sample is timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch or WORKLOADS[args.workload]["batch"]
    sample = cpu_sample_size(args.workload, batch, args.cpu_sample)
    steps = max(1, min(args.steps, 40))
    warm = max(0, min(args.warmup, 10))
    ts, kind, what = time_cpu(args.workload, sample, steps, warm)
    total = float(np.sum(ts))
    value = sample * len(ts) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(ts), "warmup": warm, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d images per step (%s); %s; torch %s CPU, %d threads, os.cpu_count()=%s"
                                   % (sample, "the full batch" if sample == batch else "a sample of the batch-%d workload" % batch,
                                      what, torch.__version__, cores, os.cpu_count())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)



# ------------------------------------------------------------------------------------------
# BASELINE config 4: the full YOLOv2 training step, batch-sharded under DDP
# ------------------------------------------------------------------------------------------
CFG4_METRIC = "yolov2_train_step_images_per_sec"


def cfg4_config(batch):
    return {"workload": "yolov2_darknet19_416_b%d_full_training_step(cudnn_convs+fused_head+backward+ddp_allreduce+sgd)" % batch,
            "batch_per_gpu": batch, "grid": [13, 13], "anchors": 5, "classes": 20, "image": [416, 416],
            "gt_boxes_per_image": "U{1..5}", "parameters": 67147837}


def run_cfg4_reference(args):
    """The unmodified reference's run_one_epoch over ONE batch on the host cores (a bounded sample: 4 images)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from odcp_b200 import synthetic, targets
    from oracle import refharness as RH
    batch = args.batch or WORKLOADS["cfg4"]["batch"]
    sample = min(args.cpu_sample or 4, batch)
    torch.set_num_threads(os.cpu_count() or 1)
    if not RH.available():
        emit({"impl": "reference", "unavailable": "the reference is not staged (oracle/_ref) on this box"})
        return
    ref = RH.load_reference("cpu")
    cls_list = [str(i) for i in range(20)]
    torch.manual_seed(1234)
    model = ref.yolov2.YOLOv2(cls_list, {c: i for i, c in enumerate(cls_list)})
    case = synthetic.make_case("cfg4", 2, sample, 13, 13, 5, 20, 416, 416, seed=104)
    x = torch.rand(sample, 416, 416, 3, generator=torch.Generator().manual_seed(7)) * 255.0
    dense = targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, 2)

    class OneBatch:
        dataset = range(sample)

        def __iter__(self):
            yield (x, *dense)

    import warnings
    warnings.filterwarnings("ignore", message="Using a target size")
    steps, warm = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    ts = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        model.run_one_epoch(1, OneBatch(), lr=1e-3, train=True, **synthetic.DEFAULT_LAMBDAS)
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    total = float(np.sum(ts))
    value = sample * len(ts) / total
    cores = torch.get_num_threads()
    emit({"impl": "reference", "metric": CFG4_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ts),
          "warmup": warm, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg4_config(batch),
          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                           "sample": "%d images per step (a sample of the batch-%d workload); the unmodified reference's "
                                     "run_one_epoch over one batch (get_loss, SGD, backward, step) on torch CPU, %d threads"
                                     % (sample, batch, cores)},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def run_cfg4(args):
    """One step = what the reference's run_one_epoch does per batch (models/yolov2.py:1237-1272) on this rank's shard
    of the batch: forward through the convolutions (stock cuDNN, channels_last, TF32 as torch defaults), the fused
    train head, backward (DDP's bucketed NCCL all-reduce of the 67.15 M parameter gradients overlapped with it), and
    the fused SGD step.  odcp_b200.train_step has the step; this function times it."""
    if args.impl == "reference":
        run_cfg4_reference(args)
        return
    from odcp_b200 import ops, synthetic, targets
    from odcp_b200.models.layout import use_channels_last_head
    from odcp_b200.optim import SGD
    from odcp_b200.train_step import ShardedTrainStep, YOLOv2Net

    world, rank, local_rank = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch or WORKLOADS["cfg4"]["batch"], args.steps, max(args.warmup, 3)
    case = synthetic.make_case("cfg4", 2, B, 13, 13, 5, 20, 416, 416, seed=104 + rank)
    torch.manual_seed(11)
    model = use_channels_last_head(YOLOv2Net()).to(dev)
    n_par = sum(p.numel() for p in model.parameters())
    gt, off = targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
    from odcp_b200 import dist as yh_dist
    m_global = yh_dist.global_box_count(case.m, device=dev)
    trainer = ShardedTrainStep(model, optimizer_cls=torch.optim.SGD if args.torch_sgd else SGD,
                               bucket_cap_mb=args.bucket_mb or None, static_box_count=m_global)
    x_host = (torch.rand(B, 416, 416, 3, generator=torch.Generator().manual_seed(7 + rank)) * 255.0).pin_memory()
    x = x_host.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()

    def timed(n, from_host):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        last = None
        for _ in range(n):
            if from_host:
                x.copy_(x_host, non_blocking=True)
            last = trainer.step(x, gt, off, case.m)
            if from_host:
                last = float(last.item())  # the step's result on the host, every step
        e.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        t = torch.tensor([a.elapsed_time(e), wall * 1e3], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), last

    timed(W, False)
    sampler = ClockSampler(nvml_index(local_rank))
    sampler.start()
    ms, _, last = timed(K, False)
    sampler.stop()
    loss_share = float(last.item())
    ke = max(3, min(K, 20))
    timed(2, True)
    _, e2e_ms, _ = timed(ke, True)

    # the head call inside the step, alone (stream-ordered, this rank's head tensor)
    with torch.no_grad():
        y = model(x).contiguous()
    kw = dict(version=2, img_hw=(416, 416), anchors=synthetic.YOLOV2_ANCHORS, lambdas=synthetic.DEFAULT_LAMBDAS, m_global=m_global)
    out = ops.train_head(y, gt, off, **kw)
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        ops.train_head(y, gt, off, out=out, **kw)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(8):
                ops.train_head(y, gt, off, out=out, **kw)
        g.replay()
        st.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(20):
            g.replay()
        e.record(st)
        st.synchronize()
    head_ms = a.elapsed_time(e) / 160
    head_bytes = 2 * y.numel() * 4 + 48 * case.m
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else FALLBACK_HBM_GBS
    if rank == 0:
        emit({
            "metric": CFG4_METRIC, "value": B * world * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (convolutions: TF32 tensor cores, torch's cuDNN default; head, loss, gradient, SGD: fp32)",
            "data": "synthetic", "config": cfg4_config(B),
            "run": {"optimizer": "torch.optim.SGD re-created per step (the reference)" if args.torch_sgd else
                                 "odcp_b200.optim.SGD (yh_sgd_step, one launch) re-created per step",
                    "ddp": None if world == 1 else "DistributedDataParallel, gradient_as_bucket_view, bucket %s MB; "
                           "all-reduce of %d bytes per step overlapped with the backward" % (args.bucket_mb or 25, 4 * n_par),
                    "l2": "activations of one step (GBs) far exceed the 126 MB L2",
                    "box_count": "the all-rank box count is known up front (static shards): no per-step all-reduce for it"},
            "allreduce_bytes_per_step": 0 if world == 1 else 4 * n_par, "parameters": n_par,
            "clocks": sampler.summary("the %d timed steps" % K),
            "e2e": {"value": B * world * ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "steps": ke, "ms_per_step": e2e_ms / ke,
                    "path": "ShardedTrainStep.step with the image batch copied from pinned host memory and the loss read "
                            "back every step (wall clock, max over ranks)"},
            "gpu_launches": 3 * K,  # ours per step: yh_train_kernel, yh_train_finalize_kernel, yh_sgd_kernel (+ cuDNN/NCCL/aten)
            "roofline": {"bound": "hbm", "kernel": "yh_train_kernel + finalize inside the step (the convolutions are cuDNN's: out of scope)",
                         "achieved": head_bytes / (head_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": head_bytes / (head_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": head_bytes, "us_per_launch": head_ms * 1e3,
                         "share_of_step": head_ms / (ms / K)},
            "cpu_baseline": None, "loss_share_of_rank0": loss_share,
        })
    if dist is not None:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                bits = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join()

    def summary(self, window):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples")}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "window": window}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs NVML reports as local to its GPU (before any pinned host memory is
    allocated), so the host side of the e2e path does not cross sockets.  Best effort: silently keeps
    the current affinity if NVML, the topology or the container's CPU set do not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(nvml_index(local_rank))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 8)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
        return sorted(pick) if pick else None
    except Exception:  # noqa: BLE001
        return None


def nvml_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


def source_stamp():
    """Hash of the kernel sources (csrc/*.cu, *.cuh, include/*.h): ties an ncu capture to the build it profiled."""
    import glob
    import hashlib
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, "object-detection-collection-pytorch_b200", "csrc")
    for f in sorted(glob.glob(os.path.join(pkg, "*.cu")) + glob.glob(os.path.join(pkg, "*.cuh")) +
                    glob.glob(os.path.join(ROOT, "include", "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the ncu --set full capture under profiles/
    (profiles/traffic.json, written by profiles/ncu_traffic.py).  Only a capture of THIS build counts: the file
    carries the source stamp of the build it profiled, a stale one is reported as null."""
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(prof))
    except Exception:  # noqa: BLE001
        return None, "no profiles/traffic.json"
    if t.get("source_stamp") != source_stamp():
        return None, "profiles/traffic.json is from another build (stamp %s != %s)" % (t.get("source_stamp"), source_stamp())
    return t.get("dram_bytes_per_launch"), t.get("source")


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, for one) print to stdout; the contract is ONE JSON line there.
    Everything else goes to stderr: fd 1 is pointed at fd 2 until emit() writes the line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    quiet_stdout()
    if args.workload == "cfg4":
        run_cfg4(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    from odcp_b200 import ops, synthetic, targets
    from odcp_b200.host import HostHeadPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    bind_to_gpu_numa_node(local_rank)  # (also with one rank: pinned host buffers on the GPU's own socket)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    K, W, R = args.steps, max(args.warmup, 3), args.sets
    B = args.batch or WORKLOADS[args.workload]["batch"]
    if args.strong:
        B = max(1, B // world)
    lam = synthetic.DEFAULT_LAMBDAS
    case = WORKLOADS[args.workload]["case"](B)
    conf_thre, iou_thre = CONF_THRE, IOU_THRE
    kw = dict(img_hw=(case.height, case.width), anchors=case.anchors)
    m_local = case.m
    m_global = m_local * world  # every rank holds a shard with the same box count (weak scaling)
    fused = not args.unfused

    # Sharded batch (N > 1): the six loss sums of all ranks are summed INSIDE the step's last kernel over NVLink
    # peer memory (odcp_b200.dist.PeerExchange, csrc/yh_finalize.cuh) -- the scalar-loss reduction of the
    # north star, in the timed region, every step, without an NCCL call.
    xch, xch_error = None, None
    if dist is not None and not args.no_collective:
        from odcp_b200.dist import PeerExchange
        try:
            xch = PeerExchange(device=dev)
        except Exception as e:  # noqa: BLE001  (no peer access / IPC on this box: say so in the line, time without it)
            xch_error = repr(e)
        # all ranks or none: a rank that could not map its peers takes everyone to the plain path
        ok = torch.tensor([0 if xch is None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            xch = None

    # rotating buffer sets: distinct addresses (inputs, outputs, workspace), > L2 in total
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    y0 = case.y.to(dev)
    sets = [dict(y=y0.clone(), gt=gt.clone(), off=off.clone(), res=None, res_nc=None, tr=None, tr_ov=None,
                 post=None) for _ in range(R)]
    set_bytes = 2 * y0.numel() * 4

    stream = torch.cuda.Stream(dev)

    # One step = train head (decode + assignment + loss + dL/dy) + post-process (threshold + NMS + class pick) of
    # the same head tensor.  Default: the fused step, ONE call (yh_v2_train_post): one kernel in which the CTA that
    # holds an image in shared memory does all of it (+ the one-warp finalize kernel) -- 2 launches, y read once.
    # --unfused: the two separate calls (3 launches, y read twice).
    # Overlap of consecutive kernels: every kernel of the path is a programmatic dependent launch.  A call may
    # additionally promise that none of its buffers is in use by any kernel launched since the last call without
    # that promise (rotating sets, include/yolohead.h "THE OVERLAP CONTRACT"); it then runs next to the tails of
    # the kernels in front of it.  The first call of every pass over the R sets makes no promise -- it waits at
    # its start for everything before it -- which bounds the overlap chain to the rotation length.
    def step(s, first, exchange=xch, key="res"):
        if fused:
            s[key] = ops.train_post(s["y"], s["gt"], s["off"], lambdas=lam, conf_thre=conf_thre, iou_thre=iou_thre,
                                    m_global=m_global, max_out=MAX_OUT, want_cls_spec=False, out=s[key],
                                    overlapped=not first, exchange=exchange, **kw)
        else:
            tr = ops.train_head(s["y"], s["gt"], s["off"], version=2, lambdas=lam, m_global=m_global,
                                out=(s[key] or {}).get("train"), input_ready=not first, exchange=exchange, **kw)
            po = ops.postprocess(s["y"], version=2, conf_thre=conf_thre, iou_thre=iou_thre, max_out=MAX_OUT,
                                 want_cls_spec=False, out=(s[key] or {}).get("post"), input_ready=True, **kw)
            s[key] = dict(train=tr, post=po)

    # the kernels on their own (for the roofline lines): stream-ordered and overlapped
    def train_only(s, overlapped=False):
        key = "tr_ov" if overlapped else "tr"
        s[key] = ops.train_head(s["y"], s["gt"], s["off"], version=2, lambdas=lam, m_global=m_global, out=s[key],
                                input_ready=overlapped, **kw)

    def post_only(s):
        s["post"] = ops.postprocess(s["y"], version=2, conf_thre=conf_thre, iou_thre=iou_thre, max_out=MAX_OUT,
                                    want_cls_spec=False, out=s["post"], **kw)

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        return g

    with torch.cuda.stream(stream):
        for s in sets:  # eager warm-up: allocates outputs/workspaces before any capture
            step(s, True)
            stream.synchronize()
            step(s, False)  # (the overlapped form as well: it owns its workspace)
            if xch is not None:
                step(s, True, exchange=None, key="res_nc")
                stream.synchronize()
                step(s, False, exchange=None, key="res_nc")
            train_only(s)
            train_only(s, True)
            post_only(s)
        stream.synchronize()
        # one graph = ROUNDS passes over the R buffer sets
        ROUNDS = 4
        G = R * ROUNDS
        g_full = capture(lambda: [step(s, i == 0) for _ in range(ROUNDS) for i, s in enumerate(sets)])
        # step counts that are not a multiple of G: one more graph of exactly the remaining steps (same chain)
        g_rem = {}
        if K % G:
            g_rem[K % G] = capture(lambda: [step(sets[j % R], j % R == 0) for j in range(K % G)])
        if W % G and W % G not in g_rem:
            g_rem[W % G] = capture(lambda: [step(sets[j % R], j % R == 0) for j in range(W % G)])
        g_train = capture(lambda: [train_only(s) for s in sets])
        g_train_ov = capture(lambda: [train_only(s, i > 0) for i, s in enumerate(sets)])
        g_post = capture(lambda: [post_only(s) for s in sets])
        g_step_iso = capture(lambda: [step(s, True) for s in sets])
        g_full_nc = None
        if xch is not None:  # the same chain without the exchange, to report what the collective costs
            g_full_nc = capture(lambda: [step(s, i == 0, exchange=None, key="res_nc") for _ in range(ROUNDS) for i, s in enumerate(sets)])

        def run_steps(n, full=g_full):
            for _ in range(n // G):
                full.replay()
            if n % G:
                g_rem[n % G].replay()

        def barrier():
            if dist is not None:
                dist.barrier()

        # Warm-up: W steps, then two rehearsals of the timed region itself, so every graph the region
        # replays (the remainder graph of K % G steps included) has been uploaded and run before time starts.
        run_steps(W)
        stream.synchronize()
        sm_hz = 1e6 * float(torch.cuda.get_device_properties(dev).clock_rate) / 1e3  # clock_rate is in kHz
        gate_cycles = int(args.gate_us * 1e-6 * sm_hz)

        align = torch.zeros(1, device=dev)

        def timed_region(ev0, ev1, full=g_full):
            """Exactly K steps between two events.  A gate kernel keeps the GPU busy while the host
            enqueues ev0, the graph launches and ev1: the region then starts with its first kernel already
            queued behind the event (no idle-GPU launch latency inside it)."""
            if gate_cycles > 0:
                torch.cuda._sleep(gate_cycles)
            if dist is not None:
                # the ranks leave the host barrier tens of microseconds apart; a tiny all-reduce on the stream (outside
                # the region) lines their streams up to a few microseconds, so a short region measures the steps and
                # not the wait of the early ranks' exchange for the late ones
                dist.all_reduce(align)
            ev0.record(stream)
            run_steps(K, full)
            ev1.record(stream)

        for _ in range(2):
            barrier()
            timed_region(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            stream.synchronize()

        # ---- timed: `n_regions` regions of exactly K steps each, device time, barrier + synchronize on both
        # sides of every region; per region the MAX over ranks, reported: the median region (min/max next to it)
        est_ms = max(K * 0.012, 1e-3)
        n_regions = args.repeats or int(min(15, max(3, round(25.0 / est_ms))))

        def time_regions(full):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_regions)]
            for ev0, ev1 in evs:
                barrier()
                torch.cuda.synchronize()
                timed_region(ev0, ev1, full)
                torch.cuda.synchronize()
                barrier()
            t = torch.tensor([a.elapsed_time(b) for a, b in evs], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return sorted(float(v) for v in t.tolist())

        sampler = ClockSampler(nvml_index(local_rank))
        barrier()
        torch.cuda.synchronize()
        sampler.start()
        region_ms = time_regions(g_full)
        window = "the %d timed regions" % n_regions
        if len(sampler.samples) < 5:  # the timed regions were too short to sample: keep the same load on
            barrier()
            for _ in range(int(0.25 / (G * 12e-6)) + 1):
                g_full.replay()
            stream.synchronize()
            window = "the %d timed regions + 0.25 s of the same graph replayed right after them" % n_regions
        sampler.stop()
        ms = region_ms[len(region_ms) // 2]
        ms_nc = None
        if g_full_nc is not None and K % G == 0:
            r_nc = time_regions(g_full_nc)
            ms_nc = r_nc[len(r_nc) // 2]
        res0 = sets[0]["res"]
        loss_value = float(res0["train"]["loss"].item())
        terms_value = [float(v) for v in res0["train"]["terms"].tolist()]

        # ---- per-kernel durations, same rotation
        def time_graph(g, reps):
            for _ in range(3):
                g.replay()
            stream.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if gate_cycles > 0:
                torch.cuda._sleep(gate_cycles)
            a.record(stream)
            for _ in range(reps):
                g.replay()
            b.record(stream)
            stream.synchronize()
            return a.elapsed_time(b) / (reps * R)

        reps = max(10, min(K // R, 500))
        train_ms = time_graph(g_train, reps)
        post_ms = time_graph(g_post, reps)
        train_ov_ms = time_graph(g_train_ov, reps)
        step_iso_ms = time_graph(g_step_iso, reps)

    images = B * world
    value = images * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (fused train head)
    kept = int(res0["post"]["keep_cnt"].clamp(max=MAX_OUT).sum().item())
    p_bytes = case.image_bytes
    train_bytes = 2 * p_bytes * B + 48 * m_local
    post_bytes = p_bytes * B + 28 * kept
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"

    def gbs(nbytes, ms_):
        return nbytes / (ms_ * 1e-3) / 1e9

    # Fused step: the dominant kernel IS the step -- yh_nms_kernel<.., TRAIN> does the train head's and the
    # post-process's work on an image it reads once; its launch duration is the stream-ordered one (every launch
    # starts after the one in front of it has completed: what ncu's serialised per-launch time corresponds to).
    # Algorithmic bytes per SURVEY 8(d): 3P + 48k + 28K' per image for train head + post-process (the fused
    # kernel's own minimum is 2P + 48k + 28K': it does not read y a second time -- both fractions are reported).
    # --unfused: the dominant kernel is yh_train_kernel (2P + 48k).  The step of the timed region (overlapped
    # launch chain) is reported under `step`.
    step_bytes = train_bytes + post_bytes
    fused_min = train_bytes + 28 * kept
    if fused:
        dom_ms, dom_bytes = step_iso_ms, step_bytes
        dom_name = ("yh_nms_kernel<IMG,TRAIN> = the fused step (one CTA per image: dense loss/gradient pass + records + "
                    "threshold + NMS + class pick, y read once) + its one-warp finalize dependent")
    else:
        dom_ms, dom_bytes = train_ms, train_bytes
        dom_name = "yh_train_kernel (fused decode+assign+loss+dL/dy) + its one-warp finalize dependent"
    achieved = gbs(dom_bytes, dom_ms)
    roofline = {"bound": "hbm", "kernel": dom_name,
                "regime": "stream-ordered launches over the rotating buffer sets (each launch starts after the previous "
                          "one has completed; only launch latency hidden); duration = CUDA-event time of the graph / launches",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "us_per_launch": dom_ms * 1e3,
                "separate_kernels": {"what": "the kernels of the two separate calls, stream-ordered launches (the drop-in "
                                             "get_loss / detect path) and back-to-back train-head launches with the overlap promise",
                                     "train_algorithmic_bytes": train_bytes,
                                     "train_us_per_launch": train_ms * 1e3, "train_frac": gbs(train_bytes, train_ms) / peak,
                                     "train_overlapped_us_per_launch": train_ov_ms * 1e3,
                                     "train_overlapped_frac": gbs(train_bytes, train_ov_ms) / peak,
                                     "post_us_per_launch": post_ms * 1e3, "post_algorithmic_bytes": post_bytes,
                                     "post_frac": gbs(post_bytes, post_ms) / peak},
                "step": {"what": "train head + post-process of the timed region (the north-star target: >= 0.60 of the HBM "
                                 "roofline on SURVEY 8(d)'s algorithmic bytes 3P + 48k + 28K' per image)",
                         "algorithmic_bytes": step_bytes, "us": ms / K * 1e3,
                         "achieved": gbs(step_bytes, ms / K), "frac": gbs(step_bytes, ms / K) / peak,
                         "stream_ordered_us": step_iso_ms * 1e3, "stream_ordered_frac": gbs(step_bytes, step_iso_ms) / peak}}
    if fused:
        roofline["fused_minimum_bytes_per_launch"] = fused_min
        roofline["frac_of_fused_minimum"] = gbs(fused_min, dom_ms) / peak
        roofline["step"]["fused_minimum_bytes"] = fused_min
        roofline["step"]["frac_of_fused_minimum"] = gbs(fused_min, ms / K) / peak
    if fused and args.workload == "headline" and B == WORKLOADS["headline"]["batch"]:
        roofline["traffic"], roofline["traffic_source"] = measured_traffic()
    else:  # (the capture under profiles/ is of the fused kernel on the headline batch)
        roofline["traffic"], roofline["traffic_source"] = None, "no ncu --set full capture of this workload's dominant kernel under profiles/"
    if fused:
        roofline["note"] = ("frac and step.frac count SURVEY 8(d)'s bytes for train head + post-process (3P per image); the fused "
                            "kernel reads y once (2P), so on that figure a step can exceed 1.0 of the HBM peak -- "
                            "frac_of_fused_minimum is the fraction on the bytes it really moves")

    # ---- e2e: host buffers in, host buffers out
    e2e = None
    if not args.no_e2e:
        ke = args.e2e_steps or max(100, min(K, 400))  # (its own step count, stated in the line: 20 steps = 8 ms of wall clock)
        pipe = HostHeadPipeline(B, case.s_h, case.s_w, case.a, case.c, img_hw=(case.height, case.width),
                                anchors=case.anchors, lambdas=lam, conf_thre=CONF_THRE, iou_thre=IOU_THRE,
                                max_out=MAX_OUT, max_boxes=m_local, depth=3, device=dev,
                                compute_dy=True, return_dy=args.e2e_return_dy)
        gt_h = targets.records_to_tensor(case.rec)
        off_h = torch.from_numpy(case.gt_off)
        for d in range(pipe.depth):  # fill every slot's pinned staging buffers once (the "producer")
            hy, hg, ho = pipe.pinned_inputs(d)
            hy.copy_(case.y)
            hg[:m_local].copy_(gt_h)
            ho.copy_(off_h)

        def run_e2e(n):
            last = None
            for i in range(n):
                if i >= pipe.depth:
                    last = pipe.result(t0 + i - pipe.depth)
                pipe.submit(None, gt_h, None, m_global=m_global, staged=True)
            for i in range(max(0, n - pipe.depth), n):
                last = pipe.result(t0 + i)
            return last

        t0 = pipe._ticket
        # warm-up: untimed repetitions until the rate is steady -- the first few hundred steps after the device-resident
        # part run up to 40 % slower (measured on fresh boxes: 0.59, 0.44, 0.41 / 0.49, 0.45, 0.42 ms per step over
        # consecutive repetitions of 100 steps; host-side caches and the PCIe link leaving its idle state), which a
        # short run would mistake for the pipeline's rate
        warm_ms, prev = [], None
        for rep_ in range(24):  # (at least eight untimed repetitions, then until two consecutive ones agree within 1.5 %: <= ~1.5 s)
            torch.cuda.synchronize()
            t0 = pipe._ticket
            tic = time.perf_counter()
            run_e2e(max(ke, pipe.depth))
            torch.cuda.synchronize()
            cur = time.perf_counter() - tic
            warm_ms.append(1e3 * cur / max(ke, pipe.depth))
            if rep_ >= 7 and prev is not None and abs(cur - prev) <= 0.015 * cur:
                break
            prev = cur
        # three repetitions of `ke` steps (wall clock between device syncs, max over ranks each); reported: the median
        e2e_runs = []
        for _ in range(3):
            barrier()
            t0 = pipe._ticket
            tic = time.perf_counter()
            res = run_e2e(ke)
            torch.cuda.synchronize()
            toc = time.perf_counter()
            e2e_s = toc - tic
            if dist is not None:
                t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            e2e_runs.append(e2e_s)
        e2e_s = sorted(e2e_runs)[1]
        # (the pipeline runs without the exchange: its loss is this rank's share, 1 / world of the reduced one --
        #  every rank holds the same shard here)
        want_loss = loss_value / world if xch is not None else loss_value
        assert abs(float(res["loss"]) - want_loss) <= 1e-5 * abs(want_loss), (float(res["loss"]), want_loss)
        e2e = {"value": images * ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(m_local),
               "d2h_bytes_per_step": pipe.d2h_bytes(), "steps": ke, "ms_per_step": 1e3 * e2e_s / ke,
               "repeats_ms_per_step": [1e3 * t_ / ke for t_ in e2e_runs], "warmup_repetitions_ms_per_step": warm_ms,
               "path": "HostHeadPipeline: pinned host y+GT -> H2D -> yh_v2_train_post (loss + dL/dy + kept boxes) -> "
                       "D2H of loss, terms and kept boxes%s; 3 slots; the head tensor goes up in two halves on two H2D streams, kernels and D2H on their own streams; "
                       "at least eight untimed repetitions and until two agree within 1.5 %%, then the median of three timed ones"
                       % (" and dL/dy" if args.e2e_return_dy else " (dL/dy is computed every step and stays on the device)")}

    collective = None
    if world > 1:
        collective = ("every step: the six loss sums of all %d ranks summed inside the step's last kernel over NVLink peer "
                      "memory (56-byte peer stores + sequence words, csrc/yh_finalize.cuh); terms/loss of the whole sharded "
                      "batch on every rank; no NCCL call in the timed region" % world) if xch is not None else \
                     ("none in the timed region (%s): every rank reports the partial terms of its shard"
                      % ("--no-collective" if args.no_collective else "peer exchange unavailable: %s" % xch_error))

    # ---- CPU baseline (rank 0, N=1 only): a bounded amount of the same CPU work as --impl reference
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = cpu_sample_size(args.workload, B, args.cpu_sample)
        ts, kind, what = time_cpu(args.workload, sample, reps=3, warm=1)
        cpu = {"value": sample / float(np.min(ts)), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": "%d images per step (%s), best of 3 after 1 warm-up (%.1f s of CPU work); %s; os.cpu_count()=%s"
                         % (sample, "the full batch" if sample == B else "a sample of the batch-%d workload" % B,
                            float(np.sum(ts)), what, os.cpu_count())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "repeats": n_regions,
            "region_ms": {"median": ms, "min": region_ms[0], "max": region_ms[-1],
                          "what": "device time of the K-step region, max over ranks, per repetition"},
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.workload, B), **({"global_batch": B * world} if args.strong else {})),
            "run": run_description(R, set_bytes + (sets[0]["res"].get("_ws").numel() if fused else 0), fused, collective),
            "clocks": sampler.summary(window), "e2e": e2e, "gpu_launches": (2 if fused else 3) * K,
            "roofline": roofline, "cpu_baseline": cpu, "loss": loss_value, "terms": terms_value,
        }
        if world > 1:
            line["collective_in_timed_region"] = xch is not None
            if ms_nc is not None:
                line["without_collective"] = {"ms_per_step": ms_nc / K, "value": images * K / (ms_nc * 1e-3)}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
