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
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--sets", type=int, default=8, help="rotating buffer sets (total must exceed L2)")
    ap.add_argument("--cpu-sample", type=int, default=32, help="images in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--two-streams", action="store_true",
                    help="issue the post-process of a step on a second stream (fork/join inside the step). "
                         "Default: one stream; the post-process is launched as a programmatic dependent with "
                         "YH_POST_INPUT_READY, so it already overlaps the train head's tail")
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
    return ap.parse_args()


def workload_config(batch, sets=None, serial=False):
    cfg = {
        "workload": "yolov2_head_13x13x5_c20_b%d_train(decode+assign+loss+bwd)+postprocess(conf%.2f,nms_iou%.2f)"
                    % (batch, CONF_THRE, IOU_THRE),
        "batch_per_gpu": batch, "grid": [13, 13], "anchors": 5, "classes": 20, "image": [416, 416],
        "gt_boxes_per_image": "U{1..5}", "candidates_per_image": "~50 of 845",
    }
    if sets is not None:
        cfg["l2"] = "%d rotating input/output buffer sets per GPU (inputs+outputs %.0f MB > 126 MB L2)" % (
            sets, sets * 2 * batch * 84500 / 1e6)
        if serial:
            cfg["streams"] = ("one stream, programmatic dependent launches; a graph replay is 4 passes over the %d buffer "
                              "sets; every kernel but the first of a pass promises that its buffers are not in use by the "
                              "kernels in front of it (rotating sets) and overlaps their tails; the first kernel of every "
                              "pass waits for everything before it" % (sets or 0))
        else:
            cfg["streams"] = "train head and post-process of a step on two streams (parallel graph branches), steps in order"
    return cfg


# ------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/) -- the baseline arm
# ------------------------------------------------------------------------------------------
def cpu_port_step(case, lambdas):
    """One pass of the reference's torch-CPU algorithm over `case`: get_loss + backward
    (dense per-box replication, autograd) and the per-image detect/NMS loop."""
    from oracle import yolo_head_oracle as O
    O.train_head_dense(case, lambdas)
    O.postprocess_torch(case.y, case.height, case.width, case.version, case.anchors, CONF_THRE, IOU_THRE)


def time_cpu_port(sample, reps, warm):
    from odcp_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    case = synthetic.headline(n=sample)
    for _ in range(warm):
        cpu_port_step(case, synthetic.DEFAULT_LAMBDAS)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_port_step(case, synthetic.DEFAULT_LAMBDAS)
        ts.append(time.perf_counter() - t0)
    return case, ts


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference does
    not exist on the GPU box and the reference is torch-eager Python (nothing to compile), so this
    arm times the oracle's torch-CPU port of it (same ops in the same order, pinned bit-exact
    against the reference's outputs in tests/golden/) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    case, ts = time_cpu_port(sample, steps, warm)
    total = float(np.sum(ts))
    value = sample * len(ts) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(ts), "warmup": warm, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d-image sample of the batch-%d workload per step (the reference's per-box "
                                   "replication grows as M*S*S*A); torch %s CPU, %d threads, os.cpu_count()=%s"
                                   % (sample, args.batch, torch.__version__, cores, os.cpu_count())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    K, W, R, B = args.steps, max(args.warmup, 3), args.sets, args.batch
    lam = synthetic.DEFAULT_LAMBDAS
    case = synthetic.headline(n=B)
    kw = dict(version=2, img_hw=(case.height, case.width), anchors=case.anchors)
    m_local = case.m
    m_global = m_local * world  # every rank holds a shard with the same box count (weak scaling)

    # rotating buffer sets: distinct addresses, > L2 in total
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    y0 = case.y.to(dev)
    sets = []
    for i in range(R):
        s = dict(y=y0.clone(), gt=gt.clone(), off=off.clone(),
                 out=dict(dy=torch.empty_like(y0), loss=torch.empty((), device=dev), terms=torch.empty(5, device=dev)))
        sets.append(s)

    stream = torch.cuda.Stream(dev)
    side = torch.cuda.Stream(dev)
    fork_ev = [torch.cuda.Event() for _ in range(2)]

    # Overlap of consecutive kernels (programmatic dependent launch).  Every kernel of this path is
    # launched as a programmatic dependent, which hides launch latency.  On top of that a call may
    # promise that its buffers are not touched by the kernels still draining in front of it
    # (`input_ready`): the kernel then runs next to their tails and only waits for them before it
    # completes (yh_v2_train_overlapped, YH_POST_INPUT_READY).  The bench rotates R buffer sets, so the
    # promise holds for every step of a graph replay except that a set comes round again in the NEXT
    # replay: the first kernel of every captured graph is therefore launched without the promise -- it
    # waits at its start for everything before it -- which bounds the overlap chain to one replay.
    def run_post(s, ready=True):
        s["post"] = ops.postprocess(s["y"], conf_thre=CONF_THRE, iou_thre=IOU_THRE, max_out=MAX_OUT,
                                    want_cls_spec=False, out=s.get("post"), input_ready=ready, **kw)

    def step(s, post=True, train=True, first=True, overlap=True):
        """One step = one train-head call + one post-process call on the same head tensor, in stream
        order (with --two-streams: fork/join inside the step, parallel branches of the captured graph).
        `first`: first step of a captured graph (its first kernel makes no promise)."""
        both = post and train and args.two_streams
        if both:
            fork_ev[0].record(stream)
            side.wait_event(fork_ev[0])
            with torch.cuda.stream(side):
                run_post(s, ready=False)
                fork_ev[1].record(side)
        if train:
            ops.train_head(s["y"], s["gt"], s["off"], lambdas=lam, m_global=m_global, out=s["out"],
                           input_ready=overlap and not first and not args.two_streams, **kw)
        if both:
            stream.wait_event(fork_ev[1])
        elif post:
            run_post(s, ready=overlap and not args.two_streams and (train or not first))

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        return g

    with torch.cuda.stream(stream):
        for s in sets:  # eager warm-up: allocates outputs/workspaces before any capture
            step(s)
        stream.synchronize()
        # one graph = ROUNDS passes over the R buffer sets; the first kernel of every pass waits for everything
        # before it (no promise), so the overlap chain stays bounded by the rotation length
        ROUNDS = 4
        G = R * ROUNDS
        g_full = capture(lambda: [step(s, first=(i == 0)) for _ in range(ROUNDS) for i, s in enumerate(sets)])
        # step counts that are not a multiple of G: one more graph of exactly the remaining steps (same chain)
        g_rem = {}
        for rem in {K % G, W % G} - {0}:
            g_rem[rem] = capture(lambda rem=rem: [step(sets[j % R], first=(j % R == 0)) for j in range(rem)])
        # per-kernel graphs: stream-ordered launches (only launch latency hidden) for the roofline of one
        # launch, and the overlapped variant for the sustained rate of back-to-back launches
        g_train = capture(lambda: [step(s, post=False, overlap=False) for s in sets])
        g_post = capture(lambda: [step(s, train=False, overlap=False) for s in sets])
        g_train_ov = capture(lambda: [step(s, post=False, first=(i == 0)) for i, s in enumerate(sets)])
        g_post_ov = capture(lambda: [step(s, train=False, first=(i == 0)) for i, s in enumerate(sets)])

        def run_steps(n):
            for _ in range(n // G):
                g_full.replay()
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

        def timed_region(ev0, ev1):
            """Exactly K steps between two events.  A gate kernel keeps the GPU busy while the host
            enqueues ev0, the graph launches and ev1: the region then starts with its first kernel already
            queued behind the event (no idle-GPU launch latency inside it)."""
            if gate_cycles > 0:
                torch.cuda._sleep(gate_cycles)
            ev0.record(stream)
            run_steps(K)
            ev1.record(stream)

        for _ in range(2):
            timed_region(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            stream.synchronize()

        # ---- timed: `reps` regions of exactly K steps each, device time, barrier + synchronize on both
        # sides of every region; per region the MAX over ranks, reported: the median region (min/max next to it)
        est_ms = max(K * 0.014, 1e-3)
        reps = args.repeats or int(min(15, max(3, round(25.0 / est_ms))))
        sampler = ClockSampler(nvml_index(local_rank))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        barrier()
        torch.cuda.synchronize()
        sampler.start()
        for ev0, ev1 in evs:
            barrier()
            torch.cuda.synchronize()
            timed_region(ev0, ev1)
            torch.cuda.synchronize()
            barrier()
        region_ms = torch.tensor([a.elapsed_time(b) for a, b in evs], device=dev, dtype=torch.float64)
        window = "the %d timed regions" % reps
        if len(sampler.samples) < 5:  # the timed regions were too short to sample: keep the same load on
            t_end = time.perf_counter() + 0.25
            while time.perf_counter() < t_end:
                g_full.replay()
            stream.synchronize()
            window = "the %d timed regions + 0.25 s of the same graph replayed right after them" % reps
        sampler.stop()
        if dist is not None:
            dist.all_reduce(region_ms, op=dist.ReduceOp.MAX)
        region_ms = sorted(float(v) for v in region_ms.tolist())
        ms = region_ms[len(region_ms) // 2]
        loss_value = float(sets[0]["out"]["loss"].item())

        # ---- per-kernel durations (train head alone / post-process alone), same rotation
        def time_graph(g, reps):
            for _ in range(3):
                g.replay()
            stream.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        post_ov_ms = time_graph(g_post_ov, reps)

    images = B * world
    value = images * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (fused train head)
    kept = sets[0]["post"]["keep_cnt"].clamp(max=MAX_OUT).sum().item()
    p_bytes = case.image_bytes
    train_bytes = 2 * p_bytes * B + 48 * m_local
    post_bytes = p_bytes * B + 28 * kept
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    # The timed region runs the kernels as an overlapped launch chain, so the dominant kernel's average launch
    # duration in that regime is the issue interval of back-to-back launches (graph time / launches).  The
    # isolated figure (every launch waits for the one in front of it to complete) is reported next to it.
    def gbs(nbytes, ms_):
        return nbytes / (ms_ * 1e-3) / 1e9

    # The dominant kernel's launch duration is the stream-ordered one (every launch starts after the one in
    # front of it has completed: what a drop-in caller behind a conv gets, and what ncu's serialised per-launch
    # time corresponds to).  The issue interval of back-to-back launches with the overlap promise -- the regime
    # of the timed region -- is reported under `sustained`, the whole step of the timed region under `step`.
    achieved = gbs(train_bytes, train_ms)
    roofline = {"bound": "hbm",
                "kernel": "yh_train_kernel (fused decode+assign+loss+dL/dy) + its one-warp finalize dependent",
                "regime": "stream-ordered launches over the rotating buffer sets (each launch starts after the previous "
                          "one has completed; only launch latency hidden); duration = CUDA-event time of the graph / launches",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": train_bytes,
                "us_per_launch": train_ms * 1e3,
                "sustained": {"what": "back-to-back launches with the overlap promise (yh_v2_train_overlapped / "
                                      "YH_POST_INPUT_READY), as in the timed region: issue interval = graph time / launches",
                              "train_us_per_launch": train_ov_ms * 1e3, "train_achieved": gbs(train_bytes, train_ov_ms),
                              "train_frac": gbs(train_bytes, train_ov_ms) / peak,
                              "post_us_per_launch": post_ov_ms * 1e3, "post_frac": gbs(post_bytes, post_ov_ms) / peak},
                "postprocess_kernel": {"us_per_launch": post_ms * 1e3, "algorithmic_bytes_per_launch": post_bytes,
                                       "achieved": gbs(post_bytes, post_ms), "frac": gbs(post_bytes, post_ms) / peak,
                                       "regime": "stream-ordered launches"},
                "step": {"what": "train head + post-process of the timed region (the north-star target: >= 0.60)",
                         "algorithmic_bytes": train_bytes + post_bytes, "us": ms / K * 1e3,
                         "achieved": gbs(train_bytes + post_bytes, ms / K),
                         "frac": gbs(train_bytes + post_bytes, ms / K) / peak}}
    roofline["traffic"], roofline["traffic_source"] = measured_traffic()

    # ---- e2e: host buffers in, host buffers out
    e2e = None
    if not args.no_e2e:
        ke = args.e2e_steps or max(20, min(K, 400))
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
        run_e2e(max(3, pipe.depth))
        torch.cuda.synchronize()
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
        assert abs(float(res["loss"]) - loss_value) <= 1e-6 * abs(loss_value), (float(res["loss"]), loss_value)
        e2e = {"value": images * ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(m_local),
               "d2h_bytes_per_step": pipe.d2h_bytes(), "steps": ke, "ms_per_step": 1e3 * e2e_s / ke,
               "path": "HostHeadPipeline: pinned host y+GT -> H2D -> yh_v2_train (loss + dL/dy) + yh_v2_postprocess -> "
                       "D2H of loss, terms and kept boxes%s; 3 slots, copy/compute streams overlapped"
                       % (" and dL/dy" if args.e2e_return_dy else " (dL/dy is computed every step and stays on the device)")}

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample
        _, ts = time_cpu_port(sample, reps=5, warm=1)
        cpu = {"value": sample / float(np.min(ts)), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d-image sample of the same workload, best of 5 after 1 warm-up (%.2f s of CPU work); "
                         "oracle/ torch-CPU port of get_loss+backward and the per-image nms loop; "
                         "os.cpu_count()=%s" % (sample, float(np.sum(ts)), os.cpu_count())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "repeats": reps,
            "region_ms": {"median": ms, "min": region_ms[0], "max": region_ms[-1],
                          "what": "device time of the K-step region, max over ranks, per repetition"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(B, R, not args.two_streams),
            "clocks": sampler.summary(window), "e2e": e2e, "gpu_launches": 3 * K,
            "roofline": roofline, "cpu_baseline": cpu, "loss": loss_value,
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
