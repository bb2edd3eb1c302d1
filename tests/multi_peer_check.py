"""Run under torchrun (one rank per GPU): the batch-sharded fused step with the in-kernel peer-memory reduction
of the loss terms (odcp_b200.dist.PeerExchange, csrc/yh_finalize.cuh) against the whole batch on one GPU.

    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 tests/multi_peer_check.py

Checks, on every rank: reduced loss / terms == whole-batch loss / terms (1e-6 relative; the sums are exact
integers, only the per-CTA float partials differ with the tiling), identical BITS on all ranks, the shard's dL/dy
and kept boxes equal to the whole batch's slice bit for bit, through the stream-ordered call, the train-only
sharded call (finalize kernel) and an overlapped chain replayed from a CUDA graph.  Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from odcp_b200 import dist as yh_dist, ops, synthetic, targets

LAM = synthetic.DEFAULT_LAMBDAS


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    xch = yh_dist.PeerExchange(device=dev)
    fails = []

    def check(cond, what):
        if not cond:
            fails.append(what)

    def close(a, b, tol=1e-6):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return bool(np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1e-30)))

    R = 4
    cases = [synthetic.make_case("mp%d" % i, 2, 64 * world + 1 + i, 13, 13, 5, 20, 416, 416, seed=500 + i, to_shift=-1.563)
             for i in range(R)]
    kw = lambda c: dict(img_hw=(c.height, c.width), lambdas=LAM, anchors=c.anchors)  # noqa: E731
    post = dict(conf_thre=0.5, iou_thre=0.45, max_out=128, want_cls_spec=False)
    whole, shard_in, lohi = [], [], []
    for c in cases:
        y = c.y.to(dev)
        gt, off = targets.records_to_tensor(c.rec, dev), torch.from_numpy(c.gt_off).to(dev)
        whole.append(ops.train_post(y, gt, off, **kw(c), **post))
        rec, soff, (lo, hi) = yh_dist.shard_case(c.rec, c.gt_off, c.n, rank, world)
        shard_in.append((y[lo:hi].contiguous(), targets.records_to_tensor(rec, dev), torch.from_numpy(soff).to(dev)))
        lohi.append((lo, hi))
    torch.cuda.synchronize()

    def compare(i, r, tag, with_post=True):
        c, w, (lo, hi) = cases[i], whole[i], lohi[i]
        check(close(r["train"]["loss"].item(), w["train"]["loss"].item()), "%s loss case %d: %r vs %r" % (
            tag, i, r["train"]["loss"].item(), w["train"]["loss"].item()))
        check(close(r["train"]["terms"].cpu().numpy(), w["train"]["terms"].cpu().numpy()), "%s terms case %d" % (tag, i))
        check(torch.equal(r["train"]["dy"], w["train"]["dy"][lo:hi]), "%s dy case %d" % (tag, i))
        bits = r["train"]["loss"].view(torch.int32).reshape(1).clone()
        allb = [torch.zeros_like(bits) for _ in range(world)]
        dist.all_gather(allb, bits)
        check(all(int(b.item()) == int(allb[0].item()) for b in allb), "%s loss bits differ between ranks, case %d" % (tag, i))
        if with_post:
            check(torch.equal(r["post"]["keep_cnt"], w["post"]["keep_cnt"][lo:hi]), "%s keep_cnt case %d" % (tag, i))
            cnt = r["post"]["keep_cnt"].cpu().numpy()
            valid = torch.from_numpy(np.arange(128)[None, :] < np.minimum(cnt, 128)[:, None]).to(dev)
            check(torch.equal(r["post"]["keep_idx"][valid], w["post"]["keep_idx"][lo:hi][valid]), "%s keep_idx case %d" % (tag, i))

    # 1. stream-ordered fused step with the exchange
    for i, c in enumerate(cases):
        y, gt, off = shard_in[i]
        r = ops.train_post(y, gt, off, m_global=c.m, exchange=xch, **kw(c), **post)
        torch.cuda.synchronize()
        compare(i, r, "fused")
    # 2. train-only sharded call (the finalize kernel does the exchange)
    for i, c in enumerate(cases):
        y, gt, off = shard_in[i]
        t = ops.train_head(y, gt, off, version=2, m_global=c.m, exchange=xch, **kw(c))
        torch.cuda.synchronize()
        compare(i, dict(train=t), "train_sharded", with_post=False)
    # 3. overlapped chain from a CUDA graph, replayed
    outs = [None] * R
    stream = torch.cuda.Stream(dev)

    def chain():
        for i, c in enumerate(cases):
            y, gt, off = shard_in[i]
            outs[i] = ops.train_post(y, gt, off, m_global=c.m, exchange=xch, out=outs[i], overlapped=i > 0, **kw(c), **post)

    with torch.cuda.stream(stream):
        chain()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            chain()
        for _ in range(3):
            for o in outs:
                o["train"]["loss"].fill_(-1.0)
                o["train"]["dy"].fill_(float("nan"))
            g.replay()
            g.replay()
            stream.synchronize()
            for i in range(R):
                compare(i, outs[i], "graph chain")
    ok = torch.tensor([0 if fails else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if fails:
        print("rank %d FAILED: %s" % (rank, fails[:6]), file=sys.stderr, flush=True)
    if rank == 0:
        print(json.dumps({"ok": bool(ok.item()), "world": world, "cases": R, "loss": float(whole[0]["train"]["loss"].item())}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
