"""BASELINE.json config 4: a full YOLOv2-shaped training step, batch-sharded across the ranks of torchrun.

    conv backbone (stock cuDNN)  ->  head tensor  ->  fused train head (libyolohead)  ->  loss.backward()
    ->  DDP gradient all-reduce over NCCL  ->  fused SGD step (libyolohead)

The backbone is OUT OF SCOPE of this package (north-star: it stays on stock cuDNN); what stands in for it here is
a Darknet-19-shaped stack of torch.nn layers with random weights, only so that the head path can be shown and
timed inside a whole step.  Each rank takes a contiguous image shard (odcp_b200.dist), passes the ALL-RANK box
count as m_global and scales its loss by the world size, so that DDP's gradient average equals the gradient of
the unsharded batch.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scratch/cfg4_train_step.py [--check]

--check   parity instead of timing: after one step from identical weights on a small BN-free backbone, the
          parameters of the sharded run must equal those of a single-process run over the whole batch.
Prints one JSON line on rank 0.  Not part of the product."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

from odcp_b200 import dist as yh_dist, synthetic, targets
from odcp_b200.models.layout import head_tensor, use_channels_last_head
from odcp_b200.models.yolov2 import YOLOv2HeadOps
from odcp_b200.optim import SGD, reset_state

DARKNET19 = [(32, 3), "M", (64, 3), "M", (128, 3), (64, 1), (128, 3), "M", (256, 3), (128, 1), (256, 3), "M",
             (512, 3), (256, 1), (512, 3), (256, 1), (512, 3), "M", (1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3),
             (1024, 3), (1024, 3)]
SMALL = [(8, 3), "M", (16, 3), "M", (16, 3), "M", (32, 3), "M", (32, 3), "M"]


class StandInYOLOv2(YOLOv2HeadOps, torch.nn.Module):
    """Conv stack + 1x1 head conv -> [N,S,S,A,5+C]; predict/get_loss/detect come from the CUDA head path."""

    def __init__(self, cfg, bn=True, num_cls=20):
        torch.nn.Module.__init__(self)
        layers, cin = [], 3
        for item in cfg:
            if item == "M":
                layers.append(torch.nn.MaxPool2d(2, 2))
                continue
            cout, k = item
            layers.append(torch.nn.Conv2d(cin, cout, k, padding=k // 2, bias=not bn))
            if bn:
                layers.append(torch.nn.BatchNorm2d(cout))
            layers.append(torch.nn.LeakyReLU(0.1))
            cin = cout
        self.cls_list = [str(i) for i in range(num_cls)]
        self.num_cls = num_cls
        self.anchor_box_size_list = [tuple(a) for a in synthetic.YOLOV2_ANCHORS]
        self.num_anchor_box = 5
        layers.append(torch.nn.Conv2d(cin, 5 * (5 + num_cls), 1))
        self.net = torch.nn.Sequential(*layers)

    def forward(self, x_batch):  # x_batch [N,H,W,3] like the reference
        out = self.net(x_batch.permute(0, 3, 1, 2))
        return head_tensor(out, self.num_anchor_box)


class Step(torch.nn.Module):
    """forward = loss, so that DDP wraps the whole thing and hooks the backward."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x, gt, gt_off, m_global, scale):
        return self.model.get_loss_compact(x, gt, gt_off, m_global=m_global) * scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lr, mom, wd = 1e-3, 0.9, 5e-4

    if args.check:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        n_total = 4 * world
        case = synthetic.make_case("cfg4_check", 2, n_total, 13, 13, 5, 20, 416, 416, seed=104)
        gen = torch.Generator().manual_seed(7)
        x_all = torch.rand(n_total, 416, 416, 3, generator=gen)
        torch.manual_seed(11)
        model = StandInYOLOv2(SMALL, bn=False).to(dev)
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        # sharded step
        rec, off, (lo, hi) = yh_dist.shard_case(case.rec, case.gt_off, n_total, rank, world)
        m_global = yh_dist.global_box_count(len(rec), device=dev)
        step = Step(model)
        wrapped = DDP(step, device_ids=[local]) if world > 1 else step
        opt = SGD(model.parameters(), lr=lr, momentum=mom, weight_decay=wd)
        opt.zero_grad()
        loss = wrapped(x_all[lo:hi].to(dev), targets.records_to_tensor(rec, dev), torch.from_numpy(off).to(dev),
                       m_global, yh_dist.ddp_gradient_scale(world))
        loss.backward()
        opt.step()
        terms, total = yh_dist.reduce_terms(model._yh_last["terms"], loss.detach() / world)
        sharded = {k: v.clone() for k, v in model.state_dict().items()}
        # whole batch, one process, same starting weights (every rank does it: same result everywhere)
        model.load_state_dict(state0)
        reset_state()
        opt = SGD(model.parameters(), lr=lr, momentum=mom, weight_decay=wd)
        opt.zero_grad()
        loss1 = model.get_loss_compact(x_all.to(dev), targets.records_to_tensor(case.rec, dev),
                                       torch.from_numpy(case.gt_off).to(dev))
        loss1.backward()
        opt.step()
        worst = 0.0
        for k, v in model.state_dict().items():
            d = (sharded[k] - v).abs().max().item()
            upd = (v - state0[k]).abs().max().item()
            worst = max(worst, d / max(upd, 1e-12))
        ok = worst < 1e-3 and abs(total.item() - loss1.item()) <= 1e-5 * abs(loss1.item())
        if rank == 0:
            print(json.dumps(dict(check="cfg4 sharded step == whole-batch step", world=world, m_global=m_global,
                                  loss_sharded=total.item(), loss_whole=loss1.item(),
                                  worst_param_diff_over_update=worst, ok=bool(ok))), flush=True)
        if world > 1:
            dist.destroy_process_group()
        sys.exit(0 if ok else 1)

    # ---- timing: Darknet-19-shaped backbone, channels_last, TF32 as torch defaults
    b = args.batch
    case = synthetic.make_case("cfg4", 2, b, 13, 13, 5, 20, 416, 416, seed=104 + rank)
    torch.manual_seed(11)
    model = use_channels_last_head(StandInYOLOv2(DARKNET19, bn=True)).to(dev)
    step = Step(model)
    wrapped = DDP(step, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else step
    x = torch.rand(b, 416, 416, 3, device=dev)
    gt, off = targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
    m_global = yh_dist.global_box_count(case.m, device=dev)
    scale = yh_dist.ddp_gradient_scale(world)

    def one(optimizer_cls):
        opt = optimizer_cls(model.parameters(), lr=lr, momentum=mom, weight_decay=wd)  # re-created per iteration
        opt.zero_grad()
        loss = wrapped(x, gt, off, m_global, scale)
        loss.backward()
        opt.step()
        return loss

    def time_steps(optimizer_cls):
        for _ in range(3):
            one(optimizer_cls)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            loss = one(optimizer_cls)
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(loss.item()) / scale

    ms_fused, loss = time_steps(SGD)
    ms_torch, _ = time_steps(torch.optim.SGD)
    # the head alone on this rank's head tensor
    with torch.no_grad():
        y = model(x).contiguous()
    from odcp_b200 import ops
    kw = dict(version=2, img_hw=(416, 416), anchors=synthetic.YOLOV2_ANCHORS, lambdas=synthetic.DEFAULT_LAMBDAS, m_global=m_global)
    outb = ops.train_head(y, gt, off, **kw)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        ops.train_head(y, gt, off, out=outb, **kw)
    e.record()
    torch.cuda.synchronize()
    head_us = a.elapsed_time(e) * 1e3 / 50
    if rank == 0:
        n_par = sum(p.numel() for p in model.parameters())
        print(json.dumps(dict(config="cfg4 full step: Darknet-19-shaped cuDNN stand-in + fused head + DDP + fused SGD",
                              n_gpus=world, batch_per_gpu=b, parameters=n_par, loss=loss,
                              ms_per_step_fused_sgd=round(ms_fused, 3), ms_per_step_torch_sgd=round(ms_torch, 3),
                              images_per_s=round(b * world / ms_fused * 1e3, 1),
                              head_us_per_call_incl_python=round(head_us, 1),
                              head_share_of_step=round(head_us / 1e3 / ms_fused, 5))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
