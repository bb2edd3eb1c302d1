"""BASELINE config 4: the full YOLOv2 training step, batch-sharded across the ranks of torchrun.

    conv backbone + neck + head convs (stock cuDNN)  ->  head tensor  ->  fused train head (libyolohead)
    ->  loss.backward()  ->  DDP gradient all-reduce over NCCL  ->  fused SGD step (libyolohead)

What is product here is the head path, the optimizer step and the sharding rules; the convolutions are OUT OF SCOPE
(north star: they stay on stock cuDNN).  `YOLOv2Net` exists so that the whole step can be run, checked and timed
without the reference on the path: the published Darknet-19 / YOLOv2 layer table in plain `torch.nn`, with the
reference's module names (reference models/backbones/darknet19.py:15-221, models/yolov2.py:72-89), so it has the
reference's 67 147 837 parameters, its `state_dict` keys (a reference checkpoint loads as it is) and its forward
(reference models/yolov2.py:91-431: normalize, net1..net7, the pass-through reorg of the 26x26 map, the two head
convs, permute + reshape) -- and the CUDA head path for predict / get_loss / detect (models/yolov2.py mixin).

Sharding: each rank takes a contiguous image shard, passes the ALL-RANK box count as `m_global` and
scales its loss by the world size, so that DDP's gradient average equals the gradient of the unsharded batch.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m odcp_b200.train_step --check
    python bench.py --workload cfg4 [--gpus N]        # the timed form
"""
from __future__ import annotations

import json
import os
import sys

import torch
from torch import nn

from . import dist as yh_dist, synthetic, targets
from .models.layout import head_tensor, use_channels_last_head
from .models.yolov2 import YOLOv2HeadOps
from .optim import SGD, reset_state

# (out_channels, kernel) per conv of net1..net7; every net but the first and the last starts with a 2x2 max-pool
DARKNET19_NETS = (
    ((32, 3),),
    ((64, 3),),
    ((128, 3), (64, 1), (128, 3)),
    ((256, 3), (128, 1), (256, 3)),
    ((512, 3), (256, 1), (512, 3), (256, 1), (512, 3)),
    ((1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3)),
    ((1024, 3), (1024, 3)),
)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
YOLOV2_PARAMETERS = 67_147_837


def _conv_block(cin, cout, k, bn):
    layers = [nn.Conv2d(cin, cout, k, padding=k // 2)]  # (bias kept, as the reference's layers have it)
    if bn:
        layers.append(nn.BatchNorm2d(cout))
    layers.append(nn.LeakyReLU(0.1))
    return layers


class Darknet19Backbone(nn.Module):
    def __init__(self, nets=DARKNET19_NETS, bn=True):
        super().__init__()
        cin = 3
        for i, convs in enumerate(nets):
            layers = [nn.MaxPool2d(2, 2)] if 0 < i < len(nets) - 1 else []
            for cout, k in convs:
                layers += _conv_block(cin, cout, k, bn)
                cin = cout
            setattr(self, "net%d" % (i + 1), nn.Sequential(*layers))
        self.num_nets = len(nets)
        self.out_channels = cin
        self.register_buffer("_mean", torch.tensor(IMAGENET_MEAN), persistent=False)
        self.register_buffer("_std", torch.tensor(IMAGENET_STD), persistent=False)

    def normalize(self, x):
        """[N,H,W,3] in 0..255 -> normalised [N,3,H,W] (reference darknet19.py:262-280)."""
        return ((x / 255 - self._mean) / self._std).permute(0, 3, 1, 2)


class YOLOv2Net(YOLOv2HeadOps, nn.Module):
    """The YOLOv2 network in stock torch.nn layers with the CUDA head path; `nets` / `bn` / `head_mid` shrink it for
    the parity check (cuDNN's backward is only reproducible enough for a 1e-3 comparison on a small BN-free stack)."""

    def __init__(self, cls_list=None, cls2idx=None, *, nets=DARKNET19_NETS, bn=True, head_mid=1024):
        nn.Module.__init__(self)
        self.cls_list = list(cls_list) if cls_list is not None else [str(i) for i in range(20)]
        self.cls2idx = cls2idx if cls2idx is not None else {c: i for i, c in enumerate(self.cls_list)}
        self.num_cls = len(self.cls_list)
        self.anchor_box_size_list = [tuple(a) for a in synthetic.YOLOV2_ANCHORS]
        self.num_anchor_box = len(self.anchor_box_size_list)
        self.head_output_dim = self.num_anchor_box * (5 + self.num_cls)
        self.backbone_model = Darknet19Backbone(nets, bn)
        c1 = nets[-3][-1][0]                      # channels of the 2x finer map the neck folds in (512)
        cin = 4 * c1 + self.backbone_model.out_channels
        self.head_model = nn.Sequential(*_conv_block(cin, head_mid, 3, bn), nn.Conv2d(head_mid, self.head_output_dim, 1))

    def forward(self, x_batch):
        b = self.backbone_model
        h = b.normalize(x_batch)
        for i in range(1, b.num_nets - 1):
            h = getattr(b, "net%d" % i)(h)
        h1 = h                                                        # [N, 512, 2S, 2S]
        h2 = getattr(b, "net%d" % b.num_nets)(getattr(b, "net%d" % (b.num_nets - 1))(h1))  # [N, 1024, S, S]
        # pass-through: the finer map's two halves in width, then in height, stacked on the channels
        # (reference models/yolov2.py:192-312)
        s_h, s_w = h2.shape[2], h2.shape[3]
        h1 = torch.cat([h1[:, :, :, i * s_w:(i + 1) * s_w] for i in range(2)], dim=1)
        h1 = torch.cat([h1[:, :, i * s_h:(i + 1) * s_h, :] for i in range(2)], dim=1)
        return head_tensor(self.head_model(torch.cat([h1, h2], dim=1)), self.num_anchor_box)


SMALL_NETS = (((8, 3),), ((16, 3),), ((16, 3),), ((32, 3),), ((32, 3),), ((32, 3),), ((32, 3),))


class LossStep(nn.Module):
    """forward = the scaled loss, so that DDP wraps the whole model and hooks its backward."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x, gt, gt_off, m_global, scale):
        return self.model.get_loss_compact(x, gt, gt_off, m_global=m_global) * scale


class ShardedTrainStep:
    """One rank's side of the sharded training step.  `step(x, gt, gt_off, m_local)` runs what the reference's
    run_one_epoch does per batch (models/yolov2.py:1237-1272: get_loss, a NEW SGD, zero_grad, backward, step) on this
    rank's shard; the parameter gradients are averaged by DDP's bucketed NCCL all-reduce during the backward."""

    def __init__(self, model, lr=1e-3, momentum=0.9, weight_decay=5e-4, optimizer_cls=SGD, bucket_cap_mb=None,
                 static_box_count=None):
        import torch.distributed as dist
        self.model, self.hyper = model, dict(lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.optimizer_cls = optimizer_cls
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.inner = LossStep(model)
        if self.world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            kw = dict(device_ids=[torch.cuda.current_device()], gradient_as_bucket_view=True)
            if bucket_cap_mb:
                kw["bucket_cap_mb"] = bucket_cap_mb
            self.wrapped = DDP(self.inner, **kw)
        else:
            self.wrapped = self.inner
        self.scale = yh_dist.ddp_gradient_scale(self.world)
        self.static_box_count = static_box_count  # all-rank box count known from the sampler: no all-reduce per step
        self.params = list(model.parameters())

    def allreduce_bytes(self):
        return 4 * sum(p.numel() for p in self.params)

    def step(self, x, gt, gt_off, m_local):
        m_global = self.static_box_count or yh_dist.global_box_count(m_local, device=x.device)
        opt = self.optimizer_cls(self.params, **self.hyper)  # re-created per iteration, as the reference does
        opt.zero_grad()
        loss = self.wrapped(x, gt, gt_off, m_global, self.scale)
        loss.backward()
        opt.step()
        return loss.detach() / self.scale  # this rank's share of the whole batch's loss (the terms add up over ranks)


def check(world, rank, dev):
    """Parity: after one step from identical weights on a small BN-free net, the parameters of the sharded run equal
    those of a single-process step over the whole batch, and the reduced loss the whole batch's."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n_total = 4 * world
    case = synthetic.make_case("cfg4_check", 2, n_total, 13, 13, 5, 20, 416, 416, seed=104)
    x_all = torch.rand(n_total, 416, 416, 3, generator=torch.Generator().manual_seed(7)) * 255.0
    torch.manual_seed(11)
    model = YOLOv2Net(nets=SMALL_NETS, bn=False, head_mid=32).to(dev)
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    rec, off, (lo, hi) = yh_dist.shard_case(case.rec, case.gt_off, n_total, rank, world)
    st = ShardedTrainStep(model)
    share = st.step(x_all[lo:hi].to(dev), targets.records_to_tensor(rec, dev), torch.from_numpy(off).to(dev), len(rec))
    _, total = yh_dist.reduce_terms(model._yh_last["terms"], share)
    sharded = {k: v.clone() for k, v in model.state_dict().items()}
    # whole batch, one process, same starting weights (every rank does it: same result everywhere)
    model.load_state_dict(state0)
    reset_state()
    opt = SGD(model.parameters(), **st.hyper)
    opt.zero_grad()
    loss1 = model.get_loss_compact(x_all.to(dev), targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev))
    loss1.backward()
    opt.step()
    # (one step moves a parameter by ~1e-6 while its own fp32 spacing is ~1e-8: a gradient that differs in its last
    #  bits -- cuDNN sums a batch of 4 and a batch of 8 in different orders -- may round the new value one ulp
    #  apart, which is no difference of the update; two ulps of the tensor's largest value are not counted)
    worst = 0.0
    for k, v in model.state_dict().items():
        d = (sharded[k] - v).abs().max().item()
        upd = (v - state0[k]).abs().max().item()
        slack = 2.0 * torch.finfo(v.dtype).eps * v.abs().max().item() if v.is_floating_point() else 0.0
        worst = max(worst, max(d - slack, 0.0) / max(upd, 1e-12))
    ok = worst < 1e-3 and abs(total.item() - loss1.item()) <= 1e-5 * abs(loss1.item())
    return dict(check="cfg4 sharded step == whole-batch step", world=world, loss_sharded=total.item(),
                loss_whole=loss1.item(), worst_param_diff_over_update=worst, ok=bool(ok))


def main():
    import argparse
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not args.check:
        raise SystemExit("timing lives in bench.py --workload cfg4; this entry point only has --check")
    out = check(world, rank, dev)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
