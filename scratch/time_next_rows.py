"""Timings of the SURVEY 8(f) rows (target builder, head layout, batched evaluation matching, fused SGD step)
next to what they replace, for DESIGN.md section 8.  CUDA events, graph replay where the call is capturable.
Not part of the product."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from odcp_b200 import ops, synthetic, targets
from odcp_b200.optim import SGD, reset_state
from oracle import yolo_head_oracle as O

dev = torch.device("cuda:0")
PEAK = 6552.6


def gpu_us(fn, reps=50, graph=True):
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        st.synchronize()
        run = fn
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                fn()
            run = g.replay
        for _ in range(3):
            run()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            run()
        b.record(st)
        st.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def cpu_ms(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return 1e3 * min(ts)


def out(**kw):
    print(json.dumps(kw), flush=True)


# ---- 1. target builder: boxes -> records + CSR offsets
for name, case in (("headline N=256", synthetic.headline()), ("cfg5 N=512", synthetic.cfg5())):
    rec = case.rec
    boxes = np.stack([rec["x1"], rec["y1"], rec["x2"], rec["y2"]], 1).astype(np.float64)
    labels = rec["cls"].astype(np.int32)
    img = rec["img"].astype(np.int32)
    db, dl, di = torch.from_numpy(boxes).to(dev), torch.from_numpy(labels).to(dev), torch.from_numpy(img).to(dev)
    kw = dict(num_images=case.n, version=2, img_hw=(case.height, case.width), grid=(case.s_h, case.s_w))
    us = gpu_us(lambda: ops.build_targets(db, dl, di, **kw), graph=False)
    ms = cpu_ms(lambda: targets.boxes_to_records(boxes, labels, img, case.height, case.width, case.s_h, case.s_w, 2))
    dense_mb = case.m * case.s_h * case.s_w * (28 * 4 + 8) / 1e6
    out(row="f1 target builder", config=name, boxes=case.m, gpu_us=round(us, 2), cpu_numpy_ms=round(ms, 3),
        dense_targets_avoided_MB=round(dense_mb, 1))

# ---- 2. head layout: the permute+reshape copy a channels_last head conv makes unnecessary
x = torch.randn(256, 125, 13, 13, device=dev)
us_copy = gpu_us(lambda: x.permute(0, 2, 3, 1).reshape(256, 13, 13, 5, 25).contiguous())
xcl = x.contiguous(memory_format=torch.channels_last)
v = xcl.permute(0, 2, 3, 1).reshape(256, 13, 13, 5, 25)
out(row="f2 head layout", nchw_permute_copy_us=round(us_copy, 2), bytes_moved_MB=round(2 * x.numel() * 4 / 1e6, 1),
    channels_last_is_view=bool(v.data_ptr() == xcl.data_ptr()), channels_last_copy_us=0.0)

# ---- 3. evaluation matching
case = synthetic.cfg3()
post = ops.postprocess(case.y.to(dev), version=2, img_hw=(416, 416), conf_thre=0.5, iou_thre=0.45, anchors=case.anchors,
                       max_out=128, want_cls_spec=False)
rng = np.random.default_rng(1)
k = rng.integers(1, 6, size=case.n)
gt_off = np.concatenate([[0], np.cumsum(k)]).astype(np.int32)
m = int(gt_off[-1])
gb = rng.uniform(0, 300, size=(m, 4))
gb[:, 2:] = gb[:, :2] + rng.uniform(20, 116, size=(m, 2))
gl = rng.integers(0, 20, size=m).astype(np.int32)
levels = [0.5 + 0.05 * i for i in range(10)]
dgb, dgl, dgo = torch.from_numpy(gb).to(dev), torch.from_numpy(gl).to(dev), torch.from_numpy(gt_off).to(dev)
us = gpu_us(lambda: ops.match_detections(post, dgb, dgl, dgo, levels), graph=False)
hb, hl, hc = post["bbox"].cpu().numpy(), post["label"].cpu().numpy(), post["keep_cnt"].cpu().numpy()


def cpu_match():
    for n in range(case.n):
        O.match_detections_np(hb[n, :hc[n]], hl[n, :hc[n]], gb[gt_off[n]:gt_off[n + 1]], gl[gt_off[n]:gt_off[n + 1]], levels)


out(row="f3 evaluation matching", images=case.n, detections=int(hc.sum()), gt_boxes=m, gpu_us=round(us, 2),
    cpu_numpy_ms=round(cpu_ms(cpu_match), 2))

# ---- 4. fused SGD step on the YOLOv2 (Darknet-19 + head) parameter shapes
cfg = [(3, 32, 3), (32, 64, 3), (64, 128, 3), (128, 64, 1), (64, 128, 3), (128, 256, 3), (256, 128, 1), (128, 256, 3),
       (256, 512, 3), (512, 256, 1), (256, 512, 3), (512, 256, 1), (256, 512, 3), (512, 1024, 3), (1024, 512, 1),
       (512, 1024, 3), (1024, 512, 1), (512, 1024, 3), (1024, 1024, 3), (1024, 1024, 3), (512, 64, 1), (1280, 1024, 3),
       (1024, 125, 1)]
shapes = []
for cin, cout, ks in cfg:
    shapes += [(cout, cin, ks, ks), (cout,), (cout,)]   # conv weight, batch-norm weight and bias
params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.01) for s in shapes]
for p in params:
    p.grad = torch.randn_like(p) * 0.01
n_el = sum(p.numel() for p in params)
reset_state()
ours = SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)
us_ours = gpu_us(ours.step)
ref = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)


def torch_fresh():   # the reference: a new optimizer (empty state) in every iteration
    ref.state.clear()
    ref.step()


us_torch = gpu_us(torch_fresh, graph=False, reps=20)
out(row="f4 fused SGD step (reference mode: new optimizer per iteration)", tensors=len(params), parameters=n_el,
    fused_us=round(us_ours, 1), algorithmic_MB=round(12 * n_el / 1e6, 1), frac_of_hbm=round(12 * n_el / us_ours / 1e3 / PEAK, 3),
    torch_optim_sgd_us=round(us_torch, 1))
reset_state()
pers = SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4, persistent_momentum=True)
us_p = gpu_us(pers.step)
ref2 = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)
ref2.step()
us_t2 = gpu_us(ref2.step, graph=False, reps=20)
out(row="f4 fused SGD step (persistent momentum)", fused_us=round(us_p, 1), algorithmic_MB=round(20 * n_el / 1e6, 1),
    frac_of_hbm=round(20 * n_el / us_p / 1e3 / PEAK, 3), torch_optim_sgd_us=round(us_t2, 1))
