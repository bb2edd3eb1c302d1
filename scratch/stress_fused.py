"""The fused step, repeated many times on several workloads (stream-ordered and as an overlapped chain over rotating
buffer sets), every run compared bit for bit with the separate calls' results: a race between the dense pass, the
records and the NMS team (mbarriers bar_list / bar_dense, named barriers) would show up as a flaky gradient or box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odcp_b200 import ops, synthetic, targets

dev = torch.device("cuda:0")
lam = synthetic.DEFAULT_LAMBDAS
CASES = (("headline", synthetic.headline(), 300), ("collisions", synthetic.with_collisions(synthetic.headline(64), 200, seed=7), 300),
         ("cfg5_n128", synthetic.cfg5(n=128), 60), ("mid", synthetic.make_case("mid", 2, 96, 13, 13, 5, 20, 416, 416, seed=48, k_lo=6, k_hi=14, to_shift=-1.563), 300),
         ("cfg2_overflow", synthetic.cfg2(), 100))
for name, case, reps in CASES:
    kw = dict(img_hw=(case.height, case.width), anchors=case.anchors)
    y = case.y.to(dev)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    tr = ops.train_head(y, gt, off, version=2, lambdas=lam, want_resp=True, **kw)
    po = ops.postprocess(y, version=2, conf_thre=0.5, iou_thre=0.45, max_out=128, want_cls_spec=False, **kw)
    cnt = po["keep_cnt"].clamp(max=128)
    mask = torch.arange(128, device=dev)[None, :] < cnt[:, None]
    R = 4
    ys = [y.clone() for _ in range(R)]
    outs = [None] * R
    bad = 0
    for i in range(reps):
        k = i % R
        outs[k] = ops.train_post(ys[k], gt, off, lambdas=lam, conf_thre=0.5, iou_thre=0.45, max_out=128, want_cls_spec=False,
                                 want_resp=True, out=outs[k], overlapped=(k != 0), **kw)
        if k == R - 1:
            torch.cuda.synchronize()
            for o in outs:
                t, p = o["train"], o["post"]
                same = torch.equal(t["dy"], tr["dy"]) and torch.equal(t["resp"], tr["resp"]) and torch.equal(p["keep_cnt"], po["keep_cnt"])
                same = same and all(torch.equal(p[key][mask], po[key][mask]) for key in ("keep_idx", "label", "score", "bbox", "conf"))
                bad += 0 if same else 1
                t["dy"].fill_(float("nan"))
                p["keep_cnt"].fill_(-3)
    print(name, "fused runs", reps, "mismatching runs", bad)
    assert bad == 0
print("fused stress ok")
