"""Repeat the headline and the dense-GT workloads many times and check that every run is bit-identical
to the first one (races in the flag / mbarrier / patch logic would show up as flaky gradients)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from odcp_b200 import ops, synthetic, targets

dev = torch.device("cuda:0")
lam = synthetic.DEFAULT_LAMBDAS
for name, case, reps in (("headline", synthetic.headline(), 300), ("cfg2+collisions", synthetic.with_collisions(synthetic.cfg2(), 60, seed=1), 300),
                         ("cfg3", synthetic.cfg3(), 300), ("cfg5", synthetic.cfg5(n=128), 60), ("cfg1", synthetic.cfg1(), 300)):
    kw = dict(version=case.version, img_hw=(case.height, case.width), anchors=case.anchors, boxes_per_cell=case.a)
    y = case.y.to(dev)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    ref = None
    bad = 0
    for i in range(reps):
        r = ops.train_head(y, gt, off, lambdas=lam, want_resp=True, **kw)
        # (alternating between the full class pick and the one that skips divisions: same labels and scores)
        p = ops.postprocess(y, conf_thre=0.5, iou_thre=0.45, max_out=128, input_ready=True, want_cls_spec=(i % 2 == 0), **kw)
        cur = [r["loss"].clone(), r["terms"].clone(), r["dy"].clone(), r["resp"].clone(), p["keep_cnt"].clone(), p["keep_idx"].clone(),
               p["label"].clone(), p["score"].clone(), p["bbox"].clone()]
        if ref is None:
            ref = cur
        else:
            cnt = ref[4].clamp(max=128)
            mask = torch.arange(128, device=dev)[None, :] < cnt[:, None]
            same = all(torch.equal(a, b) for a, b in zip(ref[:5], cur[:5]))
            same = same and torch.equal(ref[5][mask], cur[5][mask]) and torch.equal(ref[6][mask], cur[6][mask]) \
                and torch.equal(ref[7][mask], cur[7][mask]) and torch.equal(ref[8][mask], cur[8][mask])
            bad += 0 if same else 1
    torch.cuda.synchronize()
    print(name, "runs", reps, "mismatching runs", bad, "loss", float(ref[0]))
    assert bad == 0
print("stress ok")
