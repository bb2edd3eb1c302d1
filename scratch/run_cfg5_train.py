"""A few train-head launches on BASELINE config 5 (for ncu captures).  Not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odcp_b200 import ops, synthetic, targets
dev = torch.device("cuda:0")
case = getattr(synthetic, os.environ.get("YH_CASE", "cfg5"))()
y = case.y.to(dev)
gt = targets.records_to_tensor(case.rec, dev)
off = torch.from_numpy(case.gt_off).to(dev)
out = None
for _ in range(4):
    out = ops.train_head(y, gt, off, version=2, img_hw=(case.height, case.width), anchors=case.anchors,
                         lambdas=synthetic.DEFAULT_LAMBDAS, out=out)
torch.cuda.synchronize()
print(float(out["loss"]))
