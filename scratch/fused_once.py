import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odcp_b200 import ops, synthetic, targets
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
case = synthetic.headline(n)
y, gt, off = case.y.to(dev), targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
r = ops.train_post(y, gt, off, img_hw=(case.height, case.width), lambdas=synthetic.DEFAULT_LAMBDAS, anchors=case.anchors,
                   conf_thre=0.5, iou_thre=0.45, max_out=128, want_cls_spec=False)
torch.cuda.synchronize()
print("loss", float(r["train"]["loss"]), "kept", int(r["post"]["keep_cnt"].sum()))
