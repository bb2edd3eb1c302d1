"""Time the fused step's pieces on one GPU (stream-ordered and chained), for the library named by YH_LIB_PATH.
    python scratch/time_fused.py [batch]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odcp_b200 import ops, synthetic, targets

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
case = synthetic.headline(B)
lam = synthetic.DEFAULT_LAMBDAS
kw = dict(img_hw=(case.height, case.width), anchors=case.anchors, lambdas=lam)
R = 8
y0, gt, off = case.y.to(dev), targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
sets = [dict(y=y0.clone(), gt=gt.clone(), off=off.clone()) for _ in range(R)]
stream = torch.cuda.Stream(dev)
post = dict(conf_thre=0.5, iou_thre=0.45, max_out=128, want_cls_spec=False)

def fused(s, ov, key="res"):
    s[key] = ops.train_post(s["y"], s["gt"], s["off"], out=s.get(key), overlapped=ov, **kw, **post)
def train(s, ov, key="tr"):
    s[key] = ops.train_head(s["y"], s["gt"], s["off"], version=2, out=s.get(key), input_ready=ov, **kw)
def postp(s, ov):
    s["po"] = ops.postprocess(s["y"], version=2, img_hw=kw["img_hw"], anchors=kw["anchors"], out=s.get("po"), input_ready=ov, **post)

def timeit(fn, reps=200):
    with torch.cuda.stream(stream):
        fn(); stream.synchronize(); fn(); stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        for _ in range(5): g.replay()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300000)
        a.record(stream)
        for _ in range(reps): g.replay()
        b.record(stream)
        stream.synchronize()
    return a.elapsed_time(b) / (reps * R) * 1e3

out = {"lib": os.environ.get("YH_LIB_PATH", "default"), "batch": B}
out["train_iso"] = timeit(lambda: [train(s, False) for s in sets])
out["train_chain"] = timeit(lambda: [train(s, i > 0, "tr2") for i, s in enumerate(sets)])
out["fused_iso"] = timeit(lambda: [fused(s, False) for s in sets])
out["fused_chain"] = timeit(lambda: [fused(s, i > 0, "res2") for i, s in enumerate(sets)])
out["post_iso"] = timeit(lambda: [postp(s, False) for s in sets])
def unf(i, s):
    train(s, i > 0, "tr3"); postp(s, True)
out["unfused_chain"] = timeit(lambda: [unf(i, s) for i, s in enumerate(sets)])
print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()}))
