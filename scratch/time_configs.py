"""Kernel times of the BASELINE.json configurations (CUDA events, CUDA-graph replay over rotating
buffer sets sized to exceed L2), for the table in DESIGN.md.  Not part of the product."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odcp_b200 import ops, synthetic, targets

dev = torch.device("cuda:0")
lam = synthetic.DEFAULT_LAMBDAS
PEAK = 6552.6


def time_graph(fn, sets, reps=50):
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        for s in sets:
            fn(s)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for s in sets:
                fn(s)
        for _ in range(3):
            g.replay()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            g.replay()
        b.record(st)
        st.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * len(sets))


def run(name, case, conf_thre, iou_thre=0.45):
    nsets = max(2, min(8, int(400e6 // max(1, 2 * case.n * case.image_bytes)) + 1))
    kw = dict(version=case.version, img_hw=(case.height, case.width), anchors=case.anchors, boxes_per_cell=case.a)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    sets = []
    for _ in range(nsets):
        y = case.y.to(dev).clone()
        sets.append(dict(y=y, gt=gt.clone(), off=off.clone(),
                         out=dict(dy=torch.empty_like(y), loss=torch.empty((), device=dev), terms=torch.empty(5, device=dev))))

    def train(s):
        ops.train_head(s["y"], s["gt"], s["off"], lambdas=lam, out=s["out"], **kw)

    def post(s):
        s["post"] = ops.postprocess(s["y"], conf_thre=conf_thre, iou_thre=iou_thre, max_out=128, want_cls_spec=False,
                                    out=s.get("post"), **kw)

    def fused(s):
        s["fu"] = ops.train_post(s["y"], s["gt"], s["off"], img_hw=kw["img_hw"], anchors=kw["anchors"], lambdas=lam,
                                 conf_thre=conf_thre, iou_thre=iou_thre, max_out=128, want_cls_spec=False, out=s.get("fu"))

    t_fused = time_graph(fused, sets) if case.version == 2 and os.environ.get("YH_TIME_FUSED", "1") == "1" else None
    t_train = time_graph(train, sets)
    t_post = time_graph(post, sets)
    # predict(): the six decoded outputs (49 floats per predictor)
    dec_out = [ops.decode(s["y"], **kw) for s in sets[:2]]
    lib_kw = dict(kw)

    def dec(s):
        ops.decode(s["y"], **lib_kw)

    t_dec = time_graph(dec, sets[:2], reps=20)
    del dec_out
    kept = int(sets[0]["post"]["keep_cnt"].clamp(max=128).sum().item())
    p = case.image_bytes
    b_train = 2 * p * case.n + 48 * case.m
    b_post = p * case.n + 28 * kept
    print(json.dumps(dict(config=name, n=case.n, grid=[case.s_h, case.s_w], boxes=case.m, sets=nsets,
                          train_us=round(t_train, 2), train_MB=round(b_train / 1e6, 2),
                          train_frac=round(b_train / t_train / 1e3 / PEAK, 3),
                          post_us=round(t_post, 2), post_MB=round(b_post / 1e6, 2),
                          post_frac=round(b_post / t_post / 1e3 / PEAK, 3), kept_per_image=round(kept / case.n, 1),
                          fused_us=None if t_fused is None else round(t_fused, 2),
                          fused_frac_3P=None if t_fused is None else round((b_train + b_post) / t_fused / 1e3 / PEAK, 3),
                          decode_us_incl_alloc=round(t_dec, 2),
                          decode_MB=round((p * case.n + 49 * 4 * case.n * case.s_h * case.s_w * case.a) / 1e6, 2))), flush=True)


ONLY = os.environ.get("YH_TIME_ONLY", "")  # e.g. "cfg5,headline"
ALL = [("cfg1 v1 7x7 B=2 N=8", synthetic.cfg1), ("cfg2 v2 13x13 N=64", synthetic.cfg2),
       ("cfg3 v2 13x13 N=256 (inference, conf 0.5)", synthetic.cfg3), ("headline v2 13x13 N=256", synthetic.headline),
       ("cfg5 v2 19x19 N=512 50-100 boxes", synthetic.cfg5)]
for name, mk in ALL:
    if not ONLY or any(name.startswith(k) for k in ONLY.split(",")):
        run(name, mk(), 0.5)
