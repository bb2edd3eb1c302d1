"""Seeded random geometries through every kernel variant of the path (compile-time and run-time geometry, whole-image
and general staging, 512- and 1024-thread CTAs, patched / after-stream / four-per-warp records, the fused step and
its fall-back to the separate kernels): train head against the oracle's closed form, post-process against the
oracle's per-image NMS, fused step bit for bit against the separate calls.  Grid, boxes per cell, classes, batch,
box density, objectness shift, thresholds and class-aware suppression are all drawn from the seed.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from odcp_b200 import ops, synthetic, targets
from oracle import yolo_head_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
LAM = synthetic.DEFAULT_LAMBDAS


def draw_case(seed):
    rng = np.random.default_rng(1000 + seed)
    version = 2 if rng.random() < 0.7 else 1
    s_h, s_w = int(rng.integers(2, 20)), int(rng.integers(2, 20))
    a = int(rng.integers(1, 7))                      # 5a <= 32
    c = int(rng.choice([1, 2, 3, 7, 20, 27, 28, 40, 80, 90]))
    if s_h * s_w * a * (5 + c) > 120_000:            # keep the oracle in seconds
        c = 20
    n = int(rng.integers(1, 33))
    k_hi = int(rng.choice([1, 3, 5, 12, 40]))
    k_lo = int(rng.integers(0, 2)) if k_hi > 1 else 1
    cell = int(rng.choice([16, 32]))
    anchors = tuple((float(rng.uniform(0.5, s_w)), float(rng.uniform(0.5, s_h))) for _ in range(a))
    case = synthetic.make_case("fuzz%d" % seed, version, n, s_h, s_w, a, c, s_h * cell, s_w * cell, seed=2000 + seed,
                               k_lo=k_lo, k_hi=k_hi, to_shift=float(rng.uniform(-2.5, 0.5)), anchors=anchors)
    if case.m == 0:  # (the reference cannot form its means without a box)
        case = synthetic.make_case("fuzz%d" % seed, version, n, s_h, s_w, a, c, s_h * cell, s_w * cell, seed=2000 + seed,
                                   k_lo=1, k_hi=max(k_hi, 1), to_shift=-1.0, anchors=anchors)
    conf = float(rng.choice([0.3, 0.5, 0.7, 0.9]))
    iou = float(rng.choice([0.3, 0.45, 0.6]))
    return case, conf, iou, bool(rng.random() < 0.3)


@pytest.mark.parametrize("seed", range(64))
def test_random_geometry_against_the_oracle(seed, cuda_device):
    dev = cuda_device
    case, conf, iou, class_aware = draw_case(seed)
    if not synthetic.distinct_scores(case.y, case.version, case.a):
        pytest.skip("equal objectness logits inside an image: the NMS order is not unique")
    kw = dict(version=case.version, img_hw=(case.height, case.width), anchors=case.anchors, boxes_per_cell=case.a)
    y = case.y.to(dev)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)

    # train head vs the closed form
    tr = ops.train_head(y, gt, off, lambdas=LAM, want_resp=True, **kw)
    want = O.train_head_compact(case, LAM)
    assert abs(float(tr["loss"]) - want["loss"]) <= TOL * abs(want["loss"]), (float(tr["loss"]), want["loss"])
    assert np.abs(tr["terms"].cpu().numpy() - want["terms"]).max() <= TOL * np.abs(want["terms"]).max()
    resp = tr["resp"].cpu().numpy()
    for j in np.nonzero(resp != want["resp"])[0]:  # (only exact IoU ties may differ)
        top2 = np.sort(want["iou_all"][j])[-2:]
        assert top2[1] - top2[0] <= 4 * np.spacing(np.float32(top2[1])), (int(j), want["iou_all"][j])
    if np.array_equal(resp, want["resp"]):
        dy = tr["dy"].cpu().numpy()
        assert rel_err(dy, want["dy"]) <= TOL
        assert np.array_equal(dy != 0, want["dy"] != 0)

    # post-process vs the per-image NMS of the oracle
    po = ops.postprocess(y, conf_thre=conf, iou_thre=iou, class_aware=class_aware, **kw)
    wp = O.postprocess_np(case.y, case.height, case.width, case.version, case.anchors if case.version == 2 else case.a,
                          conf, iou, class_aware=class_aware)
    cnt = po["keep_cnt"].cpu().numpy()
    idx, lab = po["keep_idx"].cpu().numpy(), po["label"].cpu().numpy()
    for n, w in enumerate(wp):
        assert cnt[n] == len(w["idx"]), (n, cnt[n], len(w["idx"]))
        assert np.array_equal(idx[n, :cnt[n]], w["idx"]), n
        if case.c > 1:
            near_tie = np.sort(w["cls_spec"], -1)[:, -1] - np.sort(w["cls_spec"], -1)[:, -2] <= 1e-7 if len(w["idx"]) else np.zeros(0, bool)
            assert np.array_equal(lab[n, :cnt[n]][~near_tie], w["label"][~near_tie]), n

    # the fused step: bit for bit what the two separate calls give
    if case.version == 2:
        fs = ops.train_post(y, gt, off, img_hw=kw["img_hw"], anchors=case.anchors, lambdas=LAM, conf_thre=conf, iou_thre=iou,
                            class_aware=class_aware, want_resp=True)
        torch.cuda.synchronize()
        assert torch.equal(fs["train"]["dy"], tr["dy"]) and torch.equal(fs["train"]["resp"], tr["resp"])
        assert torch.equal(fs["train"]["iou_resp"], tr["iou_resp"])
        assert abs(float(fs["train"]["loss"]) - float(tr["loss"])) <= 2e-6 * abs(float(tr["loss"]))
        assert torch.equal(fs["post"]["keep_cnt"], po["keep_cnt"])
        mo = po["keep_idx"].shape[1]
        valid = (torch.arange(mo, device=dev)[None, :] < po["keep_cnt"].clamp(max=mo)[:, None])
        for key in ("keep_idx", "label", "score", "conf", "bbox"):
            assert torch.equal(fs["post"][key][valid], po[key][valid]), key
