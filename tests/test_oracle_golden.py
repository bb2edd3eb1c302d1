"""The CPU oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle: the dense tier must be
bit-identical to the reference, the compact tier within the north-star tolerance."""
import numpy as np
import pytest
import torch

from conftest import golden_lambdas, load_golden, rel_err
from oracle import yolo_head_oracle as O

LOSS_FILES = ["v2_loss_small.npz", "v2_loss_nonsquare.npz", "v1_loss_small.npz", "v1_loss_b3c7.npz", "v2_loss_dense.npz"]
TOL = 1e-5  # north-star: loss and gradients within 1e-5 relative in fp32


@pytest.mark.parametrize("name", LOSS_FILES)
def test_dense_tier_is_bit_identical_to_reference(name):
    case, z = load_golden(name)
    torch.set_num_threads(1)
    r = O.train_head_dense(case, golden_lambdas(z))
    assert np.float32(r["loss"]) == z["loss"]
    assert np.array_equal(r["dy"], z["dy"])
    assert np.array_equal(r["resp"], z["resp"])
    assert np.array_equal(r["iou_resp"], z["iou_resp"])


@pytest.mark.parametrize("name", LOSS_FILES)
def test_dense_tier_chunked_recombination(name):
    case, z = load_golden(name)
    r = O.train_head_dense(case, golden_lambdas(z), chunk_images=1)
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert rel_err(r["dy"], z["dy"]) <= TOL
    assert np.array_equal(r["resp"], z["resp"])


@pytest.mark.parametrize("name", LOSS_FILES)
def test_compact_tier_matches_reference(name):
    case, z = load_golden(name)
    r = O.train_head_compact(case, golden_lambdas(z))
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert rel_err(r["dy"], z["dy"]) <= TOL
    # decisions: bit-exact unless the top-2 IoUs are within a few ulp (reported, not hidden)
    mism = np.nonzero(r["resp"] != z["resp"])[0]
    for j in mism:
        top2 = np.sort(z["iou_all"][j])[-2:]
        assert top2[1] - top2[0] <= 4 * np.spacing(top2[1]), ("non-tie mismatch", j, z["iou_all"][j])
    assert np.abs(r["iou_resp"] - z["iou_resp"]).max() <= 1e-6
    # zero pattern of the gradient is exact
    assert np.array_equal(r["dy"] != 0, z["dy"] != 0)


def test_cfg2_summary():
    from odcp_b200 import synthetic
    case, z = load_golden("v2_cfg2_summary.npz")
    live = synthetic.cfg2()
    assert abs(float(live.y.double().sum()) - float(z["y_sum"])) < 1e-9, "generator drifted from the golden"
    assert np.array_equal(live.rec, case.rec)
    r = O.train_head_compact(live, synthetic.DEFAULT_LAMBDAS)
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert np.array_equal(r["resp"], z["resp"])
    dy = r["dy"]
    rows = dy[live.rec["img"], live.rec["cy"], live.rec["cx"], r["resp"]]
    assert rel_err(rows, z["dy_rows"]) <= TOL
    assert rel_err(dy[..., 4].reshape(-1)[::7], z["dy_to_sample"]) <= TOL
    assert int(np.count_nonzero(dy)) == int(z["dy_nnz"])
    assert abs(np.abs(dy.astype(np.float64)).sum() - float(z["dy_abs_sum"])) <= TOL * float(z["dy_abs_sum"])


@pytest.mark.parametrize("name,version", [("v2_predict.npz", 2), ("v1_predict.npz", 1)])
def test_decode_matches_reference(name, version):
    case, z = load_golden(name)
    anchors = case.anchors if version == 2 else case.a
    outs = O.decode_torch(case.y, case.height, case.width, version, anchors)
    for i, t in enumerate(outs):
        assert np.array_equal(t.numpy(), z["out%d" % i]), i
    d = O.decode_boxes_np(case.y.numpy(), case.height, case.width, version, anchors)
    assert rel_err(d["bbox"], z["out2"]) <= 1e-6
    assert np.abs(d["conf"] - z["out3"]).max() <= 1e-6


@pytest.mark.parametrize("name", ["v2_nms_cfg3.npz", "v2_nms_default.npz", "v1_nms.npz"])
def test_nms_matches_reference(name):
    case, z = load_golden(name)
    anchors = case.anchors if case.version == 2 else case.a
    res = O.postprocess_np(case.y, case.height, case.width, case.version, anchors,
                           float(z["conf_thre"]), float(z["iou_thre"]))
    cnt = np.array([len(r["idx"]) for r in res], dtype=np.int32)
    assert np.array_equal(cnt, z["nms_cnt"])
    assert cnt.sum() > 0
    cat = lambda k: np.concatenate([r[k] for r in res], 0)
    assert np.array_equal(cat("idx"), z["nms_idx"])
    assert np.array_equal(cat("bbox"), z["nms_bbox"])
    assert np.array_equal(cat("conf"), z["nms_conf"])
    assert np.array_equal(cat("cls_spec"), z["nms_cls_spec"])
    # the torch port of the reference's shrinking-list loop (the CPU-baseline cost model)
    res2 = O.postprocess_torch(case.y, case.height, case.width, case.version, anchors,
                               float(z["conf_thre"]), float(z["iou_thre"]))
    cat2 = lambda k: np.concatenate([r[k] for r in res2], 0)
    for k, g in (("idx", "nms_idx"), ("bbox", "nms_bbox"), ("conf", "nms_conf"), ("cls_spec", "nms_cls_spec")):
        assert np.array_equal(cat2(k), z[g]), k


def test_iou_known_answers():
    z = dict(np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "iou_kat.npz")))
    assert np.array_equal(O.iou_np(z["b1"], z["b2"]), z["iou"])
    assert np.array_equal(O.iou_torch(torch.as_tensor(z["b1"]), torch.as_tensor(z["b2"])).numpy(), z["iou"])
    assert z["iou"][1] == 0 and z["iou"][2] == 0 and z["iou"][3] == 0 and z["iou"][4] == 0
    assert abs(z["iou"][0] - 1.0) < 1e-5


def test_evaluation_matching_and_ap_match_the_reference():
    """Oracle TP matching (models/utils.py:231-262) + the package's average_precision reproduce the
    per-class APs of the reference's own evaluate_model (tests/golden/evaluate.npz)."""
    import os
    from conftest import GOLDEN
    from odcp_b200.models.utils import average_precision
    z = dict(np.load(os.path.join(GOLDEN, "evaluate.npz")))
    n, ncls = int(z["n"]), int(z["num_cls"])
    tps, labs, scs = [], [], []
    for i in range(n):
        k = int(z["keep_cnt"][i])
        lo, hi = int(z["gt_off"][i]), int(z["gt_off"][i + 1])
        tps.append(O.match_detections_np(z["det_bbox"][i, :k], z["det_label"][i, :k], z["gt_boxes"][lo:hi],
                                         z["gt_labels"][lo:hi], z["levels"]))
        labs.append(z["det_label"][i, :k])
        scs.append(z["det_score"][i, :k])
    tp, lab, sc = np.concatenate(tps), np.concatenate(labs), np.concatenate(scs)
    for c in range(ncls):
        ap = average_precision(tp[lab == c], sc[lab == c], int(np.sum(z["gt_labels"] == c)))
        assert np.array_equal(ap, z["ap"][c]), c


# ------------------------------------------------------------------------------------------
# optimizer step: the oracle's restatement against torch.optim.SGD run the way
# the reference runs it (tests/golden/make_sgd_golden.py)
# ------------------------------------------------------------------------------------------
def _sgd_golden():
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "sgd_step.npz"))
    n = int(z["n_tensors"])
    return z, n


def test_sgd_oracle_fresh_optimizer_per_iteration():
    z, n = _sgd_golden()
    ps = [z["p0_%d" % i] for i in range(n)]
    for it, lr in enumerate(z["lrs"]):
        gs = [z["fresh_g%d_%d" % (it, i)] for i in range(n)]
        ps, _ = O.sgd_step_np(ps, gs, float(lr), float(z["momentum"]), float(z["weight_decay"]), bufs=None)
    for i in range(n):
        assert np.allclose(ps[i], z["fresh_p2_%d" % i], rtol=1e-6, atol=1e-7), i
        assert not np.array_equal(ps[i], z["p0_%d" % i])


def test_sgd_oracle_persistent_momentum():
    z, n = _sgd_golden()
    ps = [z["p0_%d" % i] for i in range(n)]
    bufs = None
    for it in range(3):
        gs = [z["pers_g%d_%d" % (it, i)] for i in range(n)]
        ps, bufs = O.sgd_step_np(ps, gs, float(z["lrs"][2]), float(z["momentum"]), float(z["weight_decay"]), bufs=bufs)
    for i in range(n):
        assert np.allclose(ps[i], z["pers_p2_%d" % i], rtol=1e-6, atol=1e-7), i
    # the two modes differ (momentum accumulates only in one of them)
    assert not np.allclose(ps[4], z["fresh_p2_4"], rtol=1e-6, atol=1e-7)
