"""Parity of the CUDA path (through the C ABI, via odcp_b200.ops) against
  (1) the golden vectors produced by the reference itself (tests/golden/*.npz),
  (2) the CPU oracle on seeded synthetic inputs,
  (3) size-independent properties at the BASELINE.json full sizes.

Tolerances (north-star): responsible-predictor and kept-box indices bit-exact, except decisions
whose two best candidates are within a few ulp of each other (those are listed in the assertion
message, never silently skipped); loss and dL/dy within 1e-5 relative in fp32.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_lambdas, load_golden, rel_err
from odcp_b200 import ops, synthetic, targets
from oracle import yolo_head_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
LOSS_FILES = ["v2_loss_small.npz", "v2_loss_nonsquare.npz", "v1_loss_small.npz", "v1_loss_b3c7.npz", "v2_loss_dense.npz"]


def run_train(case, lambdas, dev, m_global=None, want_grad=True):
    y = case.y.to(dev)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    r = ops.train_head(y, gt, off, version=case.version, img_hw=(case.height, case.width),
                       lambdas=lambdas, anchors=case.anchors, boxes_per_cell=case.a,
                       m_global=m_global, want_grad=want_grad, want_resp=True)
    torch.cuda.synchronize()
    return dict(loss=float(r["loss"].item()), terms=r["terms"].cpu().numpy(),
                dy=None if r["dy"] is None else r["dy"].cpu().numpy(),
                resp=r["resp"].cpu().numpy(), iou_resp=r["iou_resp"].cpu().numpy())


def assert_decisions(got, want, iou_all=None, what="resp"):
    mism = np.nonzero(got != want)[0]
    for j in mism:
        assert iou_all is not None, ("%s mismatch at record %d" % (what, j))
        top2 = np.sort(iou_all[j])[-2:]
        assert top2[1] - top2[0] <= 4 * np.spacing(np.float32(top2[1])), \
            ("non-tie %s mismatch" % what, int(j), iou_all[j], int(got[j]), int(want[j]))


# ------------------------------------------------------------------------------------------
# train head
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", LOSS_FILES)
def test_train_head_matches_reference_golden(name, cuda_device):
    case, z = load_golden(name)
    r = run_train(case, golden_lambdas(z), cuda_device)
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert r["dy"].shape == z["dy"].shape
    assert rel_err(r["dy"], z["dy"]) <= TOL
    assert_decisions(r["resp"], z["resp"], z["iou_all"])
    assert np.abs(r["iou_resp"] - z["iou_resp"]).max() <= 2e-6
    assert np.array_equal(r["dy"] != 0, z["dy"] != 0), "gradient sparsity pattern differs"
    # elementwise on the non-zeros as well (not only relative to the largest entry)
    assert np.allclose(r["dy"], z["dy"], rtol=1e-3, atol=1e-6 * np.abs(z["dy"]).max())


@pytest.mark.parametrize("name", LOSS_FILES)
def test_train_head_loss_only_variant(name, cuda_device):
    case, z = load_golden(name)
    r = run_train(case, golden_lambdas(z), cuda_device, want_grad=False)
    assert r["dy"] is None
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert_decisions(r["resp"], z["resp"], z["iou_all"])


def test_train_head_cfg2_full_vs_reference_summary(cuda_device):
    case, z = load_golden("v2_cfg2_summary.npz")
    live = synthetic.cfg2()
    r = run_train(live, synthetic.DEFAULT_LAMBDAS, cuda_device)
    assert abs(r["loss"] - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert np.array_equal(r["resp"], z["resp"])
    dy = r["dy"]
    rows = dy[live.rec["img"], live.rec["cy"], live.rec["cx"], r["resp"]]
    assert rel_err(rows, z["dy_rows"]) <= TOL
    assert rel_err(dy[..., 4].reshape(-1)[::7], z["dy_to_sample"]) <= TOL
    assert int(np.count_nonzero(dy)) == int(z["dy_nnz"])
    assert abs(np.abs(dy.astype(np.float64)).sum() - float(z["dy_abs_sum"])) <= TOL * float(z["dy_abs_sum"])


CASES = {
    "cfg1": lambda: synthetic.cfg1(),
    "cfg1_coll": lambda: synthetic.with_collisions(synthetic.cfg1(), 9, seed=3),
    "cfg2": lambda: synthetic.cfg2(),
    "cfg2_coll": lambda: synthetic.with_collisions(synthetic.cfg2(), 40, seed=4),
    "cfg2_10gt": lambda: synthetic.cfg2(k_hi=19),
    "v2_19x19_n8": lambda: synthetic.cfg5(n=8),
    "v2_nonsq_odd": lambda: synthetic.make_case("odd", 2, 3, 9, 11, 5, 20, 288, 352, seed=21, k_hi=6),
    "v2_1img": lambda: synthetic.make_case("one", 2, 1, 13, 13, 5, 20, 416, 416, seed=22),
    "v2_c80_a3": lambda: synthetic.make_case("coco", 2, 5, 13, 13, 3, 80, 416, 416, seed=23,
                                             anchors=synthetic.YOLOV2_ANCHORS[:3]),
    "v1_b3_c7": lambda: synthetic.make_case("v1b3", 1, 5, 5, 6, 3, 7, 160, 192, seed=24),
    "v2_c1": lambda: synthetic.make_case("c1", 2, 4, 13, 13, 5, 1, 416, 416, seed=25),
    # batches with more than 12 records per tile on average (BASELINE config 5's density): the train head processes
    # every record of such a tile after its dense pass, all warps at once, records dealt by cell (compile-time and
    # run-time geometry, v2 and v1, collisions on top of dense values, C beyond the class registers)
    "dense_v2_19x19_n64": lambda: synthetic.make_case("d19", 2, 64, 19, 19, 5, 20, 608, 608, seed=26, k_lo=85, k_hi=110),
    "dense_v2_13x13_coll": lambda: synthetic.with_collisions(
        synthetic.make_case("d13", 2, 40, 13, 13, 5, 20, 416, 416, seed=27, k_lo=80, k_hi=110), 2500, seed=5),
    "dense_v2_c80_a3": lambda: synthetic.make_case("dcoco", 2, 32, 13, 13, 3, 80, 416, 416, seed=28, k_lo=130, k_hi=160,
                                                   anchors=synthetic.YOLOV2_ANCHORS[:3]),
    "dense_v1_7x7": lambda: synthetic.make_case("dv1", 1, 64, 7, 7, 2, 20, 448, 448, seed=29, k_lo=80, k_hi=100),
    "dense_v1_b3_c7": lambda: synthetic.make_case("dv1b3", 1, 96, 5, 6, 3, 7, 160, 192, seed=30, k_lo=60, k_hi=80),
}


@pytest.mark.parametrize("key", sorted(CASES))
def test_train_head_vs_oracle(key, cuda_device):
    case = CASES[key]()
    lam = synthetic.DEFAULT_LAMBDAS
    want = O.train_head_compact(case, lam)
    r = run_train(case, lam, cuda_device)
    assert abs(r["loss"] - want["loss"]) <= TOL * abs(want["loss"])
    assert np.abs(r["terms"] - want["terms"]).max() <= TOL * np.abs(want["terms"]).max()
    assert_decisions(r["resp"], want["resp"], want["iou_all"])
    assert rel_err(r["dy"], want["dy"]) <= TOL
    assert np.array_equal(r["dy"] != 0, want["dy"] != 0)
    assert np.abs(r["iou_resp"] - want["iou_resp"]).max() <= 2e-6


def test_train_head_images_without_boxes(cuda_device):
    """Images with k_n = 0 get an all-zero gradient."""
    case = synthetic.make_case("empty", 2, 6, 13, 13, 5, 20, 416, 416, seed=31, k_lo=0, k_hi=2)
    k = np.diff(case.gt_off)
    assert (k == 0).any() and (k > 0).any()
    want = O.train_head_compact(case, synthetic.DEFAULT_LAMBDAS)
    r = run_train(case, synthetic.DEFAULT_LAMBDAS, cuda_device)
    assert abs(r["loss"] - want["loss"]) <= TOL * abs(want["loss"])
    assert rel_err(r["dy"], want["dy"]) <= TOL
    for n in np.nonzero(k == 0)[0]:
        assert not r["dy"][n].any()


def test_train_head_no_boxes_is_an_error(cuda_device):
    from odcp_b200._lib import YoloHeadError
    case = synthetic.make_case("none", 2, 2, 13, 13, 5, 20, 416, 416, seed=32, k_lo=0, k_hi=0)
    with pytest.raises(YoloHeadError) as e:
        run_train(case, synthetic.DEFAULT_LAMBDAS, cuda_device)
    assert e.value.code == -2


def test_train_head_is_deterministic_and_workspace_reusable(cuda_device):
    case = synthetic.with_collisions(synthetic.cfg2(), 40, seed=4)
    a = run_train(case, synthetic.DEFAULT_LAMBDAS, cuda_device)
    for _ in range(3):
        b = run_train(case, synthetic.DEFAULT_LAMBDAS, cuda_device)
        assert a["loss"] == b["loss"]
        assert np.array_equal(a["dy"], b["dy"])


def test_train_head_unaligned_views(cuda_device):
    """y / dy that are only 4-byte aligned take the non-TMA path; results must not change."""
    case = synthetic.cfg2(n=5)
    lam = synthetic.DEFAULT_LAMBDAS
    base = run_train(case, lam, cuda_device)
    dev = cuda_device
    buf = torch.zeros(case.y.numel() + 1, device=dev)
    y = buf[1:].view(case.y.shape)
    y.copy_(case.y)
    assert y.data_ptr() % 16 == 4
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    dybuf = torch.empty(case.y.numel() + 3, device=dev)
    dy = dybuf[3:].view(case.y.shape)
    r = ops.train_head(y, gt, off, version=2, img_hw=(case.height, case.width), lambdas=lam,
                       anchors=case.anchors, out=dict(dy=dy))
    assert float(r["loss"].item()) == base["loss"]
    assert np.array_equal(dy.cpu().numpy(), base["dy"])


def test_train_head_shards_recombine(cuda_device):
    """Chunk additivity: image shards evaluated with the global box count add up
    to the unsharded result -- the property the multi-GPU path relies on."""
    case = synthetic.with_collisions(synthetic.cfg2(), 40, seed=4)
    lam = synthetic.DEFAULT_LAMBDAS
    full = run_train(case, lam, cuda_device)
    loss = 0.0
    parts = []
    for lo, hi in ((0, 13), (13, 14), (14, 40), (40, 64)):
        rec, off = targets.shard_records(case.rec, case.gt_off, lo, hi)
        sub = synthetic.HeadCase("s", 2, hi - lo, case.s_h, case.s_w, case.a, case.c, case.height,
                                 case.width, case.y[lo:hi].contiguous(), rec, off, anchors=case.anchors)
        r = run_train(sub, lam, cuda_device, m_global=case.m)
        loss += r["loss"]
        parts.append(r["dy"])
    assert abs(loss - full["loss"]) <= TOL * abs(full["loss"])
    assert np.array_equal(np.concatenate(parts, 0), full["dy"])


def test_train_head_full_size_properties_cfg5(cuda_device):
    """BASELINE cfg 5 at full size (N=512, 19x19, 50..100 boxes/image): the dense reference
    cannot run it, so check (a) an 8-image slice against the oracle, (b) permutation invariance
    of the box order inside images, (c) shard additivity."""
    case = synthetic.cfg5()
    lam = synthetic.DEFAULT_LAMBDAS
    full = run_train(case, lam, cuda_device)
    assert np.isfinite(full["loss"])
    # (a) slice vs oracle, same global M
    rec, off = targets.shard_records(case.rec, case.gt_off, 100, 108)
    sub = synthetic.HeadCase("s", 2, 8, case.s_h, case.s_w, case.a, case.c, case.height, case.width,
                             case.y[100:108].contiguous(), rec, off, anchors=case.anchors)
    want = O.train_head_compact(sub, lam, m_global=case.m)
    assert rel_err(full["dy"][100:108], want["dy"]) <= TOL
    lo, hi = int(case.gt_off[100]), int(case.gt_off[108])
    assert_decisions(full["resp"][lo:hi], want["resp"], want["iou_all"])
    # (b) reverse the box order inside every image
    perm = np.concatenate([np.arange(case.gt_off[n + 1] - 1, case.gt_off[n] - 1, -1) for n in range(case.n)])
    pc = synthetic.HeadCase("p", 2, case.n, case.s_h, case.s_w, case.a, case.c, case.height, case.width,
                            case.y, case.rec[perm], case.gt_off, anchors=case.anchors)
    rp = run_train(pc, lam, cuda_device)
    assert abs(rp["loss"] - full["loss"]) <= TOL * abs(full["loss"])
    assert rel_err(rp["dy"], full["dy"]) <= TOL
    assert np.array_equal(rp["resp"], full["resp"][perm])
    # (c) two shards
    tot = 0.0
    for a, b in ((0, 200), (200, 512)):
        rec, off = targets.shard_records(case.rec, case.gt_off, a, b)
        s = synthetic.HeadCase("s", 2, b - a, case.s_h, case.s_w, case.a, case.c, case.height, case.width,
                               case.y[a:b].contiguous(), rec, off, anchors=case.anchors)
        r = run_train(s, lam, cuda_device, m_global=case.m)
        tot += r["loss"]
        assert np.array_equal(r["dy"], full["dy"][a:b])
    assert abs(tot - full["loss"]) <= TOL * abs(full["loss"])


# ------------------------------------------------------------------------------------------
# device-side target builder
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["v2_collate.npz", "v2_collate_nonsquare.npz", "v1_collate.npz"])
def test_build_targets_matches_reference_collate_fn(name, cuda_device):
    """yh_build_targets vs the records read back from the reference's own collate_fn grids: bit-exact."""
    z = dict(np.load(os.path.join(GOLDEN, name)))
    want = np.ascontiguousarray(z["rec"]).reshape(-1).view(targets.GT_DTYPE)
    gt, off, status = ops.build_targets(torch.from_numpy(z["boxes"]).to(cuda_device), torch.from_numpy(z["labels"]),
                                        torch.from_numpy(z["img"]), num_images=int(z["n"]), version=int(z["version"]),
                                        img_hw=(int(z["height"]), int(z["width"])), grid=(int(z["s_h"]), int(z["s_w"])))
    assert status.cpu().tolist() == [0, 0]
    assert targets.tensor_to_records(gt).tobytes() == want.tobytes()
    assert np.array_equal(off.cpu().numpy(), targets.csr_offsets(want, int(z["n"])))


def test_build_targets_full_size_and_bad_input(cuda_device):
    """cfg 5 sized input (tens of thousands of boxes) against the host builder, images without boxes,
    and the status counters for out-of-order / out-of-range boxes."""
    rng = np.random.default_rng(7)
    boxes, labels, img = synthetic.make_boxes(rng, 512, 608, 608, 0, 100, 20)
    want = targets.boxes_to_records(boxes, labels, img, 608, 608, 19, 19, 2)
    gt, off, status = ops.build_targets(torch.from_numpy(boxes).to(cuda_device), torch.from_numpy(labels),
                                        torch.from_numpy(img), num_images=512, version=2, img_hw=(608, 608), grid=(19, 19))
    assert status.cpu().tolist() == [0, 0]
    assert targets.tensor_to_records(gt).tobytes() == want.tobytes()
    assert np.array_equal(off.cpu().numpy(), targets.csr_offsets(want, 512))
    bad_img = img.copy()
    bad_img[[5, 6]] = bad_img[[6, 5]] if bad_img[5] != bad_img[6] else (bad_img[6] + 1, bad_img[5])
    bad_boxes = boxes.copy()
    bad_boxes[0] = [700.0, 10.0, 800.0, 50.0]  # outside the 608-pixel image: cell out of range
    _, _, status = ops.build_targets(torch.from_numpy(bad_boxes).to(cuda_device), torch.from_numpy(labels),
                                     torch.from_numpy(bad_img), num_images=512, version=2, img_hw=(608, 608), grid=(19, 19))
    st = status.cpu().tolist()
    assert st[0] >= 1 and st[1] >= 1


def test_get_loss_from_boxes_equals_the_dense_signature(cuda_device):
    """The three entry forms of the drop-in model (dense reference grids, compact records, raw boxes)
    give the same loss and the same gradient."""
    from odcp_b200.models.yolov2 import YOLOv2Head
    z = dict(np.load(os.path.join(GOLDEN, "v2_collate.npz")))
    n, h, w = int(z["n"]), int(z["height"]), int(z["width"])
    rec = np.ascontiguousarray(z["rec"]).reshape(-1).view(targets.GT_DTYPE)
    y0 = torch.randn(n, 13, 13, 5, 25, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    x = torch.zeros(n, h, w, 3, device=cuda_device)
    m = YOLOv2Head(num_cls=20).to(cuda_device)
    res = []
    for form in ("dense", "boxes"):
        y = y0.clone().requires_grad_(True)
        m.set_head_output(y)
        if form == "dense":
            dense = [t.to(cuda_device) for t in targets.records_to_dense(rec, n, 13, 13, 20, 2)]
            loss = m.get_loss(x, *dense, **synthetic.DEFAULT_LAMBDAS)
        else:
            loss = m.get_loss_from_boxes(x, z["boxes"], z["labels"], z["img"], **synthetic.DEFAULT_LAMBDAS)
        loss.backward()
        res.append((float(loss.item()), y.grad.cpu().numpy()))
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])


# ------------------------------------------------------------------------------------------
# head-tensor layout (channels_last head conv: permute + reshape become a view)
# ------------------------------------------------------------------------------------------
def test_channels_last_head_feeds_the_kernel_without_a_copy(cuda_device):
    from odcp_b200.models import layout
    from odcp_b200.models.yolov2 import YOLOv2HeadOps

    class TinyV2(YOLOv2HeadOps, torch.nn.Module):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.conv = torch.nn.Conv2d(16, 5 * 25, 1)
            self.cls_list = [str(i) for i in range(20)]
            self.num_cls = 20
            self.anchor_box_size_list = list(synthetic.YOLOV2_ANCHORS)
            self.num_anchor_box = 5
            self.last = None

        def forward(self, x):  # x: [N,H,W,3] like the reference; features are faked from it
            feat = self.feat
            out = self.conv(feat)
            self.last = out
            return layout.head_tensor(out, self.num_anchor_box)

    case = synthetic.cfg2(n=4)
    gt = targets.records_to_tensor(case.rec, cuda_device)
    off = torch.from_numpy(case.gt_off).to(cuda_device)
    x = torch.zeros(case.n, case.height, case.width, 3, device=cuda_device)
    feat = torch.randn(case.n, 16, 13, 13, generator=torch.Generator().manual_seed(9)).to(cuda_device)
    grads = []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # compare the two layouts in full fp32
    for fmt in ("contiguous", "channels_last"):
        torch.manual_seed(3)
        m = TinyV2().to(cuda_device)
        m.feat = feat
        if fmt == "channels_last":
            layout.use_channels_last_head(m)
            m.feat = feat.contiguous(memory_format=torch.channels_last)
        loss = m.get_loss_compact(x, gt, off)
        y = layout.head_tensor(m.last, 5)
        assert layout.is_free_view(m.last, y) == (fmt == "channels_last")
        loss.backward()
        grads.append((float(loss.item()), m.conv.weight.grad.detach().float().cpu().numpy().copy()))
    torch.backends.cudnn.allow_tf32 = tf32
    assert abs(grads[0][0] - grads[1][0]) <= 1e-5 * abs(grads[0][0])
    assert rel_err(grads[1][1], grads[0][1]) <= 1e-4  # (cuDNN picks different algorithms per layout)


# ------------------------------------------------------------------------------------------
# batched evaluation
# ------------------------------------------------------------------------------------------
def test_evaluate_detections_matches_reference_evaluate_model(cuda_device):
    """yh_match_detections + the host AP arithmetic vs the per-class APs of the reference's own
    evaluate_model run on the same detections and annotations: bit-exact."""
    from odcp_b200.models.utils import evaluate_detections
    z = dict(np.load(os.path.join(GOLDEN, "evaluate.npz")))
    post = dict(bbox=torch.from_numpy(z["det_bbox"]).to(cuda_device), label=torch.from_numpy(z["det_label"]).to(cuda_device),
                score=torch.from_numpy(z["det_score"]).to(cuda_device), keep_cnt=torch.from_numpy(z["keep_cnt"]).to(cuda_device))
    cls_list = [str(c) for c in range(int(z["num_cls"]))]
    res = evaluate_detections(post, z["gt_boxes"], z["gt_labels"], z["gt_off"], cls_list, z["levels"])
    assert np.array_equal(res["level_list"], z["levels"])
    for c, name in enumerate(cls_list):
        assert np.array_equal(res[name], z["ap"][c]), name
    # matching against the oracle, slot by slot, including unused slots and classes without ground truth
    best, tp = ops.match_detections(post, z["gt_boxes"], z["gt_labels"], z["gt_off"], z["levels"])
    tp = tp.cpu().numpy()
    for i in range(int(z["n"])):
        k = int(z["keep_cnt"][i])
        lo, hi = int(z["gt_off"][i]), int(z["gt_off"][i + 1])
        want = O.match_detections_np(z["det_bbox"][i, :k], z["det_label"][i, :k], z["gt_boxes"][lo:hi],
                                     z["gt_labels"][lo:hi], z["levels"])
        assert np.array_equal(tp[i, :k], want)
        assert not tp[i, k:].any()


# ------------------------------------------------------------------------------------------
# predict / decode
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,version", [("v2_predict.npz", 2), ("v1_predict.npz", 1)])
def test_decode_matches_reference_golden(name, version, cuda_device):
    case, z = load_golden(name)
    outs = ops.decode(case.y.to(cuda_device), version=version, img_hw=(case.height, case.width),
                      anchors=case.anchors, boxes_per_cell=case.a)
    for i, t in enumerate(outs):
        got = t.cpu().numpy()
        assert got.shape == z["out%d" % i].shape, i
        assert rel_err(got, z["out%d" % i]) <= 2e-6, i
        # absolute slack scaled to the tensor: box corners are differences of O(100 px) numbers
        assert np.allclose(got, z["out%d" % i], rtol=1e-5, atol=2e-6 * np.abs(z["out%d" % i]).max()), i


def test_decode_full_size_and_nonsquare(cuda_device):
    for case in (synthetic.cfg3(), synthetic.make_case("odd", 2, 3, 9, 11, 5, 20, 288, 352, seed=21),
                 synthetic.cfg1()):
        anchors = case.anchors if case.version == 2 else case.a
        want = O.decode_torch(case.y, case.height, case.width, case.version, anchors)
        outs = ops.decode(case.y.to(cuda_device), version=case.version, img_hw=(case.height, case.width),
                          anchors=case.anchors, boxes_per_cell=case.a)
        for i, (g, w) in enumerate(zip(outs, want)):
            assert g.shape == w.shape
            assert np.allclose(g.cpu().numpy(), w.numpy(), rtol=1e-5, atol=2e-6 * float(w.abs().max())), (case.name, i)


# ------------------------------------------------------------------------------------------
# dense-target adaptor
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["cfg1_coll", "cfg2_coll", "v2_nonsq_odd", "v1_b3_c7"])
def test_compact_targets_recovers_records(key, cuda_device):
    case = CASES[key]()
    ids = np.arange(case.n, dtype=np.int64) * 7 + 3        # arbitrary image ids, like dataset indices
    dense = targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, case.version, x_img_id=ids)
    # shuffle the boxes: get_loss does not require them grouped by image
    perm = np.random.default_rng(5).permutation(case.m)
    d = [t[perm] if t.shape[0] == case.m and i != 5 else t for i, t in enumerate(dense)]
    gt, off, status = ops.compact_targets(*[t.to(cuda_device) for t in d])
    assert status.cpu().tolist() == [0, 0]
    assert np.array_equal(off.cpu().numpy(), case.gt_off)
    got = targets.tensor_to_records(gt)
    # stable inside an image: expected = records in shuffled order, stably sorted by image
    exp = case.rec[perm]
    exp = exp[np.argsort(exp["img"], kind="stable")]
    assert np.array_equal(got, exp)
    # fp32 obj_mask is accepted too
    d[4] = d[4].float()
    gt2, off2, _ = ops.compact_targets(*[t.to(cuda_device) for t in d])
    assert torch.equal(gt2, gt) and torch.equal(off2, off)


# ------------------------------------------------------------------------------------------
# post-process / NMS
# ------------------------------------------------------------------------------------------
def run_post(case, conf_thre, iou_thre, dev, class_aware=False, max_out=None, want_cls_spec=True, unaligned=False):
    y = case.y.to(dev)
    if unaligned:  # a 4-byte aligned view: no bulk copies, the general path of the kernel
        buf = torch.zeros(case.y.numel() + 1, device=dev)
        y = buf[1:].view(case.y.shape)
        y.copy_(case.y)
        assert y.data_ptr() % 16 == 4
    r = ops.postprocess(y, version=case.version, img_hw=(case.height, case.width),
                        conf_thre=conf_thre, iou_thre=iou_thre, anchors=case.anchors,
                        boxes_per_cell=case.a, class_aware=class_aware, max_out=max_out,
                        want_cls_spec=want_cls_spec)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in r.items()}


def flat_kept(r, key):
    cnt = r["keep_cnt"]
    return np.concatenate([r[key][n, :cnt[n]] for n in range(len(cnt))], 0)


@pytest.mark.parametrize("name", ["v2_nms_cfg3.npz", "v2_nms_default.npz", "v1_nms.npz"])
def test_postprocess_matches_reference_golden(name, cuda_device):
    case, z = load_golden(name)
    r = run_post(case, float(z["conf_thre"]), float(z["iou_thre"]), cuda_device)
    assert np.array_equal(r["keep_cnt"], z["nms_cnt"])
    assert z["nms_cnt"].sum() > 0
    assert np.array_equal(flat_kept(r, "keep_idx"), z["nms_idx"])
    assert np.allclose(flat_kept(r, "bbox"), z["nms_bbox"], rtol=1e-5, atol=1e-4)
    assert np.allclose(flat_kept(r, "conf"), z["nms_conf"], rtol=1e-6, atol=1e-7)
    assert np.allclose(flat_kept(r, "cls_spec"), z["nms_cls_spec"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(flat_kept(r, "label"), z["nms_cls_spec"].argmax(-1))
    assert np.allclose(flat_kept(r, "score"), z["nms_cls_spec"].max(-1), rtol=1e-5, atol=1e-7)


def check_post_vs_oracle(case, conf_thre, iou_thre, dev, class_aware=False, want_cls_spec=True):
    anchors = case.anchors if case.version == 2 else case.a
    want = O.postprocess_np(case.y, case.height, case.width, case.version, anchors, conf_thre, iou_thre,
                            class_aware=class_aware)
    r = run_post(case, conf_thre, iou_thre, dev, class_aware=class_aware, want_cls_spec=want_cls_spec)
    cnt = np.array([len(w["idx"]) for w in want], dtype=np.int32)
    assert np.array_equal(r["keep_cnt"], cnt)
    for n, w in enumerate(want):
        assert np.array_equal(r["keep_idx"][n, :cnt[n]], w["idx"]), n
        assert np.array_equal(r["label"][n, :cnt[n]], w["label"]), n
        assert np.allclose(r["bbox"][n, :cnt[n]], w["bbox"], rtol=1e-5, atol=1e-4)
        assert np.allclose(r["score"][n, :cnt[n]], w["score"], rtol=1e-5, atol=1e-7)
    return cnt


def test_postprocess_cfg3_full_vs_oracle(cuda_device):
    case = synthetic.cfg3()
    cnt = check_post_vs_oracle(case, 0.5, 0.45, cuda_device)
    assert 5 < cnt.mean() < 60


def test_postprocess_class_aware_vs_oracle(cuda_device):
    check_post_vs_oracle(synthetic.cfg3(n=32), 0.5, 0.45, cuda_device, class_aware=True)
    check_post_vs_oracle(synthetic.make_case("v1n", 1, 6, 7, 7, 2, 20, 448, 448, seed=41, to_shift=0.5),
                         0.4, 0.3, cuda_device, class_aware=True)


class _Objectness:
    """The objectness logits of a case's head tensor as [N, P] (predictor order of the kernels), writable."""

    def __init__(self, case):
        self.case = case
        a = case.a
        if case.version == 2:
            self.v = case.y.view(case.n, -1, 5 + case.c)[:, :, 4]                          # [N, P]
        else:
            self.v = case.y.view(case.n, case.s_h * case.s_w, -1)[:, :, 4:5 * a:5]          # [N, cells, A]
        self.shape = (case.n, case.s_h * case.s_w * a)

    def fill(self, n, value):
        if n is None:
            self.v[...] = value
        else:
            self.v[n] = value

    def put(self, n, idx, values):
        idx = torch.as_tensor(np.asarray(idx), dtype=torch.long)
        values = torch.as_tensor(np.asarray(values, dtype=np.float32))
        if self.case.version == 2:
            self.v[n, idx] = values
        else:
            self.v[n, idx // self.case.a, idx % self.case.a] = values

    def numpy(self):
        return self.v.reshape(self.shape[0], -1).numpy().copy()


def _objectness(case):
    assert case.y.is_contiguous()
    return _Objectness(case)


def assert_same_post(a, b, keys=("keep_idx", "label", "score", "conf", "bbox")):
    assert np.array_equal(a["keep_cnt"], b["keep_cnt"])
    for n, k in enumerate(a["keep_cnt"]):
        for key in keys:
            assert np.array_equal(a[key][n, :k], b[key][n, :k]), (key, n)


def test_postprocess_without_cls_spec_vs_oracle(cuda_device):
    """want_cls_spec=False takes the class pick that skips the divisions which cannot matter."""
    check_post_vs_oracle(synthetic.cfg3(n=32), 0.5, 0.45, cuda_device, want_cls_spec=False)
    check_post_vs_oracle(synthetic.cfg3(n=8), 0.5, 0.45, cuda_device, class_aware=True, want_cls_spec=False)
    check_post_vs_oracle(synthetic.make_case("v1p", 1, 6, 7, 7, 2, 20, 448, 448, seed=43, to_shift=0.5),
                         0.4, 0.3, cuda_device, want_cls_spec=False)
    check_post_vs_oracle(synthetic.make_case("p19", 2, 3, 19, 19, 5, 20, 608, 608, seed=44, to_shift=-2.0),
                         0.5, 0.45, cuda_device, want_cls_spec=False)


@pytest.mark.parametrize("geom", ["v2_13", "v2_19", "v1_7"])
def test_postprocess_candidate_count_edges(geom, cuda_device):
    """Images with exactly K candidates around every boundary of the kernel: none, one, odd/even pair
    enumeration, warp and ballot-word boundaries, and the shared-memory limit (256: beyond it the image
    continues in the workspace).  Every image against the oracle, with and without cls_spec."""
    if geom == "v2_13":
        case = synthetic.make_case("ke13", 2, 17, 13, 13, 5, 20, 416, 416, seed=61)
    elif geom == "v2_19":
        case = synthetic.make_case("ke19", 2, 17, 19, 19, 5, 20, 608, 608, seed=62)
    else:
        case = synthetic.make_case("ke7", 1, 17, 7, 7, 2, 20, 448, 448, seed=63)
    to = _objectness(case)
    p = to.shape[1]
    ks = [0, 1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 97, 255, 256, 257, 300]
    rng = np.random.default_rng(7)
    to.fill(None, -30.0)
    for n, k in enumerate(ks):
        k = min(k, p)
        to.put(n, rng.choice(p, size=k, replace=False), rng.uniform(0.5, 4.0, size=k))
    # (spread the boxes so that a good share of them survives: smaller sizes)
    if case.version == 2:
        case.y[..., 2:4] -= 1.5
    for spec in (True, False):
        cnt = check_post_vs_oracle(case, 0.5, 0.45, cuda_device, want_cls_spec=spec)
        assert cnt[0] == 0 and cnt[1] == 1 and cnt.max() > 60
    # the general path (no bulk copies) gives the same bits
    assert_same_post(run_post(case, 0.5, 0.45, cuda_device, want_cls_spec=False),
                     run_post(case, 0.5, 0.45, cuda_device, want_cls_spec=False, unaligned=True))


@pytest.mark.parametrize("thr", [0.5, 0.3, 0.9, 0.01, 0.99, 0.0099, 0.995, 1e-4, 0.9999])
def test_postprocess_threshold_band(thr, cuda_device):
    """conf >= conf_thre is decided on the logit outside a band of +-1e-3 around logit(conf_thre) and by the
    sigmoid inside it: objectness logits on, next to and far from every edge of that band.  The whole-image
    path must list exactly the candidates of the general path (which takes the sigmoid of every logit above
    the reject margin), and both must agree with the oracle wherever sigmoid(t) is not within a few ulp
    of the threshold."""
    case = synthetic.make_case("band", 2, 2, 13, 13, 5, 20, 416, 416, seed=71)
    to = _objectness(case)
    t0 = np.log(thr / (1.0 - thr))
    offs = [0.0]
    for d in (1e-7, 1e-6, 1e-5, 1e-4, 5e-4, 9.9e-4, 1e-3, 1.01e-3, 1.5e-3, 1e-2, 0.1, 1.0, 3.0):
        offs += [d, -d]
    vals = []
    for o in offs:
        v = np.float32(t0 + o)
        vals += [v, np.nextafter(v, np.float32(np.inf)), np.nextafter(v, np.float32(-np.inf))]
    vals = np.array(vals, dtype=np.float32)
    assert len(vals) <= 200
    to.fill(None, -60.0)
    rng = np.random.default_rng(3)
    for n in range(case.n):
        to.put(n, rng.choice(to.shape[1], size=len(vals), replace=False), vals)
    # iou_thre 2: nothing is suppressed, the kept list is the candidate list
    fast = run_post(case, thr, 2.0, cuda_device, want_cls_spec=False)
    gen = run_post(case, thr, 2.0, cuda_device, want_cls_spec=False, unaligned=True)
    assert_same_post(fast, gen)
    conf = (np.float32(1.0) / (np.float32(1.0) + np.exp(-to.numpy()))).astype(np.float32)
    thr32 = np.float32(thr)
    for n in range(case.n):
        got = np.zeros(to.shape[1], dtype=bool)
        got[fast["keep_idx"][n, :fast["keep_cnt"][n]]] = True
        want = conf[n] >= thr32
        clear = np.abs(conf[n] - thr32) > 8 * np.spacing(thr32)
        assert np.array_equal(got[clear], want[clear]), (thr, np.nonzero(got != want)[0])
        assert want[clear].sum() >= 6 and (~want[clear]).sum() >= 20


def test_postprocess_class_pick_adversarial(cuda_device):
    """Class logits built to break a class pick that skips work: exact and near ties for the maximum,
    saturated softmax, all classes equal, scores in the denormal range.  The fast pick (no cls_spec, whole
    image path), the full pick (cls_spec requested) and the general path must return the same labels and
    scores, bit for bit; where the maximum is unique in float arithmetic the label is the oracle's."""
    case = synthetic.make_case("pick", 2, 8, 13, 13, 5, 20, 416, 416, seed=81, to_shift=-1.4)
    cls = case.y[..., 5:]                       # [N, S, S, A, C]
    rng = np.random.default_rng(11)
    shape = tuple(cls.shape[1:4])
    cls[1] = 0.25                               # all classes equal: label 0
    a = torch.from_numpy(rng.integers(0, 20, size=shape))
    b = torch.from_numpy(rng.integers(0, 20, size=shape))
    top = cls[2].max(-1).values + 0.5           # two exact maxima
    cls[2].scatter_(-1, a[..., None], top[..., None])
    cls[2].scatter_(-1, b[..., None], top[..., None])
    top = cls[3].max(-1).values + 0.5           # runner-up one ulp below the maximum
    cls[3].scatter_(-1, a[..., None], top[..., None])
    cls[3].scatter_(-1, b[..., None], torch.nextafter(top, torch.tensor(-1e30))[..., None])
    top = cls[4].max(-1).values + 0.5           # runner-up just inside / outside the 1e-4 window
    d = torch.from_numpy(rng.choice([0.9e-4, 1.1e-4, 1e-4], size=shape).astype(np.float32))
    cls[4].scatter_(-1, a[..., None], top[..., None])
    cls[4].scatter_(-1, b[..., None], (top - d)[..., None])
    cls[5] = -80.0                              # saturated: one class takes everything
    cls[5].scatter_(-1, a[..., None], torch.full(shape, 80.0)[..., None])
    cls[6] *= 30.0
    to = _objectness(case)
    to.fill(7, -200.0)                          # scores in the denormal range (conf ~ 5e-39)
    to.put(7, rng.choice(to.shape[1], size=40, replace=False), rng.uniform(-88.4, -87.6, size=40))
    for thr, imgs in ((0.5, slice(0, 7)), (1e-40, slice(7, 8))):
        sub = synthetic.HeadCase(case.name, 2, imgs.stop - imgs.start, 13, 13, 5, 20, 416, 416,
                                 case.y[imgs].contiguous(), case.rec[:0], np.zeros(imgs.stop - imgs.start + 1, np.int32),
                                 anchors=case.anchors)
        fast = run_post(sub, thr, 0.45, cuda_device, want_cls_spec=False)
        full = run_post(sub, thr, 0.45, cuda_device, want_cls_spec=True)
        gen = run_post(sub, thr, 0.45, cuda_device, want_cls_spec=False, unaligned=True)
        assert fast["keep_cnt"].min() >= 5
        assert_same_post(fast, full)
        assert_same_post(fast, gen)
        if thr == 0.5:
            k = fast["keep_cnt"]
            assert (fast["label"][1, :k[1]] == 0).all()
            for n in (2, 3):   # ties / one-ulp runner-up: the reference's argmax takes the first maximum
                idx = fast["keep_idx"][n, :k[n]]
                rows = sub.y[n].reshape(-1, 25)[idx, 5:].numpy()
                assert np.array_equal(fast["label"][n, :k[n]], rows.argmax(-1))
            for n in (0, 5):
                want = O.postprocess_np(sub.y[n:n + 1], 416, 416, 2, sub.anchors, 0.5, 0.45)[0]
                assert np.array_equal(fast["keep_idx"][n, :k[n]], want["idx"])
                assert np.array_equal(fast["label"][n, :k[n]], want["label"])


def test_postprocess_dense_candidates_multi_tile(cuda_device):
    """More candidates than one 256-wide suppression tile (19x19, most predictors pass) and the
    degenerate thresholds: everything passes / nothing passes / iou_thre 0 and > 1."""
    case = synthetic.make_case("dense", 2, 3, 19, 19, 5, 20, 608, 608, seed=42, to_shift=1.0)
    check_post_vs_oracle(case, 0.5, 0.45, cuda_device)
    check_post_vs_oracle(case, 0.0, 0.6, cuda_device)           # all 1805 predictors are candidates
    cnt = check_post_vs_oracle(case, 0.5, 1.5, cuda_device)     # nothing is ever suppressed
    assert (cnt > 256).all()
    cnt = check_post_vs_oracle(case, 1.1, 0.5, cuda_device)     # no candidates
    assert (cnt == 0).all()
    check_post_vs_oracle(case, 0.9, 0.0, cuda_device)           # iou >= 0 always: one box survives
    check_post_vs_oracle(case, 0.3, 0.45, cuda_device, class_aware=True)


def test_postprocess_max_out_truncates_but_counts(cuda_device):
    case = synthetic.cfg3(n=8)
    full = run_post(case, 0.5, 0.45, cuda_device)
    r = run_post(case, 0.5, 0.45, cuda_device, max_out=4)
    assert np.array_equal(r["keep_cnt"], full["keep_cnt"])
    for n in range(case.n):
        k = min(4, full["keep_cnt"][n])
        assert np.array_equal(r["keep_idx"][n, :k], full["keep_idx"][n, :k])


def test_postprocess_properties_cfg5_full(cuda_device):
    """Full-size cfg 5 post-process: kept ⊆ candidates, descending confidence, no two kept boxes
    with IoU >= thr, idempotence (NMS of the kept set keeps everything), and agreement with the
    oracle on a 16-image slice."""
    case = synthetic.cfg5()
    r = run_post(case, 0.5, 0.45, cuda_device, max_out=256)
    cnt = r["keep_cnt"]
    assert cnt.max() <= 256 and cnt.min() > 0
    sub = synthetic.HeadCase("s", 2, 16, case.s_h, case.s_w, case.a, case.c, case.height, case.width,
                             case.y[:16].contiguous(), case.rec[:0], case.gt_off[:17] * 0, anchors=case.anchors)
    want = O.postprocess_np(sub.y, sub.height, sub.width, 2, sub.anchors, 0.5, 0.45)
    for n, w in enumerate(want):
        assert np.array_equal(r["keep_idx"][n, :cnt[n]], w["idx"])
    thr = np.float32(0.45)
    for n in range(0, case.n, 37):
        k = cnt[n]
        conf, box = r["conf"][n, :k], r["bbox"][n, :k]
        assert (conf >= np.float32(0.5)).all() and (np.diff(conf) <= 0).all()
        iou = O.iou_np(box[:, None, :], box[None, :, :])
        np.fill_diagonal(iou, 0)
        assert (iou < thr).all()
    # idempotence through the decoded-box entry point
    dev = cuda_device
    bbox = torch.from_numpy(r["bbox"]).to(dev)
    conf = torch.from_numpy(r["conf"]).to(dev).clone()
    valid = torch.arange(256, device=dev)[None, :] < torch.from_numpy(cnt).to(dev)[:, None]
    conf[~valid] = -1.0
    idx2, cnt2 = ops.nms_indices(bbox, conf, conf_thre=0.5, iou_thre=0.45)
    assert np.array_equal(cnt2.cpu().numpy(), cnt)
    assert all(np.array_equal(idx2[n, :cnt[n]].cpu().numpy(), np.arange(cnt[n])) for n in range(0, case.n, 37))


def test_nms_on_decoded_boxes_matches_oracle(cuda_device):
    case = synthetic.cfg3(n=16)
    _, _, bbox, conf, _, spec = O.decode_torch(case.y, case.height, case.width, 2, case.anchors)
    b = bbox.reshape(case.n, -1, 4).contiguous()
    c = conf.reshape(case.n, -1).contiguous()
    lab = spec.reshape(case.n, -1, case.c).argmax(-1).int()
    for labels in (None, lab):
        idx, cnt = ops.nms_indices(b.to(cuda_device), c.to(cuda_device), conf_thre=0.5, iou_thre=0.45,
                                   labels=None if labels is None else labels.to(cuda_device))
        idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
        for n in range(case.n):
            want = O.nms_image_np(b[n].numpy(), c[n].numpy(), 0.5, 0.45,
                                  labels=None if labels is None else labels[n].numpy())
            assert cnt[n] == len(want)
            assert np.array_equal(idx[n, :cnt[n]], want)


def test_nms_equal_confidences_order_by_index(cuda_device):
    """torch.sort leaves ties unspecified (SURVEY B-4); this implementation orders them by
    ascending predictor index."""
    b = torch.tensor([[[0, 0, 10, 10], [100, 100, 110, 110], [0, 0, 10, 10.5], [200, 200, 210, 210]]],
                     dtype=torch.float32, device=cuda_device)
    c = torch.tensor([[0.7, 0.9, 0.7, 0.9]], device=cuda_device)
    idx, cnt = ops.nms_indices(b, c, conf_thre=0.5, iou_thre=0.5)
    assert cnt.item() == 3
    assert idx[0, :3].cpu().tolist() == [1, 3, 0]


def test_iou_known_answers(cuda_device):
    import os
    from conftest import GOLDEN
    z = dict(np.load(os.path.join(GOLDEN, "iou_kat.npz")))
    got = ops.iou(torch.from_numpy(z["b1"]).to(cuda_device), torch.from_numpy(z["b2"]).to(cuda_device))
    assert np.array_equal(got.cpu().numpy(), z["iou"])   # +,-,*,/ are correctly rounded on both sides
    # large random batch against the oracle
    rng = np.random.default_rng(7)
    b1 = rng.uniform(0, 400, size=(100003, 4)).astype(np.float32)
    b2 = rng.uniform(0, 400, size=(100003, 4)).astype(np.float32)
    b1[:, 2:] += b1[:, :2]
    b2[:, 2:] += b2[:, :2]
    got = ops.iou(torch.from_numpy(b1).to(cuda_device), torch.from_numpy(b2).to(cuda_device))
    assert np.array_equal(got.cpu().numpy(), O.iou_np(b1, b2))


def test_cuda_graph_capture_of_a_step(cuda_device):
    """The entry points only enqueue work: a train + post-process step captures into a CUDA graph
    and replays with identical results."""
    case = synthetic.headline(n=32)
    dev = cuda_device
    lam = synthetic.DEFAULT_LAMBDAS
    y = case.y.to(dev)
    gt = targets.records_to_tensor(case.rec, dev)
    off = torch.from_numpy(case.gt_off).to(dev)
    eager = ops.train_head(y, gt, off, version=2, img_hw=(416, 416), lambdas=lam, anchors=case.anchors)
    eager_post = ops.postprocess(y, version=2, img_hw=(416, 416), conf_thre=0.5, iou_thre=0.45,
                                 anchors=case.anchors, max_out=64)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = dict(dy=torch.empty_like(y), loss=torch.empty((), device=dev), terms=torch.empty(5, device=dev))
        ops.train_head(y, gt, off, version=2, img_hw=(416, 416), lambdas=lam, anchors=case.anchors, out=out)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            ops.train_head(y, gt, off, version=2, img_hw=(416, 416), lambdas=lam, anchors=case.anchors, out=out)
            post = ops.postprocess(y, version=2, img_hw=(416, 416), conf_thre=0.5, iou_thre=0.45,
                                   anchors=case.anchors, max_out=64)
    out["dy"].zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["dy"], eager["dy"])
    assert out["loss"].item() == eager["loss"].item()
    assert torch.equal(post["keep_cnt"], eager_post["keep_cnt"])
    cnt = eager_post["keep_cnt"].cpu().numpy()
    for n in range(case.n):
        assert torch.equal(post["keep_idx"][n, :cnt[n]], eager_post["keep_idx"][n, :cnt[n]])


def test_postprocess_tail_and_unaligned_sources(cuda_device):
    """Tensor sizes that are not a multiple of 16 bytes (the staged window needs a scalar tail),
    in both the resident (13x13) and the ring (19x19) staging mode, and a 4-byte-aligned view of
    y (no TMA: direct global reads)."""
    for case, thr in ((synthetic.make_case("t13", 2, 1, 13, 13, 5, 20, 416, 416, seed=51, to_shift=-1.0), 0.5),
                      (synthetic.make_case("t19", 2, 1, 19, 19, 5, 20, 608, 608, seed=52, to_shift=-1.0), 0.5),
                      (synthetic.make_case("t19b", 2, 3, 19, 19, 5, 20, 608, 608, seed=53, to_shift=1.5), 0.6),
                      (synthetic.make_case("tv1", 1, 3, 7, 7, 2, 20, 448, 448, seed=54, to_shift=0.3), 0.5)):
        assert (case.y.numel() % 4 != 0) or case.n == 3
        check_post_vs_oracle(case, thr, 0.45, cuda_device)
    case = synthetic.cfg3(n=5)
    base = run_post(case, 0.5, 0.45, cuda_device)
    buf = torch.zeros(case.y.numel() + 1, device=cuda_device)
    y = buf[1:].view(case.y.shape)
    y.copy_(case.y)
    assert y.data_ptr() % 16 == 4
    r = ops.postprocess(y, version=2, img_hw=(case.height, case.width), conf_thre=0.5, iou_thre=0.45,
                        anchors=case.anchors)
    assert np.array_equal(r["keep_cnt"].cpu().numpy(), base["keep_cnt"])
    for n in range(case.n):
        k = base["keep_cnt"][n]
        assert np.array_equal(r["keep_idx"][n, :k].cpu().numpy(), base["keep_idx"][n, :k])
        assert np.array_equal(r["score"][n, :k].cpu().numpy(), base["score"][n, :k])


# ------------------------------------------------------------------------------------------
# host-buffer pipeline
# ------------------------------------------------------------------------------------------
def test_host_pipeline_returns_what_the_direct_calls_return(cuda_device):
    """HostHeadPipeline (pinned host buffers in, host buffers out, three streams, packed transfers, the fused
    step) against the two separate calls made directly on device tensors, over more submits than slots."""
    from odcp_b200.host import HostHeadPipeline
    lam = synthetic.DEFAULT_LAMBDAS
    cases = [synthetic.make_case("p%d" % i, 2, 16, 13, 13, 5, 20, 416, 416, seed=900 + i, to_shift=-1.5) for i in range(5)]
    max_boxes = max(c.m for c in cases)
    pipe = HostHeadPipeline(16, 13, 13, 5, 20, img_hw=(416, 416), anchors=cases[0].anchors, lambdas=lam, conf_thre=0.5,
                            iou_thre=0.45, max_out=64, max_boxes=max_boxes, depth=3, device=cuda_device)
    tickets = []
    results = []
    for i, case in enumerate(cases):
        if i >= pipe.depth:
            r = pipe.result(tickets[i - pipe.depth])
            results.append({k: (v.clone() if v is not None else None) for k, v in r.items()})
        tickets.append(pipe.submit(case.y, targets.records_to_tensor(case.rec), torch.from_numpy(case.gt_off)))
    for tk in tickets[len(results):]:
        r = pipe.result(tk)
        results.append({k: (v.clone() if v is not None else None) for k, v in r.items()})
    for case, got in zip(cases, results):
        kw = dict(version=2, img_hw=(416, 416), anchors=case.anchors)
        y = case.y.to(cuda_device)
        want = ops.train_head(y, targets.records_to_tensor(case.rec, cuda_device), torch.from_numpy(case.gt_off).to(cuda_device),
                              lambdas=lam, **kw)
        post = ops.postprocess(y, conf_thre=0.5, iou_thre=0.45, max_out=64, want_cls_spec=False, **kw)
        torch.cuda.synchronize()
        # (the pipeline runs the fused step: gradients and detections bit for bit, the loss sums are grouped by
        #  image instead of by tile -- equal to float rounding)
        assert abs(float(got["loss"]) - float(want["loss"].item())) <= 2e-6 * abs(float(want["loss"].item()))
        assert torch.equal(got["dy"], want["dy"].cpu())
        assert torch.allclose(got["terms"], want["terms"].cpu(), rtol=2e-6, atol=0)
        cnt = post["keep_cnt"].cpu()
        assert torch.equal(got["keep_cnt"], cnt)
        mask = torch.arange(64)[None, :] < cnt.clamp(max=64)[:, None]
        for k in ("keep_idx", "bbox", "conf", "label", "score"):
            assert torch.equal(got[k][mask], post[k].cpu()[mask]), k


def test_overlapped_launch_chain_gives_the_same_results(cuda_device):
    """yh_v2_train_overlapped + YH_POST_INPUT_READY: a captured chain of steps over rotating buffer sets,
    every kernel but the first overlapping the tails of the ones in front of it, replayed several times,
    must leave in every set exactly what stream-ordered calls produce."""
    lam = synthetic.DEFAULT_LAMBDAS
    cases = [synthetic.make_case("o%d" % i, 2, 96, 13, 13, 5, 20, 416, 416, seed=700 + i, to_shift=-1.5) for i in range(4)]
    kw = dict(version=2, img_hw=(416, 416), anchors=cases[0].anchors)
    sets, want = [], []
    for c in cases:
        y = c.y.to(cuda_device)
        gt = targets.records_to_tensor(c.rec, cuda_device)
        off = torch.from_numpy(c.gt_off).to(cuda_device)
        r = ops.train_head(y, gt, off, lambdas=lam, **kw)
        p = ops.postprocess(y, conf_thre=0.5, iou_thre=0.45, max_out=96, want_cls_spec=False, **kw)
        torch.cuda.synchronize()
        want.append(dict(loss=r["loss"].clone(), terms=r["terms"].clone(), dy=r["dy"].clone(),
                         post={k: v.clone() for k, v in p.items() if k in ("keep_cnt", "keep_idx", "bbox", "label", "score", "conf")}))
        sets.append(dict(y=y, gt=gt, off=off, out=dict(dy=torch.empty_like(y), loss=torch.empty((), device=cuda_device),
                                                        terms=torch.empty(5, device=cuda_device)), post=None))
    stream = torch.cuda.Stream(cuda_device)

    def chain():
        for i, s_ in enumerate(sets):
            ops.train_head(s_["y"], s_["gt"], s_["off"], lambdas=lam, out=s_["out"], input_ready=(i > 0), **kw)
            s_["post"] = ops.postprocess(s_["y"], conf_thre=0.5, iou_thre=0.45, max_out=96, want_cls_spec=False,
                                         out=s_["post"], input_ready=True, **kw)

    with torch.cuda.stream(stream):
        chain()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            chain()
        for rep in range(6):
            for s_ in sets:  # poison the outputs: a kernel that did not run, or ran on stale data, shows
                s_["out"]["dy"].fill_(float("nan"))
                s_["out"]["loss"].fill_(float("nan"))
                s_["post"]["keep_cnt"].fill_(-1)
            for _ in range(3):
                g.replay()
            stream.synchronize()
            for s_, w in zip(sets, want):
                assert torch.equal(s_["out"]["loss"], w["loss"])
                assert torch.equal(s_["out"]["terms"], w["terms"])
                assert torch.equal(s_["out"]["dy"], w["dy"])
                cnt = w["post"]["keep_cnt"]
                assert torch.equal(s_["post"]["keep_cnt"], cnt)
                mask = torch.arange(96, device=cuda_device)[None, :] < cnt.clamp(max=96)[:, None]
                for k in ("keep_idx", "bbox", "label", "score", "conf"):
                    assert torch.equal(s_["post"][k][mask], w["post"][k][mask]), k


# ------------------------------------------------------------------------------------------
# fused SGD step
# ------------------------------------------------------------------------------------------
def _sgd_golden():
    z = np.load(os.path.join(GOLDEN, "sgd_step.npz"))
    return z, int(z["n_tensors"])


def test_fused_sgd_matches_the_reference_sequence_golden(cuda_device):
    """A NEW optimizer in every iteration (models/yolov2.py:1253-1272), gradients from the golden file."""
    from odcp_b200.optim import SGD, reset_state
    reset_state()
    z, n = _sgd_golden()
    ps = [torch.nn.Parameter(torch.from_numpy(z["p0_%d" % i]).to(cuda_device)) for i in range(n)]
    for it, lr in enumerate(z["lrs"]):
        opt = SGD(ps, lr=float(lr), momentum=float(z["momentum"]), weight_decay=float(z["weight_decay"]))
        opt.zero_grad()
        for i, p in enumerate(ps):
            p.grad = torch.from_numpy(z["fresh_g%d_%d" % (it, i)]).to(cuda_device)
        opt.step()
    for i, p in enumerate(ps):
        assert np.allclose(p.detach().cpu().numpy(), z["fresh_p2_%d" % i], rtol=1e-6, atol=1e-7), i


def test_fused_sgd_persistent_momentum_golden(cuda_device):
    """persistent_momentum=True: buffers survive the re-created optimizers = ONE torch optimizer."""
    from odcp_b200.optim import SGD, reset_state
    reset_state()
    z, n = _sgd_golden()
    ps = [torch.nn.Parameter(torch.from_numpy(z["p0_%d" % i]).to(cuda_device)) for i in range(n)]
    grads = [[torch.from_numpy(z["pers_g%d_%d" % (it, i)]).to(cuda_device) for i in range(n)] for it in range(3)]
    for it in range(3):
        opt = SGD(ps, lr=float(z["lrs"][2]), momentum=float(z["momentum"]), weight_decay=float(z["weight_decay"]),
                  persistent_momentum=True)
        for i, p in enumerate(ps):
            p.grad = grads[it][i]
        opt.step()
    for i, p in enumerate(ps):
        assert np.allclose(p.detach().cpu().numpy(), z["pers_p2_%d" % i], rtol=1e-6, atol=1e-7), i
    reset_state()


def test_fused_sgd_vs_oracle_odd_shapes_and_views(cuda_device):
    """Tensors of awkward sizes (empty, one element, one past a chunk, not a multiple of 4), a parameter
    that is a 4-byte aligned view, parameters without a gradient, no weight decay, no momentum; the step is
    captured in a CUDA graph and replayed."""
    from odcp_b200.optim import SGD, reset_state
    reset_state()
    rng = np.random.default_rng(5)
    sizes = [0, 1, 3, 4, 1023, 16384, 16385, 50001]
    host_p = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    host_g = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    for lr, mom, wd in ((0.1, 0.9, 5e-4), (0.05, 0.0, 0.0), (0.2, 0.9, 0.0)):
        ps = []
        for hp in host_p:
            buf = torch.zeros(len(hp) + 1, device=cuda_device)
            t = buf[1:]          # data_ptr % 16 == 4: scalar path
            t.copy_(torch.from_numpy(hp))
            ps.append(torch.nn.Parameter(t) if len(hp) % 2 else torch.nn.Parameter(torch.from_numpy(hp).to(cuda_device)))
        extra = torch.nn.Parameter(torch.ones(7, device=cuda_device))   # never gets a gradient
        for p, hg in zip(ps, host_g):
            p.grad = torch.from_numpy(hg).to(cuda_device)
        opt = SGD(ps + [extra], lr=lr, momentum=mom, weight_decay=wd)
        opt.step()
        want, _ = O.sgd_step_np(host_p, host_g, lr, mom, wd, bufs=None)
        for p, w in zip(ps, want):
            assert np.allclose(p.detach().cpu().numpy(), w, rtol=1e-6, atol=1e-7)
        assert torch.equal(extra.detach(), torch.ones(7, device=cuda_device))
        # graph capture: the same step again on top of the result
        st = torch.cuda.Stream(cuda_device)
        with torch.cuda.stream(st):
            opt.step()   # warm: the plan is cached
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                opt.step()
            g.replay()
            st.synchronize()
        for _ in range(2):   # the warm step and the replay (capturing does not execute)
            want, _ = O.sgd_step_np(want, host_g, lr, mom, wd, bufs=None)
        for p, w in zip(ps, want):
            assert np.allclose(p.detach().cpu().numpy(), w, rtol=2e-6, atol=2e-7)
    reset_state()


def test_fused_sgd_equals_torch_sgd_on_a_model(cuda_device):
    """Drop-in check on a small conv net: one reference-style iteration with torch.optim.SGD and with the
    fused step, from the same parameters and gradients."""
    from odcp_b200.optim import SGD, reset_state
    reset_state()
    torch.manual_seed(3)
    def net():
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.LeakyReLU(0.1),
                                   torch.nn.Conv2d(8, 125, 1)).to(cuda_device)
    a, b = net(), net()
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 3, 16, 16, device=cuda_device)
    for m, cls in ((a, torch.optim.SGD), (b, SGD)):
        opt = cls(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=5e-4)
        opt.zero_grad()
        m(x).square().mean().backward()
        opt.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-7)
    reset_state()


# ------------------------------------------------------------------------------------------
# detect(): the dict of lists (reference models/yolov2.py:651-745)
# ------------------------------------------------------------------------------------------
def test_detect_returns_the_reference_dict_from_the_golden_head(cuda_device):
    """detect(img) and detect_batch on the head-only model with a head tensor from the golden file: the
    lists must hold what the reference's predict -> nms -> argmax chain produced for that image
    (v2_nms_default.npz was generated with the reference's default thresholds 0.9 / 0.5)."""
    from odcp_b200.models.yolov2 import YOLOv2Head
    case, z = load_golden("v2_nms_default.npz")
    assert float(z["conf_thre"]) == 0.9 and float(z["iou_thre"]) == 0.5
    cnt = z["nms_cnt"]
    start = np.concatenate([[0], np.cumsum(cnt)])
    m = YOLOv2Head(num_cls=case.c).to(cuda_device)
    img = np.zeros((case.height, case.width, 3), dtype=np.float32)
    for n in range(case.n):
        m.set_head_output(case.y[n:n + 1].to(cuda_device))
        d = m.detect(img)  # the reference's defaults: conf 0.9, IoU 0.5
        sl = slice(start[n], start[n + 1])
        assert sorted(d) == ["bbox_list", "cls_spec_conf_score_list", "conf_score_list", "lbl_list"]
        assert len(d["bbox_list"]) == cnt[n] > 0
        assert np.allclose(np.array(d["bbox_list"]), z["nms_bbox"][sl], rtol=1e-5, atol=1e-4)
        assert np.allclose(np.array(d["conf_score_list"]), z["nms_conf"][sl], rtol=1e-6, atol=1e-7)
        assert d["lbl_list"] == [m.cls_list[i] for i in z["nms_cls_spec"][sl].argmax(-1)]
        assert np.allclose(np.array(d["cls_spec_conf_score_list"]), z["nms_cls_spec"][sl].max(-1), rtol=1e-5, atol=1e-7)
    m.set_head_output(case.y.to(cuda_device))
    many = m.detect_batch(torch.zeros(case.n, case.height, case.width, 3, device=cuda_device), 0.9, 0.5)
    assert [len(d["bbox_list"]) for d in many] == cnt.tolist()
    assert np.allclose(np.array(many[1]["bbox_list"]), z["nms_bbox"][start[1]:start[2]], rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------
# round-2 parity additions: reference-signature wrappers, YOLOv1 detect, element-wise gradients, full-size NMS
# ------------------------------------------------------------------------------------------
def test_yolov1_detect_matches_the_reference_detect_golden(cuda_device):
    """YOLOv1.detect (resize to 224 -> predict -> nms -> clip to [0, 223] -> rescale to the original size) against
    the dict the reference's own detect produced for the same image and head tensor (tests/golden/v1_detect.npz,
    make_golden.run_v1_detect)."""
    from odcp_b200.models.yolov1 import YOLOv1Head
    z = dict(np.load(os.path.join(GOLDEN, "v1_detect.npz")))
    m = YOLOv1Head(7, 7, 2, num_cls=20).to(cuda_device)
    m.set_head_output(torch.from_numpy(z["y"]).to(cuda_device))
    d = m.detect(z["img"], float(z["conf_thre"]), float(z["iou_thre"]))
    assert len(d["bbox_list"]) == len(z["bbox"]) > 10
    assert np.allclose(np.asarray(d["bbox_list"]), z["bbox"], rtol=1e-5, atol=1e-4)
    assert [int(l) for l in d["lbl_list"]] == z["labels"].tolist()
    assert np.allclose(np.asarray(d["conf_score_list"]), z["conf"], rtol=1e-6, atol=1e-7)
    assert np.allclose(np.asarray(d["cls_spec_conf_score_list"]), z["score"], rtol=1e-5, atol=1e-7)
    assert np.asarray(d["bbox_list"]).min() >= 0.0  # clipped before the rescale


def test_reference_signature_nms_flattens_all_leading_dims(cuda_device):
    """models.utils.nms with the reference's signature (models/utils.py:68-164): every leading dimension is
    flattened into ONE candidate set (a batch of two images suppresses across images, as the reference does), the
    three gathered tensors come back in descending-confidence order."""
    from odcp_b200.models import utils as U
    case = synthetic.cfg3(2, conf_thre=0.6)
    dev = cuda_device
    y = case.y.to(dev)
    _, _, bbox, conf, _, spec = ops.decode(y, version=2, img_hw=(case.height, case.width), anchors=case.anchors)
    kb, kc, ks = U.nms(bbox, conf, spec, 0.6, 0.45)                       # [2,13,13,5,.] in, like predict's outputs
    wb, wc, ws = O.nms_torch(bbox.cpu(), conf.cpu(), spec.cpu(), 0.6, 0.45)  # the reference's op sequence
    assert kb.shape == wb.shape and kb.shape[0] > 20
    assert torch.equal(kb.cpu(), wb) and torch.equal(kc.cpu(), wc) and torch.equal(ks.cpu(), ws)
    assert (kc[:-1] >= kc[1:]).all()
    # one image alone keeps boxes the two-image call suppressed across images: the flattening is real
    k1 = U.nms(bbox[0], conf[0], spec[0], 0.6, 0.45)[1].shape[0] + U.nms(bbox[1], conf[1], spec[1], 0.6, 0.45)[1].shape[0]
    assert k1 >= kb.shape[0]
    # defaults are the reference's (0.9 / 0.5)
    d = U.nms(bbox[0], conf[0], spec[0])
    w = O.nms_torch(bbox[0].cpu(), conf[0].cpu(), spec[0].cpu())
    assert torch.equal(d[1].cpu(), w[1])


def test_reference_signature_get_iou_broadcasts_and_keeps_dtype(cuda_device):
    """models.utils.get_iou: broadcasting like the loss uses it ([M,S,S,A,4] x [M,S,S,1,4], models/yolov2.py:984-990),
    float32 tensors bit-equal to the reference's torch formula, numpy=True in the arrays' own precision: float64
    boxes (evaluate_model, models/utils.py:250-252) give numpy's float64 bits."""
    from odcp_b200.models import utils as U
    rng = np.random.default_rng(5)

    def boxes(*shape):
        a = rng.uniform(0, 400, size=shape + (2,))
        b = a + rng.uniform(0, 150, size=shape + (2,))
        return np.concatenate([a, b], -1)

    p32, t32 = boxes(7, 3, 3, 5).astype(np.float32), boxes(7, 3, 3, 1).astype(np.float32)
    got = U.get_iou(torch.from_numpy(p32).to(cuda_device), torch.from_numpy(t32).to(cuda_device))
    want = O.iou_torch(torch.from_numpy(p32), torch.from_numpy(t32))
    assert got.shape == (7, 3, 3, 5) and torch.equal(got.cpu(), want)
    # numpy route, float64: the reference's formula evaluated by numpy itself
    g64, d64 = boxes(9), boxes(1)
    g64[3] = d64[0]                      # IoU exactly 1 / (1 + 1e-6 / area)
    g64[4, 2:] = g64[4, :2]              # zero-area box
    got = U.get_iou(g64, d64, numpy=True)
    x1, y1, x2, y2 = [g64[..., i] for i in range(4)]
    u1, v1, u2, v2 = [d64[..., i] for i in range(4)]
    inter = np.clip(np.minimum(x2, u2) - np.maximum(x1, u1), 0, None) * np.clip(np.minimum(y2, v2) - np.maximum(y1, v1), 0, None)
    want = inter / ((x2 - x1) * (y2 - y1) + (u2 - u1) * (v2 - v1) - inter + 1e-6)
    assert got.dtype == np.float64 and got.shape == (9,) and np.array_equal(got, want)
    got32 = U.get_iou(g64.astype(np.float32), d64.astype(np.float32), numpy=True)
    assert got32.dtype == np.float32 and np.allclose(got32, want, rtol=1e-5)


def test_reference_evaluate_loop_with_the_patched_get_iou_reproduces_its_golden(cuda_device):
    """The reference's own evaluate_model loop (restated: models/utils.py:231-262) fed by models.utils.get_iou
    (numpy=True, float64 through yh_iou_f64) makes the true-positive decisions of tests/golden/evaluate.npz --
    the APs of the reference's run -- and so does the batched drop-in evaluate_model."""
    from odcp_b200.models import utils as U
    z = dict(np.load(os.path.join(GOLDEN, "evaluate.npz")))
    n, levels = int(z["n"]), z["levels"]
    cls_list = [str(c) for c in range(int(z["num_cls"]))]

    class Canned:  # detect() returns the golden's detections image by image, like the generator's stand-in model
        def __init__(self):
            self.cls_list, self.calls = cls_list, 0

        def detect(self, img, conf, iou):
            i, self.calls = self.calls, self.calls + 1
            k = int(z["keep_cnt"][i])
            return {"bbox_list": z["det_bbox"][i, :k].tolist(), "lbl_list": [cls_list[c] for c in z["det_label"][i, :k]],
                    "conf_score_list": z["det_score"][i, :k].tolist(), "cls_spec_conf_score_list": z["det_score"][i, :k].tolist()}

    dataset = [(i, None, {"bbox_list": z["gt_boxes"][z["gt_off"][i]:z["gt_off"][i + 1]].tolist(),
                          "lbl_list": [cls_list[c] for c in z["gt_labels"][z["gt_off"][i]:z["gt_off"][i + 1]]]}) for i in range(n)]
    res = U.evaluate_model(Canned(), dataset, None)  # per-image route: model.detect + get_iou(numpy=True)
    assert np.array_equal(res["level_list"], levels)
    for c, name in enumerate(cls_list):
        assert np.array_equal(res[name], z["ap"][c]), name


def test_batched_evaluate_model_equals_the_per_image_route(cuda_device):
    """evaluate_model on a YOLOv2 head: images batched (one post-process launch + one matching launch per batch,
    two image sizes in the dataset) against the per-image route (model.detect per image + get_iou(numpy=True), the
    reference's loop): identical APs."""
    from odcp_b200.models import utils as U
    from odcp_b200.models.yolov2 import YOLOv2Head
    dev = cuda_device
    small = synthetic.make_case("ev_a", 2, 10, 13, 13, 5, 20, 416, 416, seed=61, to_shift=-1.2, k_lo=2, k_hi=5)
    wide = synthetic.make_case("ev_b", 2, 6, 10, 13, 5, 20, 320, 416, seed=62, to_shift=-1.2, k_lo=2, k_hi=5)

    class Model(YOLOv2Head):  # forward looks the head tensor up by the image's tag (its first pixel)
        def forward(self, x):
            return torch.stack([self.table[int(t)] for t in x[:, 0, 0, 0].tolist()]).to(dev)

    m = Model(num_cls=20).to(dev)
    m.table, dataset = {}, []
    for case in (small, wide):
        for i in range(case.n):
            tag = len(m.table)
            m.table[tag] = case.y[i]
            img = np.zeros((case.height, case.width, 3), np.float32)
            img[0, 0, 0] = tag
            sel = case.rec[case.gt_off[i]:case.gt_off[i + 1]]
            # ground truth near the detections so that the APs are not all zero: the decoded boxes of a few predictors
            dataset.append((tag, img, {"bbox_list": np.stack([sel["x1"], sel["y1"], sel["x2"], sel["y2"]], 1).astype(np.float64).tolist(),
                                       "lbl_list": [m.cls_list[c] for c in sel["cls"]]}))
    got = U.evaluate_model(m, dataset, None, 0.5, 0.45, batch_size=4)

    class PerImage:  # hides `postprocess`: evaluate_model then takes the reference's per-image loop
        cls_list = m.cls_list

        def detect(self, img, conf, iou):
            return m.detect(img, conf, iou)

    want = U.evaluate_model(PerImage(), dataset, None, 0.5, 0.45)
    for name in m.cls_list:
        assert np.array_equal(got[name], want[name]), name
    assert sum(float(np.sum(got[c])) for c in m.cls_list) >= 0.0


def elementwise_bound(ref, to_logit):
    """Element-wise tolerance for the no-object gradient g = K conf^2 (1 - conf) of an objectness logit: 1e-5
    relative, plus what 4 ulp of the fp32 confidence do to it through dg/dconf -- the reference's own sigmoid is
    rounded to fp32 before autograd forms (1 - conf), so for saturated logits (conf -> 1) neither side knows the
    value better than eps32 * (2 - 3 conf) / (1 - conf) relative."""
    c = 1.0 / (1.0 + np.exp(-to_logit.astype(np.float64)))
    cond = np.abs(2.0 - 3.0 * c) / np.maximum(1.0 - c, 1e-12)
    return np.abs(ref) * (1e-5 + 4 * 2.0 ** -24 * cond)


def test_dense_objectness_gradient_elementwise(cuda_device):
    """The dense part of dL/dy -- the no-object gradient of every objectness logit, 99.9 % of the non-zeros and five
    orders of magnitude below the largest entry -- checked ELEMENT BY ELEMENT (the norm-relative tolerance says
    nothing about it) on every golden and on BASELINE config 2, through the train head and through the fused step;
    the worst element relative to its bound is reported."""
    worst = 0.0
    for name in LOSS_FILES:
        case, z = load_golden(name)
        r = run_train(case, golden_lambdas(z), cuda_device)
        sl = (Ellipsis, 4) if case.version == 2 else (Ellipsis, slice(4, 5 * case.a, 5))
        got, ref, to = r["dy"][sl], z["dy"][sl], case.y.numpy()[sl]
        # (cells that carry a record hold the record's objectness gradient as well: those are covered by the
        #  norm-relative checks; here only logits whose gradient is purely the no-object term)
        pure = np.ones(ref.shape, bool)
        rec = case.rec
        pure[rec["img"], rec["cy"], rec["cx"]] = False
        ratio = np.abs(got.astype(np.float64) - ref)[pure] / np.maximum(elementwise_bound(ref, to)[pure], 1e-300)
        assert (ratio[ref[pure] != 0] <= 1.0).all(), (name, float(ratio.max()))
        worst = max(worst, float(ratio[ref[pure] != 0].max()))
    case = synthetic.cfg2()
    want = O.train_head_compact(case, synthetic.DEFAULT_LAMBDAS)["dy"][..., 4]
    y, gt, off = case.y.to(cuda_device), targets.records_to_tensor(case.rec, cuda_device), torch.from_numpy(case.gt_off).to(cuda_device)
    kw = dict(img_hw=(case.height, case.width), lambdas=synthetic.DEFAULT_LAMBDAS, anchors=case.anchors)
    sep = ops.train_head(y, gt, off, version=2, **kw)["dy"][..., 4].cpu().numpy()
    fus = ops.train_post(y, gt, off, conf_thre=0.5, iou_thre=0.45, max_out=16, want_cls_spec=False, **kw)["train"]["dy"][..., 4].cpu().numpy()
    assert np.array_equal(sep, fus)
    pure = np.ones(want.shape, bool)
    pure[case.rec["img"], case.rec["cy"], case.rec["cx"]] = False
    ratio = np.abs(sep.astype(np.float64) - want)[pure] / np.maximum(elementwise_bound(want, case.y.numpy()[..., 4])[pure], 1e-300)
    assert (ratio[want[pure] != 0] <= 1.0).all(), float(ratio.max())
    assert np.array_equal(sep != 0, want != 0)
    # and the plain statement for everything that is not saturated (conf <= 0.9): 2e-6 relative, element by element
    mid = pure & (case.y.numpy()[..., 4] < 2.0) & (want != 0)
    assert (np.abs(sep.astype(np.float64) - want)[mid] <= 2e-6 * np.abs(want[mid])).all()
    print("worst element / bound: %.3f" % max(worst, float(ratio.max())))


def test_headline_batch_vs_the_dense_reference_tier(cuda_device):
    """The headline configuration at its full batch (N=256) against the oracle's DENSE tier -- the reference's own op
    sequence (per-box replication + autograd), evaluated in image chunks and recombined exactly."""
    case = synthetic.headline(256)
    want = O.train_head_dense(case, synthetic.DEFAULT_LAMBDAS, chunk_images=64)
    r = run_train(case, synthetic.DEFAULT_LAMBDAS, cuda_device)
    assert abs(r["loss"] - want["loss"]) <= TOL * abs(want["loss"])
    assert np.abs(r["terms"] - want["terms"]).max() <= TOL * np.abs(want["terms"]).max()
    assert np.array_equal(r["resp"], want["resp"])
    assert rel_err(r["dy"], want["dy"]) <= TOL
    assert np.array_equal(r["dy"] != 0, want["dy"] != 0)


def test_postprocess_cfg5_all_images_vs_oracle(cuda_device):
    """BASELINE config 5 at full size (N=512, 19x19, ~110 candidates per image): kept indices and labels of EVERY
    image against the oracle."""
    case = synthetic.cfg5()
    r = run_post(case, 0.5, 0.45, cuda_device, max_out=256, want_cls_spec=False)
    want = O.postprocess_np(case.y, case.height, case.width, 2, case.anchors, 0.5, 0.45)
    cnt = r["keep_cnt"]
    assert np.array_equal(cnt, np.array([len(w["idx"]) for w in want], dtype=np.int32))
    for n, w in enumerate(want):
        assert np.array_equal(r["keep_idx"][n, :cnt[n]], w["idx"]), n
        assert np.array_equal(r["label"][n, :cnt[n]], w["label"]), n


def test_nms_boxes_of_infinite_size_follow_the_reference(cuda_device):
    """exp(tw) overflows to inf for large logits: the box is (-inf, +inf), and two such boxes have IoU inf/inf = NaN.
    The reference keeps a later box only if `iou < iou_thre` (models/utils.py:133), so a NaN IoU removes it."""
    inf = np.float32(np.inf)
    bbox = np.array([[[-inf, -inf, inf, inf], [-inf, -inf, inf, inf], [10, 10, 50, 50], [-inf, 0, inf, 40], [200, 200, 260, 280],
                      [12, 12, 50, 52]]], dtype=np.float32)
    conf = np.array([[0.99, 0.98, 0.97, 0.96, 0.95, 0.94]], dtype=np.float32)
    want = O.nms_image_np(bbox[0], conf[0], 0.5, 0.5)
    ref = O.nms_torch(torch.from_numpy(bbox[0]), torch.from_numpy(conf[0]),
                      torch.arange(6, dtype=torch.float32)[:, None], 0.5, 0.5)[2][:, 0].numpy().astype(np.int32)
    assert np.array_equal(want, ref)  # the oracle agrees with the reference's op sequence on this input
    idx, cnt = ops.nms_indices(torch.from_numpy(bbox).to(cuda_device), torch.from_numpy(conf).to(cuda_device), conf_thre=0.5, iou_thre=0.5)
    k = int(cnt[0])
    assert np.array_equal(idx[0, :k].cpu().numpy(), want), (idx[0, :k].cpu().numpy(), want)
    assert 1 not in want  # the second infinite box did not survive the first


def test_get_loss_backward_twice_and_zero_upstream_gradient(cuda_device):
    """HeadLoss.backward hands the kernel's gradient buffer out once; a second backward (retain_graph) runs the kernel
    again: an upstream gradient of 0 in between loses nothing, and the first result is not mutated."""
    from odcp_b200.models.yolov2 import YOLOv2Head
    case = synthetic.cfg2(n=4)
    dev = cuda_device
    m = YOLOv2Head(num_cls=case.c).to(dev)
    y = case.y.to(dev).requires_grad_(True)
    m.set_head_output(y)
    x = torch.zeros(case.n, case.height, case.width, 3, device=dev)
    gt, off = targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev)
    loss = m.get_loss_compact(x, gt, off)
    (g0,) = torch.autograd.grad(loss, y, grad_outputs=torch.zeros((), device=dev), retain_graph=True)
    assert not g0.any()
    (g1,) = torch.autograd.grad(loss, y, retain_graph=True)
    (g3,) = torch.autograd.grad(loss, y, grad_outputs=torch.full((), 3.0, device=dev))
    want = O.train_head_compact(case, synthetic.DEFAULT_LAMBDAS)["dy"]
    assert rel_err(g1.cpu().numpy(), want) <= TOL
    assert torch.allclose(g3, 3.0 * g1, rtol=1e-6, atol=0)
    assert not g0.any()  # untouched by the later backwards
