"""The fused step (yh_v2_train_post: train head + post-process in ONE kernel, head tensor read once) against
  (1) the two separate calls: dL/dy, assignments, kept boxes, labels, scores bit for bit, loss/terms to rounding,
  (2) the CPU oracle and the reference's golden vectors,
on inputs that take every route through it: the lean path, images with more than 256 candidates (general path
inside the fused kernel), record-dense images (more records than the shared-memory record stage, collisions),
images whose start is not 16-byte aligned (every odd image of a 13x13 grid), inputs the fused kernel does not
cover (unaligned tensors, fewer than 3 classes), compile-time and run-time geometry.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from odcp_b200 import ops, synthetic, targets
from oracle import yolo_head_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
LAM = synthetic.DEFAULT_LAMBDAS


def dev_inputs(case, dev):
    return (case.y.to(dev), targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev))


def separate(case, dev, conf, iou, y=None, **kw):
    y0, gt, off = dev_inputs(case, dev)
    y = y0 if y is None else y
    tr = ops.train_head(y, gt, off, version=2, img_hw=(case.height, case.width), lambdas=LAM, anchors=case.anchors,
                        want_resp=True, want_grad=kw.get("want_grad", True))
    po = ops.postprocess(y, version=2, img_hw=(case.height, case.width), conf_thre=conf, iou_thre=iou,
                         anchors=case.anchors, class_aware=kw.get("class_aware", False), max_out=kw.get("max_out"),
                         want_cls_spec=kw.get("want_cls_spec", True))
    torch.cuda.synchronize()
    return tr, po


def fused(case, dev, conf, iou, y=None, out=None, overlapped=False, **kw):
    y0, gt, off = dev_inputs(case, dev)
    y = y0 if y is None else y
    r = ops.train_post(y, gt, off, img_hw=(case.height, case.width), lambdas=LAM, anchors=case.anchors, conf_thre=conf,
                       iou_thre=iou, want_resp=True, want_grad=kw.get("want_grad", True),
                       class_aware=kw.get("class_aware", False), max_out=kw.get("max_out"),
                       want_cls_spec=kw.get("want_cls_spec", True), out=out, overlapped=overlapped)
    torch.cuda.synchronize()
    return r


def assert_same_train(a, b):
    """Decisions and gradients bit for bit; loss and terms to float rounding (the fused kernel groups the partial
    sums by image, the train head by tile; the fixed-point totals themselves are exact)."""
    for k in ("resp", "iou_resp"):
        assert torch.equal(a[k], b[k]), k
    for k in ("loss", "terms"):
        x, y = a[k].double().cpu().numpy(), b[k].double().cpu().numpy()
        assert np.all(np.abs(x - y) <= 2e-6 * np.maximum(np.abs(x), 1e-30)), (k, x, y)
    if a["dy"] is None:
        assert b["dy"] is None
    else:
        assert torch.equal(a["dy"], b["dy"]), "dL/dy differs"


def assert_same_post(a, b):
    cnt = a["keep_cnt"].cpu().numpy()
    assert np.array_equal(cnt, b["keep_cnt"].cpu().numpy())
    max_out = a["keep_idx"].shape[1]
    valid = torch.from_numpy(np.arange(max_out)[None, :] < np.minimum(cnt, max_out)[:, None]).to(a["keep_idx"].device)
    for k in ("keep_idx", "conf", "label", "score"):
        assert torch.equal(a[k][valid], b[k][valid]), k
    assert torch.equal(a["bbox"][valid], b["bbox"][valid]), "bbox"
    if a["cls_spec"] is not None:
        assert torch.equal(a["cls_spec"][valid], b["cls_spec"][valid]), "cls_spec"


CASES = {
    # (case, conf, iou, what it exercises)
    "headline_256": (lambda: synthetic.headline(256), 0.5, 0.45),
    "headline_64": (lambda: synthetic.headline(64), 0.5, 0.45),
    "headline_8_small_tiles": (lambda: synthetic.headline(8), 0.5, 0.45),         # tiles of a few cells: many lists per image
    "one_image": (lambda: synthetic.make_case("one", 2, 1, 13, 13, 5, 20, 416, 416, seed=41, to_shift=-1.5), 0.5, 0.45),
    "cfg3_default_thresholds": (lambda: synthetic.cfg3(64, conf_thre=0.9), 0.9, 0.5),
    "cfg2_training_logits_overflow": (lambda: synthetic.cfg2(), 0.5, 0.45),       # ~420 candidates per image: fallback
    "tile_overflow_only": (lambda: synthetic.make_case("tov", 2, 256, 13, 13, 5, 20, 416, 416, seed=42, to_shift=-0.8), 0.5, 0.45),
    "cfg5_n64": (lambda: synthetic.cfg5(64), 0.5, 0.45),
    "cfg5_n512_multi_tile": (lambda: synthetic.cfg5(512), 0.5, 0.45),
    "coco_c80_a3_runtime_geometry": (lambda: synthetic.make_case("coco", 2, 40, 13, 13, 3, 80, 416, 416, seed=43, to_shift=-1.5,
                                                                 anchors=synthetic.YOLOV2_ANCHORS[:3]), 0.5, 0.45),
    "c3_nonsquare": (lambda: synthetic.make_case("odd", 2, 37, 9, 11, 5, 3, 288, 352, seed=44, to_shift=-1.2, k_hi=6), 0.5, 0.45),
    "nothing_passes": (lambda: synthetic.headline(32), 0.999999, 0.45),
    "everything_passes": (lambda: synthetic.headline(16), 0.0, 0.45),
    "tiny_grid": (lambda: synthetic.make_case("tiny", 2, 2048, 2, 2, 5, 20, 64, 64, seed=45, to_shift=-1.5), 0.5, 0.45),
    "collisions": (lambda: synthetic.with_collisions(synthetic.headline(64), 200, seed=7), 0.5, 0.45),
    "images_without_boxes": (lambda: synthetic.make_case("empty", 2, 33, 13, 13, 5, 20, 416, 416, seed=46, k_lo=0, k_hi=2,
                                                         to_shift=-1.563), 0.5, 0.45),
    # 6-14 boxes per image: the fused kernel takes four records per warp (more than 8 per image) where the train head
    # uses the one-warp form throughout (patches in the shadow of the stream, record-dense tiles after it)
    "mid_density_both_record_forms": (lambda: synthetic.make_case("mid", 2, 96, 13, 13, 5, 20, 416, 416, seed=48, k_lo=6, k_hi=14,
                                                                  to_shift=-1.563), 0.5, 0.45),
    "mid_density_collisions": (lambda: synthetic.with_collisions(synthetic.make_case("midc", 2, 48, 13, 13, 5, 20, 416, 416, seed=49,
                                                                                     k_lo=6, k_hi=14, to_shift=-1.563), 300, seed=8), 0.5, 0.45),
    # 5 + C > 32: the train head cannot patch, all its records are processed after the stream; the fused kernel's images
    # with at most 8 records take the one-warp form, the others four per warp -- C = 40 and 80 run the class loops
    # beyond the registers
    "c40_a4_sparse_records": (lambda: synthetic.make_case("c40", 2, 24, 13, 13, 4, 40, 416, 416, seed=50, to_shift=-1.5,
                                                           anchors=synthetic.YOLOV2_ANCHORS[:4]), 0.5, 0.45),
    "coco_c80_a3_dense_records": (lambda: synthetic.make_case("cocod", 2, 24, 13, 13, 3, 80, 416, 416, seed=51, to_shift=-1.5,
                                                               k_lo=10, k_hi=30, anchors=synthetic.YOLOV2_ANCHORS[:3]), 0.5, 0.45),
    # more than 12 records per tile on average: the train head takes every record of such a tile after its dense pass
    # (one warp per record, grouped by tile) where the fused kernel takes four per warp, grouped by image
    "dense_19x19_record_dense_tiles": (lambda: synthetic.make_case("d19", 2, 64, 19, 19, 5, 20, 608, 608, seed=52, k_lo=85,
                                                                         k_hi=110, to_shift=-1.59), 0.5, 0.45),
    "dense_collisions_record_dense_tiles": (lambda: synthetic.with_collisions(
        synthetic.make_case("d13", 2, 40, 13, 13, 5, 20, 416, 416, seed=53, k_lo=80, k_hi=110, to_shift=-1.563), 2500, seed=9), 0.5, 0.45),
    "dense_c80_a3_record_dense_tiles": (lambda: synthetic.make_case("dcoco", 2, 32, 13, 13, 3, 80, 416, 416, seed=54,
                                                                          to_shift=-1.5, k_lo=130, k_hi=160,
                                                                          anchors=synthetic.YOLOV2_ANCHORS[:3]), 0.5, 0.45),
    "two_classes_unfused_route": (lambda: synthetic.make_case("c2", 2, 12, 13, 13, 5, 2, 416, 416, seed=47, to_shift=-1.5), 0.5, 0.45),
}


@pytest.mark.parametrize("key", sorted(CASES))
def test_fused_step_equals_separate_calls(key, cuda_device):
    mk, conf, iou = CASES[key]
    case = mk()
    tr, po = separate(case, cuda_device, conf, iou)
    r = fused(case, cuda_device, conf, iou)
    assert_same_train(tr, r["train"])
    assert_same_post(po, r["post"])
    if key == "headline_256":  # sanity: this case is the lean path (no list overflows) with ~50 candidates per image
        cnt = r["post"]["keep_cnt"].cpu().numpy()
        assert cnt.min() > 5 and cnt.max() < 128


@pytest.mark.parametrize("opts", [dict(class_aware=True), dict(want_cls_spec=False), dict(want_grad=False),
                                  dict(max_out=7), dict(max_out=0)])
def test_fused_step_options(opts, cuda_device):
    case = synthetic.headline(48)
    tr, po = separate(case, cuda_device, 0.5, 0.45, **opts)
    r = fused(case, cuda_device, 0.5, 0.45, **opts)
    assert_same_train(tr, r["train"])
    if opts.get("max_out") == 0:
        assert torch.equal(po["keep_cnt"], r["post"]["keep_cnt"])
    else:
        assert_same_post(po, r["post"])


def test_fused_step_unaligned_tensor_takes_the_separate_kernels(cuda_device):
    case = synthetic.headline(16)
    buf = torch.zeros(case.y.numel() + 8, device=cuda_device)
    y = buf[1:1 + case.y.numel()].view(case.y.shape)
    y.copy_(case.y.to(cuda_device))
    assert y.data_ptr() % 16 != 0
    tr, po = separate(case, cuda_device, 0.5, 0.45, y=y)
    r = fused(case, cuda_device, 0.5, 0.45, y=y)
    assert_same_train(tr, r["train"])
    assert_same_post(po, r["post"])


def test_fused_headline_batch_vs_oracle(cuda_device):
    """The configuration the north-star target is quoted on (N=256), fused step against the oracle: loss, terms,
    assignments, the whole gradient, and the kept boxes of every image."""
    case = synthetic.headline(256)
    r = fused(case, cuda_device, 0.5, 0.45)
    want = O.train_head_compact(case, LAM)
    t = r["train"]
    assert abs(float(t["loss"]) - want["loss"]) <= TOL * abs(want["loss"])
    assert np.abs(t["terms"].cpu().numpy() - want["terms"]).max() <= TOL * np.abs(want["terms"]).max()
    assert np.array_equal(t["resp"].cpu().numpy(), want["resp"])
    dy = t["dy"].cpu().numpy()
    assert rel_err(dy, want["dy"]) <= TOL
    assert np.array_equal(dy != 0, want["dy"] != 0)
    wp = O.postprocess_np(case.y, case.height, case.width, 2, case.anchors, 0.5, 0.45)
    cnt = r["post"]["keep_cnt"].cpu().numpy()
    idx = r["post"]["keep_idx"].cpu().numpy()
    lab = r["post"]["label"].cpu().numpy()
    for n, w in enumerate(wp):
        assert cnt[n] == len(w["idx"]) and np.array_equal(idx[n, :cnt[n]], w["idx"]), n
        assert np.array_equal(lab[n, :cnt[n]], w["label"]), n


def test_fused_step_matches_reference_nms_golden(cuda_device):
    """Kept indices of the reference's own nms (tests/golden/v2_nms_cfg3.npz) through the fused step."""
    case, z = load_golden("v2_nms_cfg3.npz")
    conf, iou = float(z["conf_thre"]), float(z["iou_thre"])
    if case.rec is None or len(case.rec) == 0:
        pytest.skip("golden has no ground truth")
    r = fused(case, cuda_device, conf, iou)
    cnt = r["post"]["keep_cnt"].cpu().numpy()
    idx = r["post"]["keep_idx"].cpu().numpy()
    assert np.array_equal(cnt, z["nms_cnt"])
    off = np.concatenate([[0], np.cumsum(z["nms_cnt"])])
    for n in range(case.n):
        want = z["nms_idx"][off[n]:off[n + 1]]
        assert np.array_equal(idx[n, :cnt[n]], want), n
        got_box = r["post"]["bbox"][n, :cnt[n]].cpu().numpy()
        assert np.allclose(got_box, z["nms_bbox"][off[n]:off[n + 1]], rtol=1e-5, atol=1e-4), n


def test_fused_overlapped_chain_with_different_data_per_step(cuda_device):
    """A chain of overlapped fused steps over rotating buffer sets, every step on DIFFERENT data, replayed twice
    from a CUDA graph over poisoned outputs: every step's results equal the stream-ordered ones (a stale or
    overtaken write from a neighbouring step would show)."""
    dev = cuda_device
    R = 6
    cases = [synthetic.make_case("chain%d" % i, 2, 96, 13, 13, 5, 20, 416, 416, seed=300 + i, to_shift=-1.563) for i in range(R)]
    want = []
    for c in cases:
        r = fused(c, dev, 0.5, 0.45, want_cls_spec=False)
        want.append(r)
    ins = [dev_inputs(c, dev) for c in cases]
    outs = [None] * R

    def chain():
        for i, c in enumerate(cases):
            y, gt, off = ins[i]
            outs[i] = ops.train_post(y, gt, off, img_hw=(c.height, c.width), lambdas=LAM, anchors=c.anchors, conf_thre=0.5,
                                     iou_thre=0.45, want_resp=True, want_cls_spec=False, out=outs[i], overlapped=i > 0)

    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        chain()  # allocates
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            chain()
        for _ in range(2):
            for o in outs:  # poison
                o["train"]["dy"].fill_(float("nan"))
                o["train"]["loss"].fill_(-1.0)
                o["post"]["keep_cnt"].fill_(-7)
                o["post"]["keep_idx"].fill_(-1)
            g.replay()
            g.replay()
            stream.synchronize()
            for i in range(R):
                assert_same_train(want[i]["train"], outs[i]["train"])
                assert_same_post(want[i]["post"], outs[i]["post"])


def test_fused_step_dense_images_match_the_reference_gradient_golden(cuda_device):
    """Two images with 40+ boxes each (tests/golden/v2_loss_dense.npz: loss and autograd gradient of the REFERENCE):
    the fused step processes them four records per warp -- against the reference itself, not the oracle."""
    case, z = load_golden("v2_loss_dense.npz")
    assert case.m > 8 * case.n  # (the four-per-warp variant is the one that launches)
    r = fused(case, cuda_device, 0.5, 0.45, want_cls_spec=False)["train"]
    assert abs(float(r["loss"]) - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    dy = r["dy"].cpu().numpy()
    assert rel_err(dy, z["dy"]) <= TOL
    assert np.array_equal(dy != 0, z["dy"] != 0), "gradient sparsity pattern differs"
    assert np.allclose(dy, z["dy"], rtol=1e-3, atol=1e-6 * np.abs(z["dy"]).max())
    resp = r["resp"].cpu().numpy()
    for j in np.nonzero(resp != z["resp"])[0]:  # (only exact IoU ties may differ)
        top2 = np.sort(z["iou_all"][j])[-2:]
        assert top2[1] - top2[0] <= 4 * np.spacing(np.float32(top2[1])), int(j)
