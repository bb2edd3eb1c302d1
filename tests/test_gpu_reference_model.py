"""The reference's REAL model behind the drop-in, on the GPU: the unmodified `YOLOv2` (Darknet19 backbone, neck, head;
reference models/yolov2.py:41-431, random init) runs one `run_one_epoch` training batch (models/yolov2.py:1142-1278:
get_loss -> SGD(...) -> zero_grad -> backward -> step) twice from identical weights -- once as it is, once with
`patch_reference(..., fused_sgd=True)` + a channels_last head -- and must end with the same loss and the same
parameters; then `detect` of the patched model against the reference's own predict -> nms -> argmax chain on the same
head tensor.  The reference is imported from the staged copy (oracle/_ref, see oracle/stage_reference.py) or from
/root/reference; without either the test skips (the driver's GPU box has the staged copy: it travels with the snapshot).
"""
import numpy as np
import pytest
import torch

from odcp_b200 import synthetic, targets
from oracle import refharness as RH

pytestmark = pytest.mark.gpu
LAM = synthetic.DEFAULT_LAMBDAS
N = 8


class OneBatch:
    """What run_one_epoch needs from a DataLoader: iteration and `.dataset` with a length."""

    def __init__(self, batch, n):
        self.batch, self.dataset = batch, range(n)

    def __iter__(self):
        yield self.batch


def make_batch(dev):
    case = synthetic.cfg2(n=N)
    gen = torch.Generator().manual_seed(77)
    x = (torch.rand(N, 416, 416, 3, generator=gen) * 255.0).to(dev)
    # the dense per-box grids collate_fn builds (bit-exact: tests/golden/v2_collate*.npz pin records_to_dense against it)
    dense = [t.to(dev) for t in targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, 2)]
    return case, (x, *dense)


def fresh_reference_model(dev):
    RH._loaded.pop("cuda", None)  # a fresh import: the previous run may have patched the classes
    ref = RH.load_reference("cuda")
    cls_list = [str(i) for i in range(20)]
    torch.manual_seed(1234)
    model = ref.yolov2.YOLOv2(cls_list, {c: i for i, c in enumerate(cls_list)}).to(dev)
    return ref, model


def widest_gap_midpoint(values, q=None, lo=None, hi=None):
    """Middle of the widest gap between consecutive sorted `values` inside [lo, hi] (or between two quantiles)."""
    v = np.sort(np.asarray(values, dtype=np.float64))
    if q is not None:
        lo, hi = np.quantile(v, q[0]), np.quantile(v, q[1])
    v = np.concatenate([[lo], v[(v > lo) & (v < hi)], [hi]])
    i = int(np.argmax(np.diff(v)))
    assert v[i + 1] - v[i] > 1e-5, "no usable gap"
    return float(0.5 * (v[i] + v[i + 1]))


def test_reference_yolov2_training_batch_and_detect_through_the_drop_in(cuda_device, monkeypatch):
    if not RH.available():
        pytest.skip("the reference is not staged (oracle/_ref) and /root/reference does not exist")
    dev = cuda_device
    case, batch = make_batch(dev)
    # fp32 convolutions in both runs: the comparison is between memory layouts / kernels, not TF32 roundings
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)

    # ---- the reference as it is
    ref, model = fresh_reference_model(dev)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params > 60e6  # the real thing: backbone + neck + head, ~67 M parameters
    loss_ref = float(model.run_one_epoch(2, OneBatch(batch, N), lr=1e-3, train=True, **LAM))
    after_ref = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # ---- the same step through the drop-in: CUDA head path, fused SGD, channels_last head
    from odcp_b200.models import patch_reference
    from odcp_b200.models.layout import is_free_view
    ref2, model2 = fresh_reference_model(dev)
    orig_predict = ref2.yolov2.YOLOv2.predict
    for k, v in model2.state_dict().items():
        assert torch.equal(v, init[k]), k  # identical start
    done = patch_reference(ref2.yolov1, ref2.yolov2, ref2.utils, fused_sgd=True, channels_last=True)
    assert "YOLOv2" in done
    torch.manual_seed(1234)  # constructed AFTER the patch: its convolutions are channels_last from the start
    model2 = ref2.yolov2.YOLOv2(model2.cls_list, model2.cls2idx).to(dev)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, init[k]), k  # same keys, same values
    # the reference's own head() (permute + reshape, models/yolov2.py:338-362) is now a view of the conv's output
    seen = {}
    hook = model2.head_model[-1].register_forward_hook(lambda m, i, o: seen.update(out=o))
    model2.eval()  # (no BatchNorm statistics are touched by this probe)
    with torch.no_grad():
        y_view = model2(batch[0][:2])
    model2.train()
    hook.remove()
    assert is_free_view(seen["out"], y_view)
    loss_new = float(model2.run_one_epoch(2, OneBatch(batch, N), lr=1e-3, train=True, **LAM))
    after_new = {k: v.detach().clone() for k, v in model2.state_dict().items()}

    assert abs(loss_new - loss_ref) <= 1e-5 * abs(loss_ref), (loss_new, loss_ref)
    worst = 0.0
    for k in after_ref:
        a, b, i = after_ref[k].double(), after_new[k].double(), init[k].double()
        if not after_ref[k].is_floating_point():
            assert torch.equal(after_ref[k], after_new[k]), k
            continue
        upd = (a - i).abs().max().item()
        err = (a - b).abs().max().item()
        # relative to the size of the update itself (cuDNN's backward is not bit-reproducible between layouts)
        ulp = 1e-7 * max(a.abs().max().item(), 1e-30)  # (a parameter that only sees weight decay moves by a few ulps)
        assert err <= 2e-3 * upd + ulp, (k, err, upd)
        worst = max(worst, max(err - ulp, 0.0) / max(upd, 1e-30))
    assert worst < 2e-3

    # ---- detect: the patched model against the reference's own chain on the SAME head tensor
    model2.eval()
    img = (torch.rand(416, 416, 3, generator=torch.Generator().manual_seed(5)) * 255.0).numpy().astype(np.float32)
    with torch.no_grad():
        x = torch.tensor(np.asarray([img])).to(dev)
        _, _, bbox, conf, _, spec = orig_predict(model2, x)  # the reference's predict (its torch ops, on the GPU)
        RH._loaded.pop("cpu", None)  # a fresh, unpatched import of the reference's utils: ITS nms (CPU only)
        ref_utils = RH.load_reference("cpu").utils
    # Thresholds the comparison is well-posed for: torch's CUDA sigmoid/exp and the kernel's differ in the last ulp, and
    # a random-init head puts every objectness next to 0.5, so a threshold ON a confidence (or on a pair's IoU) would
    # test rounding, not the path (north_star: ties within 1 ulp of a threshold are excluded).  Both thresholds are put
    # in the middle of the widest gap of the values they are compared with.
    conf_thre = widest_gap_midpoint(conf.flatten().cpu().numpy().astype(np.float64), q=(0.85, 0.95))
    cand = bbox.reshape(-1, 4)[conf.flatten() >= conf_thre].cpu().numpy().astype(np.float64)
    assert 20 <= cand.shape[0] <= 200, cand.shape
    pair_iou = ref_utils.get_iou(cand[:, None, :], cand[None, :, :], numpy=True)[np.triu_indices(cand.shape[0], 1)]
    iou_thre = widest_gap_midpoint(pair_iou, lo=0.35, hi=0.65)
    got = model2.detect(img, conf_thre, iou_thre)
    kb, kc, ks = ref_utils.nms(bbox.cpu(), conf.cpu(), spec.cpu(), conf_thre, iou_thre)
    assert len(got["bbox_list"]) == kb.shape[0] > 0
    assert np.allclose(np.asarray(got["bbox_list"]), kb.numpy(), rtol=1e-4, atol=1e-3)
    assert np.allclose(np.asarray(got["conf_score_list"]), kc.numpy(), rtol=1e-5, atol=1e-6)
    assert [model2.cls_list.index(l) for l in got["lbl_list"]] == ks.argmax(-1).tolist()
    assert np.allclose(np.asarray(got["cls_spec_conf_score_list"]), ks.max(-1).values.numpy(), rtol=1e-4, atol=1e-6)
