"""CPU-side checks: the C-ABI library loads and exports every symbol include/yolohead.h declares,
argument validation works without a GPU, the host-side target/shard logic is right, and the
N>1 recombination holds over a real 2-process gloo group.  No kernel runs here."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

from odcp_b200 import _lib, dist as yh_dist, synthetic, targets
from oracle import yolo_head_oracle as O


def header_symbols():
    text = open(os.path.join(ROOT, "include", "yolohead.h")).read()
    return sorted(set(re.findall(r"YH_API\s+[\w\s\*]+?\b(yh_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for name in ("yh_v2_train", "yh_v1_train", "yh_v2_decode", "yh_v1_decode", "yh_compact_targets",
                 "yh_v2_postprocess", "yh_v1_postprocess", "yh_nms", "yh_iou", "yh_abi_version", "yh_last_error"):
        assert name in syms
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes binding mirrors the header one to one


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every exported entry point to the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [name for name in header_symbols() if name not in doc]
    assert not missing, missing


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    raw = C.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert getattr(raw, name) is not None, name
    assert lib.yh_abi_version() == _lib.ABI_VERSION
    assert lib.yh_train_workspace_bytes() >= 64
    assert lib.yh_postprocess_workspace_bytes(4, 845) > 0
    assert lib.yh_compact_workspace_bytes(10, 4) > 0


def test_bad_arguments_are_reported_not_crashed():
    """Validation happens before anything touches the device: negative code + message."""
    lib = _lib.load()
    null = C.c_void_p(0)
    lam = (C.c_float * 5)(5, 5, 1, .5, 1)
    anchors = (C.c_float * 10)(*[1.0] * 10)
    rc = lib.yh_v2_train(null, 1, 13, 13, 5, 20, anchors, 416.0, 416.0, null, null, 0, 0, lam,
                         null, null, null, null, null, null, 0, null)
    assert rc < 0 and lib.yh_last_error()
    rc = lib.yh_v2_train(null, 0, 13, 13, 5, 20, anchors, 416.0, 416.0, null, null, 0, 0, lam,
                         null, null, null, null, null, null, 0, null)
    assert rc == -1
    with pytest.raises(_lib.YoloHeadError):
        _lib.call("yh_nms", null, null, null, 1, 10, 0.5, 0.5, 10, null, null, null, 0, null)


def test_missing_library_is_loud(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "object-detection-collection-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), (dirpath, f)


def test_records_dense_round_trip_and_offsets():
    case = synthetic.cfg2(n=6)
    dense = targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, 2)
    obj = dense[4].numpy()
    assert obj.dtype == np.float64 and obj.sum() == case.m  # one cell per box, fp64 like the reference
    j = np.arange(case.m)
    assert np.array_equal(dense[0].numpy()[j, case.rec["cy"], case.rec["cx"], 0], case.rec["stx"])
    assert np.array_equal(dense[3].numpy()[j, case.rec["cy"], case.rec["cx"]].argmax(-1), case.rec["cls"])
    off = targets.csr_offsets(case.rec, case.n)
    assert np.array_equal(off, case.gt_off) and off[-1] == case.m
    t = targets.records_to_tensor(case.rec)
    assert t.dtype == torch.int32 and tuple(t.shape) == (case.m, 12)
    assert np.array_equal(targets.tensor_to_records(t), case.rec)
    with pytest.raises(ValueError):
        targets.csr_offsets(case.rec[::-1], case.n)


def test_image_shards_cover_the_batch():
    for n, w in ((256, 8), (7, 2), (5, 4), (3, 8)):
        spans = [yh_dist.image_shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = synthetic.with_collisions(synthetic.cfg2(n=12), 6, seed=3)
        lam = synthetic.DEFAULT_LAMBDAS
        rec, off, (lo, hi) = yh_dist.shard_case(case.rec, case.gt_off, case.n, rank, world)
        m_global = yh_dist.global_box_count(len(rec))
        sub = synthetic.HeadCase("shard", 2, hi - lo, case.s_h, case.s_w, case.a, case.c, case.height,
                                 case.width, case.y[lo:hi].contiguous(), rec, off, anchors=case.anchors)
        r = O.train_head_compact(sub, lam, m_global=m_global)  # stands in for the kernel on this rank
        terms, loss = yh_dist.reduce_terms(torch.tensor(r["terms"], dtype=torch.float32),
                                           torch.tensor(r["loss"], dtype=torch.float32))
        dys = [torch.zeros(1) for _ in range(world)] if rank == 0 else None
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, r["dy"]))
        if rank == 0:
            full = O.train_head_compact(case, lam)
            dy = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])], 0)
            out.put(dict(m_global=m_global, m=case.m, loss=float(loss), want_loss=full["loss"],
                         terms=terms.numpy(), want_terms=np.asarray(full["terms"]),
                         dy_err=float(np.abs(dy - full["dy"]).max() / np.abs(full["dy"]).max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_recombine_to_the_unsharded_result():
    """world_size 2 over gloo: shard by image, pass the global box count, all-reduce six floats."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["m_global"] == res["m"]
    assert abs(res["loss"] - res["want_loss"]) <= 1e-5 * abs(res["want_loss"])
    assert np.allclose(res["terms"], res["want_terms"], rtol=1e-5, atol=0)
    assert res["dy_err"] <= 1e-5


@pytest.mark.parametrize("name", ["v2_collate.npz", "v2_collate_nonsquare.npz", "v1_collate.npz"])
def test_host_records_match_the_reference_collate_fn(name):
    """targets.boxes_to_records restates collate_fn's float64 arithmetic: bit-identical to the
    records read back from the reference's own dense grids (tests/golden/make_golden.py)."""
    from conftest import GOLDEN
    z = dict(np.load(os.path.join(GOLDEN, name)))
    want = np.ascontiguousarray(z["rec"]).reshape(-1).view(targets.GT_DTYPE)
    got = targets.boxes_to_records(z["boxes"], z["labels"], z["img"], int(z["height"]), int(z["width"]),
                                   int(z["s_h"]), int(z["s_w"]), int(z["version"]))
    assert got.tobytes() == want.tobytes()
    assert str(z["obj_dtype"]) == "float64"  # SURVEY B-7: the reference's obj_mask is fp64


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
def test_mixin_composes_with_the_unmodified_reference_model():
    """`class YOLOv2(YOLOv2HeadOps, ref.YOLOv2)` takes predict / get_loss /
    detect from the CUDA path and everything else (forward, backbone, collate_fn, train loop) from the
    reference, without new parameters.  Import only: nothing runs without a GPU."""
    import types
    sys.path.insert(0, "/root/reference")

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            return lambda *a, **k: None
    for name in ("albumentations", "albumentations.pytorch"):
        sys.modules.setdefault(name, _Stub(name))
    try:
        import models.yolov2 as ref_v2
        from odcp_b200.models._head import HeadOps
        from odcp_b200.models.yolov2 import YOLOv2HeadOps

        class YOLOv2(YOLOv2HeadOps, ref_v2.YOLOv2):
            pass

        for name in ("predict", "get_loss", "detect"):
            assert getattr(YOLOv2, name) is getattr(HeadOps, name)
        for name in ("forward", "collate_fn", "train_model", "run_one_epoch"):
            assert getattr(YOLOv2, name) is getattr(ref_v2.YOLOv2, name)
        import inspect
        ours = list(inspect.signature(HeadOps.get_loss).parameters)
        theirs = list(inspect.signature(ref_v2.YOLOv2.get_loss).parameters)
        assert len(ours) == len(theirs) == 14 and ours[-5:] == theirs[-5:]  # self + 13 arguments, same lambda names
        # the optimizer: run_one_epoch resolves the module-level name `SGD` (models/yolov2.py:7, :1254)
        import torch.optim
        from odcp_b200.models import patch_reference
        from odcp_b200.optim import SGD as FusedSGD
        assert ref_v2.SGD is torch.optim.SGD
        assert "SGD(" in inspect.getsource(ref_v2.YOLOv2.run_one_epoch)
        patch_reference(ref_yolov2=ref_v2, fused_sgd=True)
        assert ref_v2.SGD is FusedSGD and ref_v2.YOLOv2.get_loss is HeadOps.get_loss
        for kw in ("lr", "momentum", "weight_decay"):  # the keywords the reference passes
            assert kw in inspect.signature(FusedSGD.__init__).parameters
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "config"]:
            del sys.modules[k]


def test_sgd_step_validates_its_tensor_list():
    """Host side of the fused SGD step: argument validation happens before anything touches the device."""
    lib = _lib.load()
    sizes = np.array([5, 16384], dtype=np.int64)
    pp = np.array([0x1000, 0x20000], dtype=np.uint64)
    gp = pp + np.uint64(0x10000000)
    assert lib.yh_sgd_step(None, None, None, None, 0, 0.1, 0.9, 5e-4, 1, None) == 0   # nothing to do
    bad = pp.copy(); bad[0] = 0x1002
    assert lib.yh_sgd_step(bad.ctypes.data, gp.ctypes.data, None, sizes.ctypes.data, 2, 0.1, 0.9, 5e-4, 1, None) < 0
    assert b"aligned" in lib.yh_last_error()
    bad = pp.copy(); bad[1] = 0
    assert lib.yh_sgd_step(bad.ctypes.data, gp.ctypes.data, None, sizes.ctypes.data, 2, 0.1, 0.9, 5e-4, 1, None) < 0
    big = np.array([5, 1 << 31], dtype=np.int64)
    assert lib.yh_sgd_step(pp.ctypes.data, gp.ctypes.data, None, big.ctypes.data, 2, 0.1, 0.9, 5e-4, 1, None) < 0
    assert lib.yh_sgd_step(pp.ctypes.data, None, None, sizes.ctypes.data, 2, 0.1, 0.9, 5e-4, 1, None) < 0


def test_fused_sgd_refuses_cpu_tensors():
    import torch
    from odcp_b200.optim import SGD
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        SGD([p], lr=0.1, momentum=0.9, weight_decay=5e-4).step()


def test_pair_enumeration_of_the_postprocess_covers_the_triangle():
    """Host mirror of the dense pair enumeration in yh_nms.cu (rest_img, phase D): the i < j pairs of K ranked
    candidates are the cells of a K/2 x (K-1) rectangle, addressed through a float reciprocal instead of an
    integer division.  Every pair exactly once, and the float quotient equals the integer one, for every K the
    shared-memory path handles."""
    f32 = np.float32
    for k in range(0, 257):
        ke = k + (k & 1)
        cols, total = ke - 1, (ke >> 1) * (ke - 1)
        if total == 0:
            continue
        inv = f32(1.0) / f32(cols)
        # the device uses an approximate reciprocal: allow it to be off by a few ulp in either direction
        for fudge in (f32(1.0), f32(1.0) + f32(4e-7), f32(1.0) - f32(4e-7)):
            pr = np.arange(total, dtype=np.int64)
            r = ((pr.astype(f32) + f32(0.5)) * (inv * fudge)).astype(np.int64)
            assert np.array_equal(r, pr // cols), (k, float(fudge))
        c = pr - r * cols
        i = np.where(c >= r, r, ke - 1 - r)
        j = np.where(c >= r, c + 1, ke - 1 - c)
        ok = j < k
        assert (i[ok] < j[ok]).all()
        pairs = set(zip(i[ok].tolist(), j[ok].tolist()))
        assert len(pairs) == ok.sum() == k * (k - 1) // 2


def test_threshold_band_of_the_postprocess_is_safe():
    """Host mirror of postprocess_impl's logit band: for thresholds in [0.01, 0.99] the kernel decides
    sigmoid(t) >= thr on the logit when |t - logit(thr)| >= 1e-3.  In float32 the sigmoid at the band's edges must
    be on the right side of the threshold by far more than its own rounding (a few ulp)."""
    f32 = np.float32
    for thr in np.concatenate([np.linspace(0.01, 0.99, 197), [0.5, 0.3, 0.9]]):
        t0 = np.log(thr / (1.0 - thr))
        lo, hi = f32(t0 - 1e-3), f32(t0 + 1e-3)
        sig = lambda t: f32(1.0) / (f32(1.0) + np.exp(-t, dtype=f32))
        ulp = np.spacing(f32(thr))
        assert sig(hi) - f32(thr) > 40 * ulp, thr
        assert f32(thr) - sig(lo) > 40 * ulp, thr


def test_yolov2_net_is_the_reference_network():
    """odcp_b200.train_step.YOLOv2Net (the stock-cuDNN network of BASELINE config 4) against the reference's YOLOv2:
    same parameter count, same state_dict keys, and -- with the reference's weights loaded -- the same head tensor."""
    from oracle import refharness as RH
    if not RH.available():
        pytest.skip("the reference is not staged (oracle/_ref) and /root/reference does not exist")
    import torch
    from odcp_b200 import train_step as T
    ref = RH.load_reference("cpu")
    cls_list = [str(i) for i in range(20)]
    torch.manual_seed(0)
    theirs = ref.yolov2.YOLOv2(cls_list, {c: i for i, c in enumerate(cls_list)}).eval()
    ours = T.YOLOv2Net(cls_list).eval()
    assert sum(p.numel() for p in ours.parameters()) == T.YOLOV2_PARAMETERS == sum(p.numel() for p in theirs.parameters())
    assert list(ours.state_dict().keys()) == list(theirs.state_dict().keys())
    ours.load_state_dict(theirs.state_dict())
    x = torch.rand(1, 416, 416, 3, generator=torch.Generator().manual_seed(3)) * 255.0
    with torch.no_grad():
        assert torch.equal(ours(x), theirs(x))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host cores alone (no GPU): one JSON line with the base contract's keys,
    `impl: reference`, a cpu_baseline describing the run and an e2e object without copies; under a launcher only
    rank 0 prints."""
    import json
    import subprocess
    from oracle import refharness
    if not refharness.available():
        pytest.skip("no reference sources (neither /root/reference nor the staged oracle/_ref)")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2"]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "yolo_head_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("yolov2_head_13x13x5_c20_b256")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert res.returncode == 0 and not res.stdout.strip()  # (the other ranks exit 0 without work)
