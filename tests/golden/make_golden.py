"""Generate the golden vectors under tests/golden/ by executing the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests or fixtures, so parity is pinned by running its functions
(models/utils.py get_iou/nms, models/yolov1.py and models/yolov2.py predict/get_loss) on seeded
synthetic inputs and storing inputs + outputs.  Nothing from the reference is copied; it is
imported from where it lies with two shims: an `albumentations` stub (not
installed here, only used for module-level transform objects) and head-only subclasses whose
`forward` returns an injected head tensor instead of running the conv backbone.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):  # (inspect / importlib probe modules for __file__, __path__, ...)
            raise AttributeError(name)
        return lambda *a, **k: None


for _name in ("albumentations", "albumentations.pytorch"):
    sys.modules.setdefault(_name, _Stub(_name))

import models.utils as ref_utils  # noqa: E402
import models.yolov1 as ref_v1    # noqa: E402
import models.yolov2 as ref_v2    # noqa: E402

from odcp_b200 import synthetic, targets  # noqa: E402


class HeadOnlyV2(ref_v2.YOLOv2):
    def __init__(self, num_cls):
        torch.nn.Module.__init__(self)
        self.anchor_box_size_list = list(synthetic.YOLOV2_ANCHORS)
        self.num_anchor_box = len(self.anchor_box_size_list)
        self.anchor_box_width_list = torch.tensor([b[0] for b in self.anchor_box_size_list])
        self.anchor_box_height_list = torch.tensor([b[1] for b in self.anchor_box_size_list])
        self.cls_list = [str(i) for i in range(num_cls)]
        self.cls2idx = {c: i for i, c in enumerate(self.cls_list)}
        self.num_cls = num_cls

    def forward(self, x):
        return self._y


class HeadOnlyV1(ref_v1.YOLOv1):
    def __init__(self, s_h, s_w, b, num_cls):
        torch.nn.Module.__init__(self)
        self.num_grid_cell_in_height = s_h
        self.num_grid_cell_in_width = s_w
        self.num_anchor_box = b
        self.cls_list = [str(i) for i in range(num_cls)]
        self.cls2idx = {c: i for i, c in enumerate(self.cls_list)}
        self.num_cls = num_cls

    def forward(self, x):
        return self._y


def ref_model(case):
    if case.version == 2:
        return HeadOnlyV2(case.c)
    return HeadOnlyV1(case.s_h, case.s_w, case.a, case.c)


def run_loss(case, lambdas):
    m = ref_model(case)
    y = case.y.clone().requires_grad_(True)
    m._y = y
    x = torch.zeros(case.n, case.height, case.width, 3)
    dense = targets.records_to_dense(case.rec, case.n, case.s_h, case.s_w, case.c, case.version)
    loss = m.get_loss(x, *dense, **lambdas)
    loss.backward()
    # responsible predictor per record, from the reference's own predict + get_iou
    with torch.no_grad():
        bbox = m.predict(x)[2]
        rec = case.rec
        pb = bbox[torch.as_tensor(rec["img"].astype(np.int64)), torch.as_tensor(rec["cy"].astype(np.int64)),
                  torch.as_tensor(rec["cx"].astype(np.int64))]                        # [M,A,4]
        gt = torch.as_tensor(np.stack([rec["x1"], rec["y1"], rec["x2"], rec["y2"]], 1))
        iou = ref_utils.get_iou(pb, gt[:, None, :])
        iou_r, r = torch.max(iou, dim=-1)
    return dict(loss=np.float32(loss.item()), dy=y.grad.numpy().copy(), resp=r.numpy().astype(np.int32),
                iou_resp=iou_r.numpy().copy(), iou_all=iou.numpy().copy())


def run_predict(case):
    m = ref_model(case)
    m._y = case.y
    x = torch.zeros(case.n, case.height, case.width, 3)
    with torch.no_grad():
        return [t.numpy().copy() for t in m.predict(x)]


def run_nms(case, conf_thre, iou_thre):
    """Reference nms per image; the flat predictor index rides along as an extra class column."""
    m = ref_model(case)
    m._y = case.y
    x = torch.zeros(case.n, case.height, case.width, 3)
    with torch.no_grad():
        _, _, bbox, conf, _, spec = m.predict(x)
    out = []
    for i in range(case.n):
        b = bbox[i].reshape(-1, 4)
        c = conf[i].reshape(-1)
        s = spec[i].reshape(-1, case.c)
        idx = torch.arange(len(c), dtype=torch.float32)[:, None]
        kb, kc, ks = ref_utils.nms(b, c, torch.cat([s, idx], -1), conf_thre, iou_thre)
        out.append(dict(idx=ks[:, -1].numpy().astype(np.int32), bbox=kb.numpy(), conf=kc.numpy(),
                        cls_spec=ks[:, :-1].numpy()))
    return out


def case_arrays(case, with_y=True):
    d = dict(version=case.version, n=case.n, s_h=case.s_h, s_w=case.s_w, a=case.a, c=case.c,
             height=case.height, width=case.width, rec=case.rec.view(np.int32).reshape(-1, 12),
             gt_off=case.gt_off)
    if with_y:
        d["y"] = case.y.numpy()
    return d


def pack_nms(res):
    cnt = np.array([len(r["idx"]) for r in res], dtype=np.int32)
    cat = lambda k: np.concatenate([r[k] for r in res], 0)
    return dict(nms_cnt=cnt, nms_idx=cat("idx"), nms_bbox=cat("bbox"), nms_conf=cat("conf"),
                nms_cls_spec=cat("cls_spec"))


def run_collate(version, n, height, width, seed, num_cls=20):
    """The reference's own collate_fn (models/yolov2.py:1380-1555, models/yolov1.py:1160-1355,
    augmentation off; v1's albumentations resize replaced by the identity: the images already have the
    network size) on synthetic boxes; its dense per-box grids are read back at the cell it filled."""
    rng = np.random.default_rng(seed)
    boxes, labels, img = synthetic.make_boxes(rng, n, height, width, 1, 5, num_cls)
    if version == 2:
        m = HeadOnlyV2(num_cls)
    else:
        m = HeadOnlyV1(height // 32, width // 32, 2, num_cls)
        m.resize = lambda image, bboxes, labels: dict(image=image, bboxes=bboxes, labels=labels)
    batch = []
    for i in range(n):
        sel = img == i
        batch.append((i, np.zeros((height, width, 3), np.uint8),
                      {"bbox_list": [list(b) for b in boxes[sel]], "lbl_list": [str(l) for l in labels[sel]]}))
    out = m.collate_fn(batch, augmentation=False)
    _, sig_txty, twth, coord, cls_tgt, obj, x_img_id, bbox_img_id = [np.asarray(t) for t in out]
    mm = len(boxes)
    flat = obj.reshape(mm, -1)
    assert (flat.sum(1) == 1).all()
    cell = flat.argmax(1)
    s_w = obj.shape[2]
    cy, cx = cell // s_w, cell % s_w
    j = np.arange(mm)
    rec = np.zeros(mm, dtype=targets.GT_DTYPE)
    rec["img"] = bbox_img_id
    rec["cy"], rec["cx"] = cy, cx
    rec["cls"] = cls_tgt[j, cy, cx].argmax(-1)
    rec["stx"], rec["sty"] = sig_txty[j, cy, cx, 0], sig_txty[j, cy, cx, 1]
    rec["tw"], rec["th"] = twth[j, cy, cx, 0], twth[j, cy, cx, 1]
    for c, k in enumerate(("x1", "y1", "x2", "y2")):
        rec[k] = coord[j, cy, cx, c]
    return dict(version=version, n=n, height=height, width=width, s_h=obj.shape[1], s_w=s_w, boxes=boxes,
                labels=labels.astype(np.int32), img=img.astype(np.int32),
                rec=rec.view(np.int32).reshape(-1, 12), obj_dtype=str(obj.dtype), sig_dtype=str(sig_txty.dtype))


def run_evaluate(seed, n=12, num_cls=4, height=416, width=416):
    """The reference's own evaluate_model (models/utils.py:171-338) driven by a stand-in model whose
    detect() returns canned detections (jittered copies of the ground truth plus clutter): its per-class
    APs pin the batched evaluation."""
    rng = np.random.default_rng(seed)
    boxes, labels, img = synthetic.make_boxes(rng, n, height, width, 2, 5, num_cls)
    dets = []
    for i in range(n):
        g = boxes[img == i]
        gl = labels[img == i]
        jit = g + rng.normal(0, 12.0, size=g.shape)
        clutter = synthetic.make_boxes(rng, 1, height, width, 3, 3, num_cls)
        db = np.concatenate([jit, clutter[0]], 0).astype(np.float32)
        dl = np.concatenate([np.where(rng.random(len(gl)) < 0.8, gl, rng.integers(0, num_cls, len(gl))), clutter[1]]).astype(np.int32)
        ds = rng.random(len(db)).astype(np.float32)
        order = np.argsort(-ds)
        dets.append((db[order], dl[order], ds[order]))
    cls_list = [str(c) for c in range(num_cls)]

    class FakeModel:
        def __init__(self):
            self.cls_list = cls_list
            self.calls = 0

        def detect(self, image, conf, iou):
            b, l, s_ = dets[self.calls]
            self.calls += 1
            return {"bbox_list": b.tolist(), "lbl_list": [cls_list[i] for i in l], "conf_score_list": s_.tolist(),
                    "cls_spec_conf_score_list": s_.tolist()}

    dataset = [(i, None, {"bbox_list": boxes[img == i].tolist(), "lbl_list": [cls_list[c] for c in labels[img == i]]})
               for i in range(n)]
    res = ref_utils.evaluate_model(FakeModel(), dataset, None)
    max_out = max(len(d[0]) for d in dets)
    det_bbox = np.zeros((n, max_out, 4), np.float32)
    det_label = np.zeros((n, max_out), np.int32)
    det_score = np.zeros((n, max_out), np.float32)
    cnt = np.zeros(n, np.int32)
    for i, (b, l, s_) in enumerate(dets):
        cnt[i] = len(b)
        det_bbox[i, :len(b)], det_label[i, :len(b)], det_score[i, :len(b)] = b, l, s_
    off = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(img, minlength=n), out=off[1:])
    return dict(n=n, num_cls=num_cls, gt_boxes=boxes, gt_labels=labels.astype(np.int32), gt_off=off, det_bbox=det_bbox,
                det_label=det_label, det_score=det_score, keep_cnt=cnt, levels=np.asarray(res["level_list"], np.float64),
                ap=np.stack([res[c] for c in cls_list]))


def run_v1_detect(seed, height=90, width=120, conf_thre=0.5, iou_thre=0.5):
    """The reference's own YOLOv1.detect (models/yolov1.py:439-554) on an injected head tensor.  albumentations is not
    installed offline, so its two Resize transforms are stood in for by what they do (documented behaviour of
    albumentations.Resize with BboxParams(format="pascal_voc")): the image goes through cv2.resize with bilinear
    interpolation, boxes are normalised by the source image size and scaled to the target size in float64."""
    import cv2
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)
    y = torch.randn(1, 7, 7, 30, generator=torch.Generator().manual_seed(seed))
    y[..., 4:10:5] += 0.3
    m = HeadOnlyV1(7, 7, 2, 20)
    m._y = y

    def resize_to(h, w, src_h, src_w):
        def apply(image, bboxes, labels):
            out = []
            for b in bboxes:
                x1, y1, x2, y2 = [float(v) for v in b[:4]]
                out.append((x1 / src_w * w, y1 / src_h * h, x2 / src_w * w, y2 / src_h * h))
            return dict(image=cv2.resize(image, (w, h), interpolation=cv2.INTER_LINEAR), bboxes=out, labels=list(labels))
        return apply

    m.resize = resize_to(224, 224, height, width)
    m.get_resize_transform = lambda h, w: resize_to(h, w, 224, 224)
    d = m.detect(img, conf_thre, iou_thre)
    return dict(img=img, y=y.numpy(), conf_thre=conf_thre, iou_thre=iou_thre,
                bbox=np.asarray(d["bbox_list"], np.float64).reshape(-1, 4), labels=np.asarray([int(l) for l in d["lbl_list"]], np.int32),
                conf=np.asarray(d["conf_score_list"], np.float64), score=np.asarray(d["cls_spec_conf_score_list"], np.float64))


def main():
    lam = synthetic.DEFAULT_LAMBDAS
    if "--only-new" in sys.argv:
        r = run_v1_detect(401)
        np.savez_compressed(os.path.join(HERE, "v1_detect.npz"), **r)
        print("v1_detect", len(r["bbox"]), "boxes")
        return
    if "--only-dense" in sys.argv:
        # --- v2 loss/grad on images DENSE with ground truth (30-40 boxes per image + same-cell collisions): what the
        # fused step's four-records-per-warp form and the train head's record paths see on BASELINE config 5
        torch.set_num_threads(1)
        c = synthetic.with_collisions(synthetic.make_case("g_v2_dense", 2, 2, 13, 13, 5, 20, 416, 416, seed=15, k_lo=30, k_hi=40), 12, seed=3)
        r = run_loss(c, lam)
        np.savez_compressed(os.path.join(HERE, "v2_loss_dense.npz"), **case_arrays(c), **r)
        print("v2_loss_dense", c.m, r["loss"])
        return
    r = run_v1_detect(401)
    np.savez_compressed(os.path.join(HERE, "v1_detect.npz"), **r)
    np.savez_compressed(os.path.join(HERE, "evaluate.npz"), **run_evaluate(301))
    np.savez_compressed(os.path.join(HERE, "v2_collate.npz"), **run_collate(2, 6, 416, 416, 201))
    np.savez_compressed(os.path.join(HERE, "v2_collate_nonsquare.npz"), **run_collate(2, 3, 352, 480, 202))
    np.savez_compressed(os.path.join(HERE, "v1_collate.npz"), **run_collate(1, 5, 224, 224, 203))
    torch.set_num_threads(1)  # single-threaded reductions: the most reproducible reference run

    # --- v2 loss/grad, small, with same-cell collisions
    c = synthetic.with_collisions(synthetic.make_case("g_v2_small", 2, 3, 13, 13, 5, 20, 416, 416, seed=11), 4, seed=1)
    r = run_loss(c, lam)
    np.savez_compressed(os.path.join(HERE, "v2_loss_small.npz"), **case_arrays(c), **r)
    print("v2_loss_small", c.m, r["loss"])

    # --- v2, images dense with ground truth (see --only-dense)
    c = synthetic.with_collisions(synthetic.make_case("g_v2_dense", 2, 2, 13, 13, 5, 20, 416, 416, seed=15, k_lo=30, k_hi=40), 12, seed=3)
    r = run_loss(c, lam)
    np.savez_compressed(os.path.join(HERE, "v2_loss_dense.npz"), **case_arrays(c), **r)
    print("v2_loss_dense", c.m, r["loss"])

    # --- v2 non-square grid (validation path feeds native-size images, SURVEY B-12)
    c = synthetic.make_case("g_v2_nonsq", 2, 2, 10, 13, 5, 20, 320, 416, seed=12, k_hi=4)
    r = run_loss(c, lam)
    np.savez_compressed(os.path.join(HERE, "v2_loss_nonsquare.npz"), **case_arrays(c), **r)
    print("v2_loss_nonsquare", c.m, r["loss"])

    # --- v1 loss/grad (cfg 1 shape, N=4) with collisions
    c = synthetic.with_collisions(synthetic.make_case("g_v1", 1, 4, 7, 7, 2, 20, 448, 448, seed=13), 3, seed=2)
    r = run_loss(c, lam)
    np.savez_compressed(os.path.join(HERE, "v1_loss_small.npz"), **case_arrays(c), **r)
    print("v1_loss_small", c.m, r["loss"])

    # --- non-default lambdas, 3 classes, 3 anchors worth of head (v1 with B=3)
    lam2 = dict(lambda_xy=1.5, lambda_wh=2.5, lambda_conf=0.75, lambda_noobj=0.25, lambda_cls=3.0)
    c = synthetic.make_case("g_v1_b3", 1, 2, 5, 6, 3, 7, 160, 192, seed=14, k_hi=3)
    r = run_loss(c, lam2)
    np.savez_compressed(os.path.join(HERE, "v1_loss_b3c7.npz"), **case_arrays(c), **r,
                        lambdas=np.array(list(lam2.values())))
    print("v1_loss_b3c7", c.m, r["loss"])

    # --- predict outputs
    for name, c in (("v2_predict", synthetic.make_case("g_v2_pred", 2, 2, 13, 13, 5, 20, 416, 416, seed=15)),
                    ("v1_predict", synthetic.make_case("g_v1_pred", 1, 3, 7, 7, 2, 20, 448, 448, seed=16))):
        p = run_predict(c)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **case_arrays(c),
                            **{"out%d" % i: t for i, t in enumerate(p)})
        print(name, [t.shape for t in p])

    # --- NMS: cfg-3 shaped (conf 0.5 / IoU 0.45) and the reference defaults (0.9 / 0.5)
    c = synthetic.cfg3(n=4)
    np.savez_compressed(os.path.join(HERE, "v2_nms_cfg3.npz"), **case_arrays(c),
                        conf_thre=0.5, iou_thre=0.45, **pack_nms(run_nms(c, 0.5, 0.45)))
    c = synthetic.make_case("g_v2_nms_default", 2, 3, 13, 13, 5, 20, 416, 416, seed=17, to_shift=0.634)
    np.savez_compressed(os.path.join(HERE, "v2_nms_default.npz"), **case_arrays(c),
                        conf_thre=0.9, iou_thre=0.5, **pack_nms(run_nms(c, 0.9, 0.5)))
    c = synthetic.make_case("g_v1_nms", 1, 4, 7, 7, 2, 20, 448, 448, seed=18, to_shift=-0.5)
    np.savez_compressed(os.path.join(HERE, "v1_nms.npz"), **case_arrays(c),
                        conf_thre=0.4, iou_thre=0.3, **pack_nms(run_nms(c, 0.4, 0.3)))
    print("nms done")

    # --- get_iou known answers (including identical, disjoint, touching, degenerate boxes)
    rng = np.random.default_rng(19)
    b1 = rng.uniform(0, 400, size=(64, 4)).astype(np.float32)
    b1[:, 2:] = b1[:, :2] + rng.uniform(0, 200, size=(64, 2)).astype(np.float32)
    b2 = rng.uniform(0, 400, size=(64, 4)).astype(np.float32)
    b2[:, 2:] = b2[:, :2] + rng.uniform(0, 200, size=(64, 2)).astype(np.float32)
    b2[0] = b1[0]                          # identical
    b2[1] = b1[1] + 1000                   # disjoint
    b2[2] = [b1[2, 2], b1[2, 1], b1[2, 2] + 5, b1[2, 3]]   # touching edge
    b2[3] = 0                              # all-zero target (the dense grids' empty cells)
    b1[4, 2:] = b1[4, :2]                  # zero-area prediction
    iou = ref_utils.get_iou(torch.as_tensor(b1), torch.as_tensor(b2)).numpy()
    np.savez_compressed(os.path.join(HERE, "iou_kat.npz"), b1=b1, b2=b2, iou=iou)

    # --- cfg 2 at full size (N=64): summary only (y is regenerated from the seed)
    c = synthetic.cfg2()
    r = run_loss(c, lam)
    dyf = r["dy"].reshape(c.n, c.s_h, c.s_w, c.a, 5 + c.c)
    rec = c.rec
    rows = dyf[rec["img"], rec["cy"], rec["cx"], r["resp"]]
    np.savez_compressed(os.path.join(HERE, "v2_cfg2_summary.npz"), **case_arrays(c, with_y=False),
                        y_sum=np.float64(c.y.double().sum().item()), y_abs_sum=np.float64(c.y.double().abs().sum().item()),
                        loss=r["loss"], resp=r["resp"], iou_resp=r["iou_resp"], dy_rows=rows,
                        dy_to_sample=dyf[..., 4].reshape(-1)[::7].copy(),
                        dy_abs_sum=np.float64(np.abs(r["dy"].astype(np.float64)).sum()),
                        dy_nnz=np.int64(np.count_nonzero(r["dy"])))
    print("cfg2", c.m, r["loss"])


if __name__ == "__main__":
    main()
