"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded inputs.

    python tests/golden/make_golden.py            # needs /root/reference; run in the build container only

The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path (tests/test_gpu_*.py).
Shims (no edits to the reference):
  * torch.Tensor.cuda / torch.cuda availability: the reference calls .cuda() unconditionally (losses.py:310,359-361;
    utils.py:96-98,113); on this CPU-only box .cuda() becomes a no-op (clone for leaf tensors that require grad, so the
    in-place writes at losses.py:311-328 stay legal);
  * matplotlib stub modules: util_track/kf.py:10 imports it and MC3D_crop_tracker imports kf.
"""
import csv
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

REF = "/root/reference"


def _shim():
    torch.Tensor.cuda = lambda self, *a, **k: self.clone() if (self.requires_grad and self.is_leaf) else self
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _import_retinanet(three_d):
    for k in [k for k in sys.modules if k == "retinanet" or k.startswith("retinanet.")]:
        del sys.modules[k]
    path = os.path.join(REF, "pytorch_retinanet_detector_directional") if three_d else REF
    sys.path.insert(0, path)
    try:
        from retinanet import anchors, losses, model, utils
    finally:
        sys.path.pop(0)
    return losses, utils, model, anchors


def _np(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (_np(v) if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------------------------------------------------ losses
def golden_losses():
    for three_d in (True, False):
        losses, _, _, _ = _import_retinanet(three_d)
        H, W = (64, 64)
        g = synth.gen(11 if three_d else 12)
        anc = synth.anchors(H, W)
        A = anc.shape[1]
        maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
        ann = maker(4, 6, H, W, g, n_pad=2, empty_images=(1,), **synth.TINY)
        ann[3, 2:6] = -1.0  # image 3: padding rows in the middle of the valid ones would be unusual; keep them trailing
        cls, reg = synth.head_outputs(4, A, 8, 12 if three_d else 4, g)
        cls[0, :5, :] = torch.tensor([0.0, 1.0, 5e-5, 0.99995, 0.5, 1e-4, 0.9999, 0.3])  # clamp edges
        cls.requires_grad_(True)
        reg.requires_grad_(True)
        out = losses.FocalLoss()(cls, reg, anc, ann)
        total = sum(o.sum() * wgt for o, wgt in zip(out, (1.0, 0.7, 1.3)))
        total.backward()
        # per-image assignment from the reference's own calc_iou + torch.max
        iou_max, iou_arg = [], []
        for j in range(4):
            rows = ann[j][ann[j][:, 20 if three_d else 4] != -1]
            if rows.shape[0] == 0:
                iou_max.append(torch.zeros(A)); iou_arg.append(torch.zeros(A, dtype=torch.int64)); continue
            if three_d:
                xs, ys = rows[:, 0:16:2], rows[:, 1:16:2]
                b2d = torch.stack((xs.min(1)[0], ys.min(1)[0], xs.max(1)[0], ys.max(1)[0]), 1)
            else:
                b2d = rows[:, :4]
            iou = losses.calc_iou(anc[0], b2d)
            m, a = torch.max(iou, dim=1)
            iou_max.append(m); iou_arg.append(a)
        rows0 = ann[0][ann[0][:, -1 if not three_d else 20] != -1]
        save("loss3d" if three_d else "loss2d", anchors=anc, annotations=ann, classification=cls, regression=reg,
             losses=torch.cat([o.detach() for o in out]), grad_weights=np.array([1.0, 0.7, 1.3]),
             dcls=cls.grad, dreg=reg.grad, iou_max=torch.stack(iou_max), iou_argmax=torch.stack(iou_arg),
             iou_matrix0=losses.calc_iou(anc[0], rows0[:, 16:20] if three_d else rows0[:, :4]))


# ------------------------------------------------------------------------------------------------------------ decode
def golden_decode():
    _, utils3, _, _ = _import_retinanet(True)
    g = synth.gen(21)
    anc = synth.anchors(64, 80)
    A = anc.shape[1]
    reg = torch.randn(3, A, 12, generator=g) * 0.3
    out3 = utils3.BBoxTransform()(anc, reg)
    _, utils2, _, _ = _import_retinanet(False)
    deltas = torch.randn(3, A, 4, generator=g) * 0.5
    out2 = utils2.BBoxTransform()(anc, deltas)
    clipped = utils2.ClipBoxes()(out2.clone(), torch.zeros(3, 3, 64, 80))
    save("decode", anchors=anc, regression=reg, decoded3d=out3, deltas=deltas, decoded2d=out2, clipped2d=clipped,
         image_hw=np.array([64, 80]))


# --------------------------------------------------------------------------------------------------------------- nms
def golden_nms():
    from torchvision.ops import nms
    _, _, model3, _ = _import_retinanet(True)
    g = synth.gen(31)
    boxes, scores = synth.clustered_boxes(600, g)
    scores[100:140] = scores[100]           # ties: stable order decides
    boxes[200:210] = boxes[200]             # identical boxes
    boxes[300] = torch.tensor([50.0, 50.0, 40.0, 40.0])   # degenerate (x2 < x1)
    out = {"boxes": boxes, "scores": scores}
    for thr in (0.5, 0.3, 0.1, 0.8, 0.2):
        out[f"keep_{thr}"] = nms(boxes, scores, thr)
    idxs = torch.randint(0, 4, (600,), generator=g)
    out["idxs"] = idxs
    out["keep_batched_0.5"] = model3.batched_nms(boxes, scores, idxs, 0.5)
    # threshold-equality case: IoU exactly 0.5 must be KEPT (suppress iff IoU > thr)
    eq_boxes = torch.tensor([[0.0, 0.0, 2.0, 2.0], [0.0, 0.0, 2.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    eq_scores = torch.tensor([0.9, 0.8, 0.7])
    out["eq_boxes"], out["eq_scores"], out["eq_keep_0.5"] = eq_boxes, eq_scores, nms(eq_boxes, eq_scores, 0.5)
    save("nms", **out)


# ---------------------------------------------------------------------------------------- model post-processing glue
class _Feeder(torch.nn.Module):
    """stands in for regressionModel / classificationModel: hands back pre-made per-level slices"""

    def __init__(self, full, level_sizes):
        super().__init__()
        self.parts = list(torch.split(full, level_sizes, dim=1))
        self.i = 0

    def forward(self, feature):
        part = self.parts[self.i % len(self.parts)]
        self.i += 1
        return part

    def reset(self):
        self.i = 0


def _level_sizes(H, W):
    return [((H + 2 ** l - 1) // 2 ** l) * ((W + 2 ** l - 1) // 2 ** l) * 9 for l in (3, 4, 5, 6, 7)]


def golden_postprocess():
    H, W = 96, 128
    sizes = _level_sizes(H, W)
    A = sum(sizes)
    # ---- 3D model: default branch (B = 1) and MULTI_FRAME (B = 3)
    _, _, model3, _ = _import_retinanet(True)
    g = synth.gen(41)
    net = model3.resnet18(num_classes=8, pretrained=False)
    net.eval()
    cls = synth.detection_scores(3, A, 8, g, objects=12, per_object=9, lo=0.05, hi=1.0, background=0.04)
    cls[:, 50:80, 2] = 0.25          # a run of tied scores
    reg = torch.randn(3, A, 12, generator=g) * 0.1
    reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]) + torch.randn(3, A, 4, generator=g) * 0.05
    img = torch.zeros(3, 3, H, W)
    with torch.no_grad():
        net.regressionModel, net.classificationModel = _Feeder(reg[:1], sizes), _Feeder(cls[:1], sizes)
        s1, c1, b1 = net(img[:1])
        net.regressionModel, net.classificationModel = _Feeder(reg, sizes), _Feeder(cls, sizes)
        sm, cm, bm, im = net(img, MULTI_FRAME=True)
        net.regressionModel, net.classificationModel = _Feeder(reg, sizes), _Feeder(cls, sizes)
        bl, cl = net(img, LOCALIZE=True)
    # a denser variant so that the ladder has to climb above its first rung (more than 10000 candidates per class);
    # inputs are regenerated from the seed by the tests (synth.dense_detection_inputs), outputs are stored as digests
    Hd, Wd = 256, 320
    sizes_d = _level_sizes(Hd, Wd)
    cls_d, reg_d = synth.dense_detection_inputs(43, Hd, Wd)
    with torch.no_grad():
        net.regressionModel, net.classificationModel = _Feeder(reg_d, sizes_d), _Feeder(cls_d, sizes_d)
        sd, cd, bd = net(torch.zeros(1, 3, Hd, Wd))
    save("post3d", image_hw=np.array([H, W]), classification=cls, regression=reg, scores=s1, classes=c1, boxes=b1,
         mf_scores=sm, mf_classes=cm, mf_boxes=bm, mf_im=im, loc_boxes_digest=synth.digest(bl),
         dense_hw=np.array([Hd, Wd]), dense_seed=np.array([43]), dense_count=np.bincount(_np(cd), minlength=8),
         dense_scores_digest=synth.digest(sd), dense_boxes_digest=synth.digest(bd), dense_head_scores=sd[:64],
         dense_head_boxes=bd[:64])
    # ---- 2D model
    _, _, model2, _ = _import_retinanet(False)
    g = synth.gen(42)
    net2 = model2.resnet18(num_classes=8, pretrained=False)
    net2.eval()
    cls2 = synth.detection_scores(1, A, 8, g, objects=15, per_object=9, lo=0.05, hi=1.0, background=0.04)
    reg2 = torch.randn(1, A, 4, generator=g) * 0.5
    with torch.no_grad():
        net2.regressionModel, net2.classificationModel = _Feeder(reg2, sizes), _Feeder(cls2, sizes)
        s2, c2, b2 = net2(torch.zeros(1, 3, H, W))
    save("post2d", image_hw=np.array([H, W]), classification=cls2, regression=reg2, scores=s2, classes=c2, boxes=b2)


# -------------------------------------------------------------------------------------------------------- homography
def _dlt(space, image):
    """3x4 projection from (x,y,z) -> (u,v) pairs by the direct linear transform (float64)."""
    n = space.shape[0]
    Xh = np.concatenate([space, np.ones((n, 1))], axis=1)
    rows = []
    for i in range(n):
        u, v = image[i]
        rows.append(np.concatenate([Xh[i], np.zeros(4), -u * Xh[i]]))
        rows.append(np.concatenate([np.zeros(4), Xh[i], -v * Xh[i]]))
    _, _, vt = np.linalg.svd(np.asarray(rows))
    P = vt[-1].reshape(3, 4)
    return P / P[2, 3]


def golden_homography():
    sys.path.insert(0, REF)
    import homography as ref_h
    sys.path.pop(0)
    # ---- the reference's own CSV rows: state, state_to_space (float32, cols 27-34), state_to_im (float64, cols 11-26)
    rows = []
    with open(os.path.join(REF, "3D_tracking_results.csv")) as fh:
        rd = csv.reader(fh)
        header = next(rd)
        for r in rd:
            if len(r) > 44 and r[36] == "p1c1":
                rows.append(r)
    rows = rows[:: max(1, len(rows) // 400)][:400]
    st = np.array([[float(r[39]), float(r[40]), float(r[43]), float(r[42]), float(r[44]), float(r[35])] for r in rows],
                  dtype=np.float32)
    space_csv = np.array([[np.float32(x) for x in r[27:35]] for r in rows], dtype=np.float32)
    im_csv = np.array([[float(x) for x in r[11:27]] for r in rows], dtype=np.float64).reshape(-1, 8, 2)
    hg = ref_h.Homography()
    space = hg.state_to_space(torch.from_numpy(st))
    assert np.array_equal(_np(space)[:, :4, :2].reshape(-1, 8), space_csv), "reference state_to_space != CSV"
    y0 = _np(space)[:, 0, 1]
    lo, hi = y0 <= 60, y0 > 60
    sp64 = _np(space).astype(np.float64)
    P_lo = _dlt(sp64[lo].reshape(-1, 3), im_csv[lo].reshape(-1, 2))
    P_hi = _dlt(sp64[hi].reshape(-1, 3), im_csv[hi].reshape(-1, 2))
    # ---- reference objects with synthetic correspondences: 3 cameras, two homographies (wrapper)
    Pm, Hm = synth.camera_matrices(3)
    names = synth.CAMERAS[:3]
    hg1, hg2 = ref_h.Homography(), ref_h.Homography()
    for i, n in enumerate(names):
        hg1.correspondence[n] = {"P": Pm[i, 0], "H": Hm[i, 0], "H_inv": np.linalg.inv(Hm[i, 0])}
        hg2.correspondence[n] = {"P": Pm[i, 1], "H": Hm[i, 1], "H_inv": np.linalg.inv(Hm[i, 1])}
    hg1.default_correspondence = hg2.default_correspondence = names[0]
    wr = ref_h.Homography_Wrapper(hg1, hg2)
    g = synth.gen(51)
    states, cam = synth.vehicle_states(300, g, n_cams=3)
    states = torch.cat((states, torch.rand(300, 1, generator=g)), dim=1)      # a 7th (velocity) column is ignored
    cam_names = [names[i] for i in cam.tolist()]
    out = dict(csv_states=st, csv_space=space_csv, csv_im=im_csv, csv_P_lo=P_lo, csv_P_hi=P_hi, P=Pm, H=Hm,
               states=states, cam=cam)
    out["space"] = hg1.state_to_space(states)
    out["im_single"] = hg1.state_to_im(states, name=names[1])
    out["im_list"] = hg1.state_to_im(states, name=cam_names)
    out["im_wrapper_list"] = wr.state_to_im(states, name=cam_names)
    out["im_wrapper_single"] = wr.state_to_im(states, name=names[2])
    det = out["im_wrapper_list"].float() + torch.randn(300, 8, 2, generator=g) * 0.5      # detector-like float32 corners
    heights = hg1.guess_heights(["sedan", "semi", "van", "nothing"] * 75)
    out["det"], out["heights"] = det, heights
    out["space_from_im_list"] = hg1.im_to_space(det, name=cam_names, heights=heights)
    out["space_from_im_wrapper"] = wr.im_to_space(det, name=cam_names, heights=heights)
    out["state_single"] = hg1.im_to_state(det, name=names[1], heights=heights)
    out["state_list"] = hg1.im_to_state(det, name=cam_names, heights=heights)
    out["state_wrapper_list"] = wr.im_to_state(det, name=cam_names, heights=heights)
    out["state_from_space"] = hg1.space_to_state(out["space_from_im_list"])
    out["state_from_space_f32"] = hg1.space_to_state(out["space"])
    # the trackers' two-pass refinement (MC3D_crop_tracker.py:364-370)
    boxes = wr.im_to_state(det, heights=heights, name=cam_names)
    repro = wr.state_to_im(boxes, name=cam_names)
    refined = wr.height_from_template(repro, heights, det)
    out["hft_f64_f32_f32"] = refined
    out["state_refined"] = wr.im_to_state(det, heights=refined, name=cam_names)
    out["hft_all_f64"] = hg1.height_from_template(repro, heights.double(), det.double())
    out["hft_all_f32"] = hg1.height_from_template(repro.float(), heights, det)
    out["space_to_im_f32pts"] = wr.space_to_im(out["space"], name=cam_names)
    save("homography", **out)


# ----------------------------------------------------------------------------------------------------------- tracker
def golden_tracker():
    sys.path.insert(0, REF)
    import homography as ref_h
    import MC3D_crop_tracker as mc
    sys.path.pop(0)
    T = mc.MC_Crop_Tracker
    hg = ref_h.Homography()
    me = types.SimpleNamespace(hg=hg, phi_match=0.1)
    me.md_iou = lambda a, b: T.md_iou(me, a, b)
    g = synth.gen(61)
    states, _ = synth.vehicle_states(120, g, n_cams=1)
    states[:, 0] = 100 + torch.rand(120, generator=g) * 300        # dense enough to overlap
    second = states.clone()
    second[:, :2] += torch.randn(120, 2, generator=g) * torch.tensor([4.0, 1.0])
    second = second[torch.randperm(120, generator=g)][:100]
    scores = torch.rand(120, generator=g)

    def fp(s):
        sp = hg.state_to_space(s.clone())
        b = torch.zeros([sp.shape[0], 4])
        b[:, 0] = torch.min(sp[:, 0:4, 0], dim=1)[0]; b[:, 2] = torch.max(sp[:, 0:4, 0], dim=1)[0]
        b[:, 1] = torch.min(sp[:, 0:4, 1], dim=1)[0]; b[:, 3] = torch.max(sp[:, 0:4, 1], dim=1)[0]
        return b
    fa, fb = fp(states), fp(second)
    f, s = fa.shape[0], fb.shape[0]
    A = fa.unsqueeze(1).repeat(1, s, 1).double()
    Bm = fb.unsqueeze(0).repeat(f, 1, 1).double()
    iou = T.md_iou(me, A, Bm)
    degenerate = torch.tensor([[[1.0, 1.0, 1.0, 1.0]]], dtype=torch.float64)
    Pm, _ = synth.camera_matrices(1)
    hg.correspondence["p1c1"] = {"P": Pm[0, 0], "H": np.eye(3), "H_inv": np.eye(3)}
    corners = hg.state_to_im(states, name="p1c1").float()
    save("tracker", states=states, second=second, scores=scores, footprint=fa, md_iou=iou, cost=1.0 - iou,
         md_iou_degenerate=T.md_iou(me, degenerate, degenerate),
         space_nms_0_1=T.space_nms(me, states, scores, threshold=0.1),
         space_nms_0_4=T.space_nms(me, states, scores, threshold=0.4),
         corners=corners, im_nms_0_3=T.im_nms(me, corners, scores, threshold=0.3),
         im_nms_groups=T.im_nms(me, corners, scores, threshold=0.3, groups=torch.zeros(120)))


def best_box_inputs(seed=91, n=37, d=9):
    """seeded select_best_box scenario shared with the tests: n tracked objects (a priori states), d candidate detections
    per object (jittered copies, some far away), confidences and classes"""
    g = synth.gen(seed)
    prior, _ = synth.vehicle_states(n, g, n_cams=1)
    preds = prior.unsqueeze(1).repeat(1, d, 1)
    preds[:, :, :2] += torch.randn(n, d, 2, generator=g) * torch.tensor([6.0, 1.5])
    preds[:, :, 2:5] *= 1.0 + 0.1 * torch.randn(n, d, 3, generator=g)
    preds[:, d - 1, 0] += 500.0                                   # one detection per object without any overlap
    confs = torch.rand(n, d, generator=g)
    classes = torch.randint(0, 8, (n, d), generator=g)
    return prior, preds, confs, classes


def golden_best_box():
    """MC_Crop_Tracker.select_best_box (MC3D_crop_tracker.py:974-1028) and MOT_Evaluator.iou (mot_evaluator.py:87-118, the
    scalar IoU with 1e-6 added to the union) called unbound on seeded inputs."""
    sys.path.insert(0, REF)
    import homography as ref_h
    import MC3D_crop_tracker as mc
    import mot_evaluator as ev
    sys.path.pop(0)
    T = mc.MC_Crop_Tracker
    prior, preds, confs, classes = best_box_inputs()
    n, d = confs.shape
    out = {}
    for W in (0.4, 0.0, 1.0):
        me = types.SimpleNamespace(hg=ref_h.Homography(), W=W)
        me.md_iou = lambda a, b: T.md_iou(me, a, b)
        best, cls, cf = T.select_best_box(me, prior.clone(), preds.clone().reshape(-1, 6), confs.clone(), classes.clone(), n)
        tag = str(W).replace(".", "_")
        out[f"best_{tag}"], out[f"cls_{tag}"], out[f"conf_{tag}"] = best, cls, cf
    # the evaluator's scalar IoU on every (i, j) of two small box sets (float64 python arithmetic on tensor elements)
    g = synth.gen(92)
    a = torch.rand(23, 4, generator=g, dtype=torch.float64) * 50
    a[:, 2:] += a[:, :2] + 1.0
    b = a[torch.randperm(23, generator=g)][:17] + torch.randn(17, 4, generator=g, dtype=torch.float64) * 3
    E = ev.MOT_Evaluator
    eps_iou = torch.zeros(23, 17, dtype=torch.float64)
    for i in range(23):
        for j in range(17):
            eps_iou[i, j] = E.iou(None, a[i], b[j])
    save("best_box", prior=prior, preds=preds, confs=confs, classes=classes, eval_a=a, eval_b=b, eval_iou=eps_iou, **out)


def parse_inputs(seed=95, n_obj=60, n_cams=3):
    """seeded parse_detections scenario shared with the tests: detector output [d,20] (8 projected corners + 2D box) for
    objects seen by their camera, jittered duplicates (so both NMS stages bite), scores around the cut, integer labels"""
    g = synth.gen(seed)
    Pm, Hm = synth.camera_matrices(n_cams)
    st, cam = synth.vehicle_states(n_obj, g, n_cams=n_cams)
    dup = torch.randint(0, n_obj, (n_obj,), generator=g)
    st = torch.cat((st, st[dup] + torch.randn(n_obj, 6, generator=g) * torch.tensor([0.6, 0.15, 0.3, 0.1, 0.1, 0.0])))
    cam = torch.cat((cam, cam[dup]))
    d = st.shape[0]
    scores = torch.rand(d, generator=g)
    labels = torch.randint(0, 8, (d,), generator=g)
    return st, cam.long(), scores, labels, Pm, Hm


def golden_parse():
    """MC_Crop_Tracker.parse_detections (MC3D_crop_tracker.py:319-383) and remove_overlaps' NMS (:482-518) from the
    unmodified reference, called unbound on a namespace that carries the fields they read."""
    sys.path.insert(0, REF)
    import homography as ref_h
    import MC3D_crop_tracker as mc
    sys.path.pop(0)
    from torchvision.ops import nms as tv_nms
    T = mc.MC_Crop_Tracker
    st, cam, scores, labels, Pm, Hm = parse_inputs()
    names = synth.CAMERAS[:Pm.shape[0]]
    hg1, hg2 = ref_h.Homography(), ref_h.Homography()
    for i, n in enumerate(names):
        hg1.correspondence[n] = {"P": Pm[i, 0], "H": Hm[i, 0], "H_inv": np.linalg.inv(Hm[i, 0])}
        hg2.correspondence[n] = {"P": Pm[i, 1], "H": Hm[i, 1], "H_inv": np.linalg.inv(Hm[i, 1])}
    hg1.default_correspondence = hg2.default_correspondence = names[0]
    wr = ref_h.Homography_Wrapper(hg1, hg2)
    corners = wr.state_to_im(st, name=[names[i] for i in cam.tolist()]).float()             # [d,8,2]
    box2d = torch.stack((corners[:, :, 0].min(1).values, corners[:, :, 1].min(1).values,
                         corners[:, :, 0].max(1).values, corners[:, :, 1].max(1).values), dim=1)
    boxes = torch.cat((corners.reshape(-1, 16), box2d), dim=1)                               # the detector's [d,20] rows
    out = dict(boxes=boxes, cam=cam, scores=scores, labels=labels, P=Pm, H=Hm)
    for tag, kw in (("nms", dict(perform_nms=True, refine_height=False)), ("nms_refined", dict(perform_nms=True, refine_height=True)),
                    ("plain", dict(perform_nms=False, refine_height=False))):
        me = types.SimpleNamespace(sigma_d=0.35, phi_nms_im=0.3, phi_nms_space=0.1, est_ts=False, cameras=names, hg=wr)
        me.im_nms = lambda *a, **k: T.im_nms(me, *a, **k)
        me.space_nms = lambda *a, **k: T.space_nms(me, *a, **k)
        b, l, s, c = T.parse_detections(me, scores.clone(), labels.clone(), boxes.clone(), cam.clone(), **kw)
        out[f"states_{tag}"], out[f"labels_{tag}"], out[f"scores_{tag}"], out[f"cams_{tag}"] = b, l, s, c
    # remove_overlaps: the state -> footprint -> nms(frames alive) core (:495-508) on the first half of the states
    view = st[:40]
    sp = hg1.state_to_space(view.clone())
    fp = torch.zeros([sp.shape[0], 4])
    fp[:, 0] = torch.min(sp[:, 0:4, 0], dim=1)[0]; fp[:, 2] = torch.max(sp[:, 0:4, 0], dim=1)[0]
    fp[:, 1] = torch.min(sp[:, 0:4, 1], dim=1)[0]; fp[:, 3] = torch.max(sp[:, 0:4, 1], dim=1)[0]
    alive = torch.randint(1, 300, (40,), generator=synth.gen(96))
    out["overlap_states"], out["overlap_alive"] = view, alive
    out["overlap_keep"] = tv_nms(fp.float(), alive.float(), 0.2)
    save("parse", **out)


def ts_bias_inputs(seed=81, n_obj=45, n_cams=4):
    """seeded estimate_ts_bias scenario shared with the tests: objects seen by 1-3 cameras (jittered duplicates), a
    filter view with both directions, per-camera timestamps"""
    g = synth.gen(seed)
    base, _ = synth.vehicle_states(n_obj, g, n_cams=1)
    base[:, 0] = 200 + torch.rand(n_obj, generator=g) * 1500
    boxes, cams = [], []
    for k in range(n_obj):
        seen = torch.randperm(n_cams, generator=g)[: int(torch.randint(1, 4, (1,), generator=g))]
        for c in seen.tolist():
            b = base[k].clone()
            b[:2] += torch.randn(2, generator=g) * torch.tensor([2.0, 0.3])
            boxes.append(b)
            cams.append(c)
    order = torch.randperm(len(boxes), generator=g)
    boxes = torch.stack(boxes)[order].contiguous()
    cams = torch.tensor(cams)[order].contiguous()
    n_filter = 30
    objs = torch.zeros(n_filter, 7)
    objs[:, :5] = synth.vehicle_states(n_filter, g, n_cams=1)[0][:, :5]
    objs[:, 5] = torch.where(torch.rand(n_filter, generator=g) < 0.5, -1.0, 1.0)
    objs[:, 6] = 90 + torch.rand(n_filter, generator=g) * 40
    timestamps = [10.0, 10.013, 9.991, 10.02]
    return boxes, cams, objs, timestamps


def golden_ts_bias():
    """MC_Crop_Tracker.estimate_ts_bias (MC3D_crop_tracker.py:237-315) called unbound, twice (the bias accumulates), plus
    the one-direction case (:262-265 falls back to mu_v)"""
    sys.path.insert(0, REF)
    import homography as ref_h
    import MC3D_crop_tracker as mc
    sys.path.pop(0)
    T = mc.MC_Crop_Tracker
    boxes, cams, objs, timestamps = ts_bias_inputs()
    out = {}
    for tag, view in (("both", objs), ("eastbound_only", objs[objs[:, 5] == 1])):
        me = types.SimpleNamespace(hg=ref_h.Homography(), phi_nms_space=0.1, ts_alpha=0.05, timestamps=list(timestamps),
                                   ts_bias=[0 for _ in timestamps])
        me.md_iou = lambda a, b: T.md_iou(me, a, b)
        me.filter = types.SimpleNamespace(view=lambda with_direction=False, v=view: (list(range(len(v))), v.clone()), mu_v=105.0)
        T.estimate_ts_bias(me, boxes.clone(), cams)
        out[f"bias1_{tag}"] = np.array(me.ts_bias, dtype=np.float64)
        me.timestamps = [t + 1 / 30.0 + 0.001 * k for k, t in enumerate(timestamps)]
        T.estimate_ts_bias(me, boxes.clone(), cams)
        out[f"bias2_{tag}"] = np.array(me.ts_bias, dtype=np.float64)
    save("ts_bias", boxes=boxes, cams=cams, objs=objs, timestamps=np.array(timestamps), **out)


def kf_inputs(seed=71, n=40, m=25):
    """seeded Kalman-filter scenario shared with the tests (the model matrices mimic a fitted kf_params INIT dict)"""
    g = synth.gen(seed)
    S, M = 6, 5
    F = torch.eye(S)
    H = torch.zeros(M, S); H[:M, :M] = torch.eye(M)
    A = torch.randn(S, S, generator=g) * 0.3
    Q = A @ A.t() + torch.eye(S) * 0.5
    Bm = torch.randn(M, M, generator=g) * 0.2
    R = Bm @ Bm.t() + torch.eye(M) * 0.8
    Cm = torch.randn(S, S, generator=g)
    P0 = Cm @ Cm.t() + torch.eye(S) * 5.0
    init = {"P": P0, "F": F, "H": H, "Q": Q, "R": R, "mu_Q": torch.zeros(S), "mu_R": torch.randn(M, generator=g) * 0.1,
            "mu_v": torch.tensor([75.0])}
    st, _ = synth.vehicle_states(n, g)
    det = st[:, :5].clone()
    directions = st[:, 5].clone()
    times = torch.rand(n, generator=g).double() * 0.2
    dts = (torch.rand(n, generator=g).double() * 0.1 + 0.01)
    rows = torch.randperm(n, generator=g)[:m]
    z = det[rows] + torch.randn(m, M, generator=g) * 0.5
    return init, det, directions, times, dts, rows, z


def golden_kf():
    sys.path.insert(0, REF)
    from util_track.kf import Torch_KF
    sys.path.pop(0)
    init, det, directions, times, dts, rows, z = kf_inputs()
    kf = Torch_KF(torch.device("cpu"), INIT=init, ADD_MEAN_R=True)
    ids = list(range(100, 100 + len(det)))
    kf.add(det, ids, directions, times, init_speed=True)
    out = {"X0": kf.X.clone(), "P0": kf.P.clone(), "T0": kf.T.clone()}
    kf.predict()                                   # default dt, python float
    out.update(X1=kf.X.clone(), P1=kf.P.clone(), T1=kf.T.clone())
    kf.predict(dt=dts)                             # per-object float64 dt
    out.update(X2=kf.X.clone(), P2=kf.P.clone(), T2=kf.T.clone())
    kf.update(z, [ids[int(r)] for r in rows])
    out.update(X3=kf.X.clone(), P3=kf.P.clone())
    _, view = kf.view(dt=dts, with_direction=True)
    out.update(view=view)
    kf.remove([ids[3], ids[17]])
    kf.predict(dt=0.05)
    out.update(X4=kf.X.clone(), P4=kf.P.clone())
    save("kf", **out)


# ----------------------------------------------------------------------------------------------------------- anchors
ANCHOR_SHAPES_FULL = ((112, 112), (75, 133), (64, 80))          # stored completely
ANCHOR_SHAPES_HASHED = ((540, 960), (1080, 1920), (1001, 777))  # stored as sha256 of the float32 bytes + row count


def golden_anchors():
    """the reference's own Anchors.forward (retinanet/anchors.py:21-40; both copies are identical) on zero images"""
    import hashlib
    _, _, _, anchors3 = _import_retinanet(True)
    _, _, _, anchors2 = _import_retinanet(False)
    avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False                     # anchors.py:37 would call .cuda()
    try:
        out = {}
        for h, w in ANCHOR_SHAPES_FULL + ANCHOR_SHAPES_HASHED:
            a = anchors3.Anchors()(torch.zeros(1, 3, h, w))
            assert torch.equal(a, anchors2.Anchors()(torch.zeros(1, 3, h, w)))
            a = np.ascontiguousarray(_np(a)[0])
            if (h, w) in ANCHOR_SHAPES_FULL:
                out[f"anchors_{h}x{w}"] = a
            out[f"sha256_{h}x{w}"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8)
            out[f"count_{h}x{w}"] = np.int64(a.shape[0])
    finally:
        torch.cuda.is_available = avail
    save("anchors", **out)


if __name__ == "__main__":
    if "--only-anchors" in sys.argv:
        _shim()
        golden_anchors()
        sys.exit(0)
    if "--only-parse" in sys.argv:
        _shim()
        golden_parse()
        sys.exit(0)
    if "--only-best-box" in sys.argv:
        _shim()
        golden_best_box()
        sys.exit(0)
    if "--only-ts-bias" in sys.argv:
        _shim()
        golden_ts_bias()
        sys.exit(0)
    if not os.path.isdir(REF):
        sys.exit("needs the reference checkout at /root/reference")
    torch.manual_seed(0)
    torch.set_num_threads(4)
    _shim()
    golden_losses()
    golden_decode()
    golden_nms()
    golden_postprocess()
    golden_homography()
    golden_tracker()
    golden_kf()
    golden_anchors()
    golden_ts_bias()
    golden_best_box()
    golden_parse()
