"""GPU parity of the tracker-frame widening: estimate_ts_bias's pair mining + bias update against the
unmodified method's golden vectors and the oracle, and the one-graph frame (`FrameGeometry`) against the separate calls
and the oracle.  Pair lists, keep lists and bias floats must be identical; the cost matrix is FP64 of the same formula."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import synth
from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


def test_cross_camera_pairs_vs_oracle():
    from geom3d_b200 import tracker_geometry as tg
    from oracle import tracker_oracle as to
    for seed, d, n_cams, spread in ((1, 300, 5, 600.0), (2, 65, 2, 150.0), (3, 1000, 18, 2500.0), (4, 33, 3, 1e6)):
        g = synth.gen(900 + seed)
        st, _ = synth.vehicle_states(d, g, n_cams=1)
        st[:, 0] = 100 + torch.rand(d, generator=g) * spread
        cams = torch.randint(0, n_cams, (d,), generator=g)
        exp = to.cross_camera_pairs(st, cams, 0.1)
        got = tg.cross_camera_pairs(st.cuda(), cams.cuda(), 0.1)
        assert got.dtype == torch.int64 and torch.equal(got.cpu(), exp), (seed, exp.shape, got.shape)
        assert torch.equal(tg.cross_camera_pairs(st, cams, 0.1), exp)                      # CPU tensors in, CPU out
    one = synth.vehicle_states(1, synth.gen(5), n_cams=1)[0]
    assert tg.cross_camera_pairs(one.cuda(), torch.zeros(1).cuda(), 0.1).shape == (0, 2)
    assert tg.cross_camera_pairs(torch.zeros(0, 6).cuda(), torch.zeros(0).cuda(), 0.1).shape == (0, 2)
    same = one.repeat(40, 1)                                                                # one camera: no pair qualifies
    assert tg.cross_camera_pairs(same.cuda(), torch.zeros(40).cuda(), 0.1).shape == (0, 2)
    assert tg.cross_camera_pairs(same.cuda(), torch.arange(40).cuda() % 2, 0.1).shape == (400, 2)


def test_estimate_ts_bias_golden():
    """MC3D_crop_tracker.py:237-315: bias lists equal to the unmodified method's floats, two consecutive calls"""
    from geom3d_b200 import tracker_geometry as tg
    gd = load_golden("ts_bias")
    boxes, cams, objs = gd["boxes"], gd["cams"], gd["objs"]
    ts = [float(t) for t in gd["timestamps"]]
    for tag, view in (("both", objs), ("eastbound_only", objs[objs[:, 5] == 1])):
        for put in (lambda t: t.cuda(), lambda t: t):
            bias = [0, 0, 0, 0]
            assert tg.estimate_ts_bias(put(boxes), put(cams), put(view), ts, bias, 105.0, 0.1, 0.05) is bias
            assert bias == gd[f"bias1_{tag}"].tolist(), tag
            ts2 = [t + 1 / 30.0 + 0.001 * k for k, t in enumerate(ts)]
            tg.estimate_ts_bias(put(boxes), put(cams), put(view), ts2, bias, 105.0, 0.1, 0.05)
            assert bias == gd[f"bias2_{tag}"].tolist(), tag
    bias = [0.5, 0.25]
    assert tg.estimate_ts_bias(boxes[:0], cams[:0], objs, ts, bias, 105.0) == [0.5, 0.25]   # no detections: untouched
    assert tg.estimate_ts_bias(boxes, cams, objs[:0], ts, bias, 105.0) == [0.5, 0.25]       # empty filter: untouched


@pytest.mark.parametrize("graph", [True, False])
def test_frame_geometry_equals_separate_calls_and_oracle(graph):
    from geom3d_b200 import ops, tracker_geometry as tg
    from oracle import homography_oracle as ho, tracker_oracle as to
    P, _ = synth.camera_matrices(6)
    Pd = torch.from_numpy(P).cuda()
    frame = tg.FrameGeometry(Pd, capacity_pre=640, capacity_det=700, phi_space=0.1, phi_im=0.3, graph=graph)
    for k, (n_pre, n_det) in enumerate(((500, 700), (640, 300), (37, 41), (1, 1), (300, 0), (0, 129), (600, 699))):
        g = synth.gen(700 + k)
        det, cam = synth.vehicle_states(n_det, g, n_cams=6)
        det[:, 0] = 100 + torch.rand(n_det, generator=g) * 900                 # dense enough for both NMS to bite
        pre = det[torch.randint(0, max(n_det, 1), (n_pre,), generator=g)].clone() if n_det else synth.vehicle_states(n_pre, g)[0]
        pre[:, :2] += torch.randn(n_pre, 2, generator=g) * torch.tensor([3.0, 0.5])
        sc = torch.rand(n_det, generator=g)
        out = frame(pre.cuda(), det.cuda(), sc.cuda(), cam.cuda())
        assert out["cost"].shape == (n_pre, n_det)
        if n_det == 0:
            assert out["space_keep"].numel() == 0 and out["im_keep"].numel() == 0 and out["corners"].shape == (0, 8, 2)
            continue
        # the separate drop-in calls
        assert torch.equal(out["space_keep"], tg.space_nms(det.cuda(), sc.cuda(), 0.1)), (k, "space")
        corners = ops.state_to_im(det.cuda(), Pd, cam.cuda(), wrapper=True)
        assert torch.equal(out["corners"], corners)
        assert torch.equal(out["im_keep"], tg.im_nms(corners, sc.cuda(), 0.3)), (k, "im")
        if n_pre and n_det:
            assert torch.equal(out["cost"], tg.association_cost(pre.cuda(), det.cuda()))
        # the oracle
        if n_pre and n_det and k < 3:
            assert torch.equal(out["cost"].cpu(), to.association_cost(pre, det))
            assert torch.equal(out["space_keep"].cpu(), to.space_nms(det, sc, 0.1))
            Pc = torch.from_numpy(P)[cam.long()]
            assert torch.equal(out["im_keep"].cpu(), to.im_nms(ho.wrapper_state_to_im(det, Pc[:, 0], Pc[:, 1]), sc, 0.3))
    with pytest.raises(ValueError):
        frame(torch.zeros(641, 6).cuda(), torch.zeros(1, 6).cuda(), torch.zeros(1).cuda(), torch.zeros(1).cuda())


def test_select_best_box_golden_and_oracle():
    """select_best_box (MC3D_crop_tracker.py:974-1028): identical rows to the unmodified method's (golden) for three
    weights, CPU and CUDA inputs, and to the oracle on a larger seeded case"""
    from geom3d_b200 import tracker_geometry as tg
    from oracle import tracker_oracle as to
    gd = load_golden("best_box")
    n = gd["confs"].shape[0]
    for W, tag in ((0.4, "0_4"), (0.0, "0_0"), (1.0, "1_0")):
        for put in (lambda t: t.cuda(), lambda t: t):
            best, cls, cf = tg.select_best_box(put(gd["prior"]), put(gd["preds"].reshape(-1, 6)), put(gd["confs"]),
                                               put(gd["classes"]), n, W)
            assert torch.equal(best.cpu(), gd[f"best_{tag}"]) and torch.equal(cls.cpu(), gd[f"cls_{tag}"])
            assert torch.equal(cf.cpu(), gd[f"conf_{tag}"])
    g = synth.gen(93)
    n, d = 700, 13
    prior, _ = synth.vehicle_states(n, g, n_cams=1)
    preds = prior.unsqueeze(1).repeat(1, d, 1)
    preds[:, :, :2] += torch.randn(n, d, 2, generator=g) * torch.tensor([8.0, 2.0])
    confs, classes = torch.rand(n, d, generator=g), torch.randint(0, 8, (n, d), generator=g)
    want = to.select_best_box(prior, preds, confs, classes, n, 0.3)
    got = tg.select_best_box(prior.cuda(), preds.cuda(), confs.cuda(), classes.cuda(), n, 0.3)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(got, want))


def test_evaluator_iou_with_union_epsilon():
    """the scalar iou of mot_evaluator.py:87-118 adds 1e-6 to the union; g3d_pairwise_iou_f64(eps) against the
    unmodified function's values (golden, float64 boxes) and the oracle on float32 boxes incl. degenerate ones"""
    from geom3d_b200 import tracker_geometry as tg
    from oracle import tracker_oracle as to
    gd = load_golden("best_box")
    a, b = gd["eval_a"], gd["eval_b"]
    # the kernel takes float32 boxes (promoted to float64 inside, as the trackers' .double()): compare on boxes that are
    # exactly representable, i.e. the float32 roundings of the golden boxes, through the oracle pinned by the golden file
    assert torch.equal(to.pairwise_iou_eps(a, b), gd["eval_iou"])
    a32, b32 = a.float(), b.float()
    got = tg.pairwise_iou(a32.cuda(), b32.cuda(), eps=1e-6)
    assert got.dtype == torch.float64 and torch.equal(got.cpu(), to.pairwise_iou_eps(a32, b32, 1e-6))
    deg = torch.tensor([[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 2.0, 2.0]])
    got = tg.pairwise_iou(deg.cuda(), deg.cuda(), eps=1e-6).cpu()
    assert float(got[0, 0]) == 0.0 and not torch.isnan(got).any()       # 0 / 1e-6, not the NaN of the eps-free md_iou
    assert torch.isnan(tg.pairwise_iou(deg.cuda(), deg.cuda()).cpu()[0, 0])


def test_cfg5_full_size_frame_vs_oracle():
    """BASELINE.json configs[4] at its full size: 2000 tracked objects x 2000 detections over 18 cameras - the cost matrix
    bit for bit, both keep lists identical to the oracle's"""
    from geom3d_b200 import tracker_geometry as tg
    from oracle import homography_oracle as ho, tracker_oracle as to
    P, _ = synth.camera_matrices(18)
    Pd = torch.from_numpy(P).cuda()
    g = synth.gen(7)                                                  # bench.py's cfg5 generator
    s5, c5 = synth.vehicle_states(2000, g)
    j5 = s5.clone()
    j5[:, :2] += torch.randn(2000, 2, generator=g) * torch.tensor([3.0, 0.5])
    sc5 = torch.rand(2000, generator=g)
    frame = tg.FrameGeometry(Pd, 2000, 2000, phi_space=0.1, phi_im=0.3)
    out = frame(s5.cuda(), j5.cuda(), sc5.cuda(), c5.cuda())
    assert torch.equal(out["cost"].cpu(), to.association_cost(s5, j5))
    assert torch.equal(out["space_keep"].cpu(), to.space_nms(j5, sc5, 0.1))
    Pc = torch.from_numpy(P)[c5.long()]
    corners = ho.wrapper_state_to_im(j5, Pc[:, 0], Pc[:, 1])
    assert torch.equal(out["im_keep"].cpu(), to.im_nms(corners, sc5, 0.3))
    assert 0 < out["space_keep"].numel() < 2000 and 0 < out["im_keep"].numel() < 2000


def _wrapper_from(P, H, n_cams):
    from geom3d_b200.homography_impl import Homography, Homography_Wrapper
    hg1, hg2 = Homography(), Homography()
    names = synth.CAMERAS[:n_cams]
    for i, n in enumerate(names):
        hg1.add_correspondence_matrices(n, H[i, 0].numpy(), P[i, 0].numpy())
        hg2.add_correspondence_matrices(n, H[i, 1].numpy(), P[i, 1].numpy())
    return Homography_Wrapper(hg1, hg2), names


def test_parse_detections_on_the_device_golden():
    """MC_Crop_Tracker.parse_detections (MC3D_crop_tracker.py:319-383) with the detections staying on the GPU
    - score cut, im_nms, im_to_state (+ fused two-pass height refinement), space_nms - against the unmodified method's
    output (golden): the same detections survive, in the same order; states within 1e-5"""
    from conftest import assert_close_rel
    from geom3d_b200 import tracker_geometry as tg
    gd = load_golden("parse")
    wr, names = _wrapper_from(gd["P"], gd["H"], gd["P"].shape[0])
    for tag, kw in (("nms", dict(perform_nms=True, refine_height=False)), ("nms_refined", dict(perform_nms=True, refine_height=True)),
                    ("plain", dict(perform_nms=False, refine_height=False))):
        for put in (lambda t: t.cuda(), lambda t: t):
            st, lb, sc, cm = tg.parse_detections(wr, put(gd["scores"]), put(gd["labels"]), put(gd["boxes"]), put(gd["cam"]), names,
                                                 sigma_d=0.35, phi_nms_im=0.3, phi_nms_space=0.1, **kw)
            assert torch.equal(lb.cpu(), gd[f"labels_{tag}"]) and torch.equal(sc.cpu(), gd[f"scores_{tag}"]), tag
            assert torch.equal(cm.cpu(), gd[f"cams_{tag}"])
            assert st.dtype == torch.float32
            assert_close_rel(st.cpu(), gd[f"states_{tag}"], 1e-5, f"states {tag}")
    assert tg.parse_detections(wr, gd["scores"].cuda() * 0, gd["labels"].cuda(), gd["boxes"].cuda(), gd["cam"].cuda(), names,
                               0.35, 0.3, 0.1) == ([], [], [], [])
    assert tg.parse_detections(wr, gd["scores"][:0], gd["labels"][:0], gd["boxes"][:0], gd["cam"][:0], names, 0.35, 0.3, 0.1) == ([], [], [], [])


def test_remove_overlaps_golden_and_oracle():
    """MC_Crop_Tracker.remove_overlaps (MC3D_crop_tracker.py:482-518): footprint NMS with the frames alive as confidence"""
    from geom3d_b200 import tracker_geometry as tg
    from oracle import tracker_oracle as to
    gd = load_golden("parse")
    keep, removed = tg.remove_overlaps(gd["overlap_states"].cuda(), gd["overlap_alive"].cuda(), 0.2)
    assert torch.equal(keep.cpu(), gd["overlap_keep"]) and int(removed.sum()) == 40 - gd["overlap_keep"].numel()
    keep0, removed0 = tg.remove_overlaps(gd["overlap_states"].cuda(), gd["overlap_alive"].cuda(), 0.0)
    assert keep0.numel() == 40 and not bool(removed0.any())                       # phi_over <= 0: nothing is removed (:489)
    g = synth.gen(97)
    st, _ = synth.vehicle_states(1500, g, n_cams=1)
    st[:, 0] = 100 + torch.rand(1500, generator=g) * 2000
    alive = torch.randint(1, 500, (1500,), generator=g)
    keep, _ = tg.remove_overlaps(st.cuda(), alive.cuda(), 0.3)
    assert torch.equal(keep.cpu(), to.remove_overlaps(st, alive, 0.3))
