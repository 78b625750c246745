"""CPU: the oracle (oracle/*.py) against vectors produced by the unmodified reference (tests/golden/make_golden.py).
Index / integer / bit-pattern results must be identical; float results are compared at 1e-6 relative (the oracle uses
the same ATen ops in the same order, so they are in fact bit-identical almost everywhere)."""
import numpy as np
import pytest
import torch

import synth
import os

from conftest import GOLDEN, assert_close_rel, load_golden
from oracle import anchors_oracle, decode_oracle, homography_oracle as ho, losses_oracle as lo, nms_oracle, tracker_oracle


@pytest.mark.parametrize("name", ["loss3d", "loss2d"])
def test_losses_forward_backward(golden, name):
    gd = golden(name)
    three_d = name == "loss3d"
    cls = gd["classification"].clone().requires_grad_(True)
    reg = gd["regression"].clone().requires_grad_(True)
    out = lo.focal_loss(cls, reg, gd["anchors"], gd["annotations"])
    losses, info = out[:-1], out[-1]
    assert len(losses) == (3 if three_d else 2)
    got = torch.cat([l.detach() for l in losses])
    assert_close_rel(got, gd["losses"], 1e-6, "losses")
    for j, (iou_max, iou_arg, pos, neg, _) in enumerate(info):
        assert torch.equal(iou_max, gd["iou_max"][j]), f"IoU_max image {j} not bit-exact"
        assert torch.equal(iou_arg, gd["iou_argmax"][j]), f"IoU_argmax image {j}"
    w = gd["grad_weights"]
    sum(l.sum() * float(w[i]) for i, l in enumerate(losses)).backward()
    assert_close_rel(cls.grad, gd["dcls"], 1e-5, "dcls")
    assert_close_rel(reg.grad, gd["dreg"], 1e-5, "dreg")


@pytest.mark.parametrize("name", ["loss3d", "loss2d"])
def test_calc_iou_bit_exact(golden, name):
    gd = golden(name)
    ann = gd["annotations"][0]
    rows = lo.valid_rows(ann[:, :21] if name == "loss3d" else ann, name == "loss3d")
    b = rows[:, 16:20] if name == "loss3d" else rows[:, :4]
    assert torch.equal(lo.calc_iou(gd["anchors"][0], b), gd["iou_matrix0"])


def test_all_empty_batch_raises_like_reference():
    anc = synth.anchors(64, 64)
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, synth.gen(1))
    ann = -torch.ones(2, 3, 27)
    with pytest.raises(RuntimeError):
        lo.focal_loss(cls, reg, anc, ann)


def test_decode(golden):
    gd = golden("decode")
    assert torch.equal(decode_oracle.decode3d(gd["anchors"], gd["regression"]), gd["decoded3d"])
    d2 = decode_oracle.decode2d(gd["anchors"], gd["deltas"])
    assert torch.equal(d2, gd["decoded2d"])
    h, w = gd["image_hw"].tolist()
    assert torch.equal(decode_oracle.clip(d2, h, w), gd["clipped2d"])


def test_nms(golden):
    gd = golden("nms")
    for thr in (0.5, 0.3, 0.1, 0.8, 0.2):
        assert torch.equal(nms_oracle.nms(gd["boxes"], gd["scores"], thr), gd[f"keep_{thr}"]), f"thr {thr}"
    assert torch.equal(nms_oracle.batched_nms(gd["boxes"], gd["scores"], gd["idxs"], 0.5), gd["keep_batched_0.5"])
    assert torch.equal(nms_oracle.nms(gd["eq_boxes"], gd["eq_scores"], 0.5), gd["eq_keep_0.5"])
    assert gd["eq_keep_0.5"].tolist() == [0, 1, 2]        # IoU == threshold is kept


def test_nms_matches_installed_torchvision_random():
    from torchvision.ops import nms as tv_nms
    for seed in range(5):
        b, s = synth.clustered_boxes(400, synth.gen(100 + seed))
        s = (s * 20).round() / 20           # many ties
        for thr in (0.5, 0.3):
            assert torch.equal(nms_oracle.nms(b, s, thr), tv_nms(b, s, thr))


def test_postprocess_3d(golden):
    gd = golden("post3d")
    h, w = gd["image_hw"].tolist()
    anc = synth.anchors(h, w)
    tr = decode_oracle.decode3d(anc, gd["regression"])
    assert np.array_equal(synth.digest(tr), gd["loc_boxes_digest"].numpy())
    s, c, b = nms_oracle.detect_3d(gd["classification"][:1], tr[:1])
    assert torch.equal(s, gd["scores"]) and torch.equal(c, gd["classes"]) and torch.equal(b, gd["boxes"])
    s, c, b, im = nms_oracle.detect_multi_frame(gd["classification"], tr)
    assert torch.equal(s, gd["mf_scores"]) and torch.equal(c, gd["mf_classes"])
    assert torch.equal(b, gd["mf_boxes"]) and torch.equal(im, gd["mf_im"])


def test_postprocess_3d_dense_ladder(golden):
    gd = golden("post3d")
    h, w = gd["dense_hw"].tolist()
    cls, reg = synth.dense_detection_inputs(int(gd["dense_seed"][0]), h, w)
    tr = decode_oracle.decode3d(synth.anchors(h, w), reg)
    s, c, b = nms_oracle.detect_3d(cls, tr)
    assert np.array_equal(np.bincount(c.numpy(), minlength=8), gd["dense_count"].numpy())
    assert np.array_equal(synth.digest(s), gd["dense_scores_digest"].numpy())
    assert np.array_equal(synth.digest(b), gd["dense_boxes_digest"].numpy())


def test_postprocess_2d(golden):
    gd = golden("post2d")
    h, w = gd["image_hw"].tolist()
    tr = decode_oracle.clip(decode_oracle.decode2d(synth.anchors(h, w), gd["regression"]), h, w)
    s, c, b = nms_oracle.detect_2d(gd["classification"], tr)
    assert torch.equal(s, gd["scores"]) and torch.equal(c, gd["classes"]) and torch.equal(b, gd["boxes"])


def test_ladder_rungs_are_float32_casts_of_repeated_products():
    mask, last = nms_oracle.ladder_threshold(torch.tensor([0.5, 0.2, 0.9]), 1e-25, keep=1)
    assert mask.tolist() == [False, False, True]
    t = 1e-25
    while np.float32(t) != last:
        t *= 10 ** .2
    assert 0.5 <= t < 0.9


def test_homography_csv_rows(golden):
    """the reference's own stored outputs: 3D_tracking_results.csv"""
    gd = golden("homography")
    st = gd["csv_states"]
    space = ho.state_to_space(st)
    assert torch.equal(space[:, :4, :2].reshape(-1, 8), gd["csv_space"]), "state_to_space vs CSV cols 27-34 (bit-exact)"
    im = ho.wrapper_state_to_im(st, gd["csv_P_lo"], gd["csv_P_hi"])
    # the CSV stores states printed as float32 and P is a DLT fit: agreement to ~1e-6 relative
    assert_close_rel(im, gd["csv_im"], 5e-6, "state_to_im vs CSV cols 11-26")


def test_homography_transforms(golden):
    gd = golden("homography")
    P, H = gd["P"].numpy(), gd["H"].numpy()
    st, cam = gd["states"], gd["cam"].long()
    assert torch.equal(ho.state_to_space(st), gd["space"])
    tol = 1e-12
    assert_close_rel(ho.state_to_im(st, P[1, 0]), gd["im_single"], tol, "state_to_im single")
    Pl = torch.from_numpy(P)[cam]
    assert_close_rel(ho.state_to_im(st, Pl[:, 0]), gd["im_list"], tol, "state_to_im list")
    assert_close_rel(ho.wrapper_state_to_im(st, Pl[:, 0], Pl[:, 1]), gd["im_wrapper_list"], tol, "wrapper list")
    assert_close_rel(ho.wrapper_state_to_im(st, P[2, 0], P[2, 1]), gd["im_wrapper_single"], tol, "wrapper single")
    assert_close_rel(ho.wrapper_space_to_im(gd["space"], Pl[:, 0], Pl[:, 1]), gd["space_to_im_f32pts"], tol, "space_to_im")
    det, hts = gd["det"], gd["heights"]
    Hl = torch.from_numpy(H)[cam]
    assert_close_rel(ho.im_to_space(det, Hl[:, 0], hts), gd["space_from_im_list"], tol, "im_to_space list")
    assert_close_rel(ho.wrapper_im_to_space(det, Hl[:, 0], Hl[:, 1], hts), gd["space_from_im_wrapper"], tol, "im_to_space wrapper")
    assert_close_rel(ho.im_to_state(det, H[1, 0], hts), gd["state_single"], 1e-6, "im_to_state single")
    assert_close_rel(ho.im_to_state(det, Hl[:, 0], hts), gd["state_list"], 1e-6, "im_to_state list")
    assert_close_rel(ho.wrapper_im_to_state(det, Hl[:, 0], Hl[:, 1], hts), gd["state_wrapper_list"], 1e-6, "im_to_state wrapper")
    assert torch.equal(ho.space_to_state(gd["space_from_im_list"]), gd["state_from_space"])
    assert torch.equal(ho.space_to_state(gd["space"]), gd["state_from_space_f32"])
    s1, h1 = ho.refined_im_to_state(det, Hl[:, 0], Hl[:, 1], Pl[:, 0], Pl[:, 1], hts)
    assert_close_rel(h1, gd["hft_f64_f32_f32"], 1e-9, "refined heights")
    assert_close_rel(s1, gd["state_refined"], 1e-6, "refined state")


def test_tracker(golden):
    gd = golden("tracker")
    st, sec, sc = gd["states"], gd["second"], gd["scores"]
    assert torch.equal(tracker_oracle.footprint(st), gd["footprint"])
    cost = tracker_oracle.association_cost(st, sec)
    assert torch.equal(cost, gd["cost"])
    assert torch.isnan(tracker_oracle.md_iou(torch.ones(1, 1, 4, dtype=torch.float64), torch.ones(1, 1, 4, dtype=torch.float64))).all()
    assert torch.isnan(gd["md_iou_degenerate"]).all()
    assert torch.equal(tracker_oracle.space_nms(st, sc, 0.1), gd["space_nms_0_1"])
    assert torch.equal(tracker_oracle.space_nms(st, sc, 0.4), gd["space_nms_0_4"])
    assert torch.equal(tracker_oracle.im_nms(gd["corners"], sc, 0.3), gd["im_nms_0_3"])
    assert torch.equal(tracker_oracle.im_nms(gd["corners"], sc, 0.3, groups=torch.zeros(120)), gd["im_nms_groups"])


def test_kf_oracle_matches_reference_golden():
    """oracle/kf_oracle.py against the UNMODIFIED Torch_KF (tests/golden/kf.npz): default-dt predict, per-object float64 dt
    predict, update, view"""
    import importlib.util
    from oracle import kf_oracle as ko
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    init, det, directions, times, dts, rows, z = mg.kf_inputs()
    gd = load_golden("kf")
    F, Q, H, R, mu_R = init["F"].float(), init["Q"].float(), init["H"].float(), init["R"].float(), init["mu_R"].float()
    X, P = gd["X0"], gd["P0"]
    X, P = ko.predict(X, P, directions, 1 / 30.0, F, Q)
    assert torch.allclose(X, gd["X1"], rtol=1e-6, atol=0) and torch.allclose(P, gd["P1"], rtol=1e-6, atol=1e-6)
    X, P = ko.predict(X, P, directions, dts, F, Q)
    assert torch.allclose(X, gd["X2"], rtol=1e-6, atol=0) and torch.allclose(P, gd["P2"], rtol=1e-6, atol=1e-6)
    X3, P3 = ko.update(X, P, rows, z, H, R, mu_R)
    assert torch.allclose(X3, gd["X3"], rtol=1e-6, atol=1e-6) and torch.allclose(P3, gd["P3"], rtol=1e-5, atol=1e-5)
    Xv, _ = ko.predict(X3, P3, directions, dts, F, Q)
    view = torch.cat((Xv[:, :-1], directions.float().unsqueeze(1), Xv[:, -1:]), dim=1)
    assert torch.allclose(view, gd["view"], rtol=1e-6, atol=1e-6)


def test_anchors_oracle_and_host_generator_equal_reference_tables():
    """anchors.py:21-40: every stored table and every sha256 of the reference's Anchors.forward, bit for bit"""
    import hashlib
    from geom3d_b200.anchors_impl import anchors_for_image
    gd = np.load(os.path.join(GOLDEN, "anchors.npz"))
    shapes = [tuple(map(int, k[7:].split("x"))) for k in gd.files if k.startswith("sha256_")]
    assert len(shapes) == 6
    for h, w in shapes:
        for table in (anchors_oracle.anchors(h, w), anchors_for_image(h, w)):
            assert table.dtype == np.float32 and table.shape == (int(gd[f"count_{h}x{w}"]), 4)
            assert hashlib.sha256(np.ascontiguousarray(table).tobytes()).digest() == gd[f"sha256_{h}x{w}"].tobytes(), (h, w)
            if f"anchors_{h}x{w}" in gd.files:
                assert np.array_equal(table, gd[f"anchors_{h}x{w}"])


def test_estimate_ts_bias(golden):
    """MC3D_crop_tracker.py:237-315: the oracle's bias lists equal the unmodified method's, call after call"""
    gd = golden("ts_bias")
    boxes, cams, objs = gd["boxes"], gd["cams"], gd["objs"]
    ts = [float(t) for t in gd["timestamps"]]
    assert tracker_oracle.cross_camera_pairs(boxes, cams, 0.1).shape[0] > 20
    for tag, view in (("both", objs), ("eastbound_only", objs[objs[:, 5] == 1])):
        b1 = tracker_oracle.estimate_ts_bias(boxes, cams, view, ts, [0, 0, 0, 0], 105.0, 0.1, 0.05)
        assert b1 == gd[f"bias1_{tag}"].tolist(), tag
        ts2 = [t + 1 / 30.0 + 0.001 * k for k, t in enumerate(ts)]
        b2 = tracker_oracle.estimate_ts_bias(boxes, cams, view, ts2, b1, 105.0, 0.1, 0.05)
        assert b2 == gd[f"bias2_{tag}"].tolist(), tag


def test_select_best_box_and_evaluator_iou(golden):
    """tracker_oracle.select_best_box / pairwise_iou_eps against the UNMODIFIED MC_Crop_Tracker.select_best_box
    (MC3D_crop_tracker.py:974-1028) and MOT_Evaluator.iou (mot_evaluator.py:87-118) - tests/golden/best_box.npz"""
    gd = golden("best_box")
    n = gd["confs"].shape[0]
    for W, tag in ((0.4, "0_4"), (0.0, "0_0"), (1.0, "1_0")):
        best, cls, cf = tracker_oracle.select_best_box(gd["prior"], gd["preds"], gd["confs"], gd["classes"], n, W)
        assert torch.equal(best, gd[f"best_{tag}"]) and torch.equal(cls, gd[f"cls_{tag}"]) and torch.equal(cf, gd[f"conf_{tag}"])
    got = tracker_oracle.pairwise_iou_eps(gd["eval_a"], gd["eval_b"], 1e-6)
    assert torch.equal(got, gd["eval_iou"]), float((got - gd["eval_iou"]).abs().max())


def test_postprocess_batch_flattening_equals_loop_over_the_reference_expressions():
    """B > 1 in the default branch: the reference's squeeze / boolean-mask expressions (3D model.py:365-395) evaluated
    literally on a small batch equal the oracle's flattened form"""
    g = synth.gen(55)
    B, A, C = 3, 400, 4
    cls = synth.detection_scores(B, A, C, g, objects=6, per_object=7)
    boxes = torch.rand(B, A, 20, generator=g) * 100
    boxes[..., 18:20] = boxes[..., 16:18] + 5 + torch.rand(B, A, 2, generator=g) * 30
    S, K, X = [], [], []
    for i in range(C):                                             # the reference's loop, names and all
        scores = torch.squeeze(cls[:, :, i])
        keep, keep_count, threshold = 10000, 1000000, 1e-25
        while keep_count > keep:
            scores_over_thresh = (scores > threshold)
            keep_count = scores_over_thresh.sum()
            threshold *= (10 ** .2)
        if scores_over_thresh.sum() == 0:
            continue
        scores = scores[scores_over_thresh]
        anchorBoxes = torch.squeeze(boxes)[scores_over_thresh]
        idx = nms_oracle.nms(anchorBoxes[:, 16:20], scores, 0.5)
        S.append(scores[idx]); K.append(torch.tensor([i] * idx.shape[0])); X.append(anchorBoxes[idx])
    got = nms_oracle.detect_3d(cls, boxes)
    assert torch.equal(got[0], torch.cat(S)) and torch.equal(got[1], torch.cat(K)) and torch.equal(got[2], torch.cat(X))


def test_parse_detections_and_remove_overlaps(golden):
    """tracker_oracle.parse_detections / remove_overlaps against the UNMODIFIED MC_Crop_Tracker.parse_detections
    (MC3D_crop_tracker.py:319-383; integer labels -> the "other" height 5 from guess_heights) and the nms core of
    remove_overlaps (:495-508) - tests/golden/parse.npz"""
    gd = golden("parse")
    H, P = gd["H"], gd["P"]
    heights = torch.full((gd["scores"].shape[0],), 5.0)
    for tag, kw in (("nms", dict(perform_nms=True, refine_height=False)), ("nms_refined", dict(perform_nms=True, refine_height=True)),
                    ("plain", dict(perform_nms=False, refine_height=False))):
        st, lb, sc, cm = tracker_oracle.parse_detections(gd["scores"], gd["labels"], gd["boxes"], gd["cam"], H[:, 0], H[:, 1],
                                                          P[:, 0], P[:, 1], 0.35, 0.3, 0.1, heights, **kw)
        assert torch.equal(lb, gd[f"labels_{tag}"]) and torch.equal(sc, gd[f"scores_{tag}"]) and torch.equal(cm, gd[f"cams_{tag}"])
        assert_close_rel(st, gd[f"states_{tag}"], 1e-6, f"states {tag}")
    keep = tracker_oracle.remove_overlaps(gd["overlap_states"], gd["overlap_alive"], 0.2)
    assert torch.equal(keep, gd["overlap_keep"])
