"""GPU parity: homography transforms and tracker geometry through the drop-in classes / C ABI.
Bar: state_to_space bit-exact (float32 add/mul); projections within 1e-5 relative (they agree to ~1e-12: FP64)."""
import numpy as np
import pytest
import torch

import synth
from conftest import assert_close_rel

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _hgs(P, H, names):
    from geom3d_b200.homography_impl import Homography, Homography_Wrapper
    hg1, hg2 = Homography(), Homography()
    for i, n in enumerate(names):
        hg1.add_correspondence_matrices(n, H[i, 0], P[i, 0])
        hg2.add_correspondence_matrices(n, H[i, 1], P[i, 1])
    return hg1, hg2, Homography_Wrapper(hg1, hg2)


def test_csv_rows_of_the_reference(golden):
    gd = golden("homography")
    from geom3d_b200.homography_impl import Homography, Homography_Wrapper
    hg1, hg2 = Homography(), Homography()
    hg1.add_correspondence_matrices("p1c1", np.eye(3), gd["csv_P_lo"].numpy())
    hg2.add_correspondence_matrices("p1c1", np.eye(3), gd["csv_P_hi"].numpy())
    st = gd["csv_states"].cuda()
    space = hg1.state_to_space(st)
    assert space.dtype == torch.float32 and space.shape == (st.shape[0], 8, 3)
    assert torch.equal(space[:, :4, :2].reshape(-1, 8).cpu(), gd["csv_space"]), "state_to_space vs CSV (bit-exact)"
    im = Homography_Wrapper(hg1, hg2).state_to_im(st, name="p1c1")
    assert im.dtype == torch.float64
    assert_close_rel(im.cpu(), gd["csv_im"], 5e-6, "state_to_im vs CSV")


def test_transforms_golden(golden):
    gd = golden("homography")
    names = synth.CAMERAS[:3]
    hg1, hg2, wr = _hgs(gd["P"].numpy(), gd["H"].numpy(), names)
    st, cam = gd["states"].cuda(), gd["cam"]
    cam_names = [names[i] for i in cam.tolist()]
    assert torch.equal(hg1.state_to_space(st).cpu(), gd["space"])
    assert_close_rel(hg1.state_to_im(st, name=names[1]).cpu(), gd["im_single"], TOL, "state_to_im single")
    assert_close_rel(hg1.state_to_im(st, name=cam_names).cpu(), gd["im_list"], TOL, "state_to_im list of names")
    assert_close_rel(hg1.state_to_im(st, name=cam.cuda()).cpu(), gd["im_list"], TOL, "state_to_im index tensor")
    assert_close_rel(wr.state_to_im(st, name=cam_names).cpu(), gd["im_wrapper_list"], TOL, "wrapper list")
    assert_close_rel(wr.state_to_im(st, name=names[2]).cpu(), gd["im_wrapper_single"], TOL, "wrapper single")
    assert_close_rel(wr.space_to_im(gd["space"].cuda(), name=cam_names).cpu(), gd["space_to_im_f32pts"], TOL, "space_to_im")
    det, hts = gd["det"].cuda(), gd["heights"].cuda()
    assert_close_rel(hg1.im_to_space(det, name=cam_names, heights=hts).cpu(), gd["space_from_im_list"], TOL, "im_to_space")
    assert_close_rel(wr.im_to_space(det, name=cam_names, heights=hts).cpu(), gd["space_from_im_wrapper"], TOL, "im_to_space wrapper")
    for got, key in ((hg1.im_to_state(det, name=names[1], heights=hts), "state_single"),
                     (hg1.im_to_state(det, name=cam_names, heights=hts), "state_list"),
                     (wr.im_to_state(det, name=cam_names, heights=hts), "state_wrapper_list")):
        assert got.dtype == torch.float32 and got.shape == (300, 6)
        assert_close_rel(got.cpu(), gd[key], TOL, key)
    assert_close_rel(hg1.space_to_state(gd["space_from_im_list"].cuda()).cpu(), gd["state_from_space"], 1e-6, "space_to_state f64")
    assert torch.equal(hg1.space_to_state(gd["space"].cuda()).cpu(), gd["state_from_space_f32"])
    assert hg1.im_to_space(det, name=names[0]) is None            # no heights: prints and returns None
    # height_from_template in the three dtype mixes the trackers produce
    repro = wr.state_to_im(wr.im_to_state(det, heights=hts, name=cam_names), name=cam_names)
    h_a = wr.height_from_template(repro, hts, det)
    assert h_a.dtype == torch.float64
    assert_close_rel(h_a.cpu(), gd["hft_f64_f32_f32"], TOL, "height_from_template f64/f32/f32")
    assert_close_rel(hg1.height_from_template(repro, hts.double(), det.double()).cpu(), gd["hft_all_f64"], TOL, "hft f64")
    h_c = hg1.height_from_template(repro.float(), hts, det)
    assert h_c.dtype == torch.float32
    assert_close_rel(h_c.cpu(), gd["hft_all_f32"], TOL, "hft f32")
    # fused two-pass refinement
    s1, h1 = wr.im_to_state_refined(det, name=cam_names, heights=hts, return_heights=True)
    assert_close_rel(h1.cpu(), gd["hft_f64_f32_f32"], TOL, "refined heights")
    assert_close_rel(s1.cpu(), gd["state_refined"], TOL, "refined state")
    # CPU callers get CPU results computed by the kernels
    cpu = wr.state_to_im(gd["states"], name=cam_names)
    assert not cpu.is_cuda
    assert_close_rel(cpu, gd["im_wrapper_list"], TOL, "cpu round trip")


def test_round_trip_and_all_cameras_vs_oracle():
    from oracle import homography_oracle as ho
    P, H = synth.camera_matrices(18)
    hg1, hg2, wr = _hgs(P, H, synth.CAMERAS)
    g = synth.gen(2)
    st, cam = synth.vehicle_states(20000, g)
    im = wr.state_to_im(st.cuda(), name=cam.cuda())
    Pl = torch.from_numpy(P)[cam.long()]
    assert_close_rel(im.cpu(), ho.wrapper_state_to_im(st, Pl[:, 0], Pl[:, 1]), TOL, "state_to_im vs oracle")
    # im -> state -> im round trip of the bottom corners (Homography.test_transformation's property, homography.py:554-604)
    back = hg1.im_to_state(hg1.state_to_im(st.cuda(), name=cam.cuda()), name=cam.cuda(), heights=st[:, 4].cuda())
    assert_close_rel(back[:, :4].cpu(), st[:, :4], 1e-4, "round trip x,y,l,w")
    assert torch.equal(back[:, 5].cpu(), st[:, 5])
    allc = wr.state_to_im_all(st[:500].cuda())
    assert allc.shape == (500, 18, 8, 2)
    for c in (0, 7, 17):
        exp = ho.wrapper_state_to_im(st[:500], P[c, 0], P[c, 1])
        assert_close_rel(allc[:, c].cpu(), exp, TOL, f"all-cameras cam {c}")
    f32 = wr.state_to_im(st.cuda(), name=cam.cuda(), out_dtype=torch.float32)
    assert f32.dtype == torch.float32
    assert_close_rel(f32.cpu(), im.cpu().float(), 1e-6, "float32 opt-in output")
    assert wr.state_to_im(torch.zeros(0, 6).cuda(), name="p1c1").shape == (0, 8, 2)


def test_tracker_geometry_golden(golden):
    from geom3d_b200 import tracker_geometry as tg
    gd = golden("tracker")
    st, sec, sc = gd["states"].cuda(), gd["second"].cuda(), gd["scores"].cuda()
    assert torch.equal(tg.state_footprint(st).cpu(), gd["footprint"])
    cost = tg.association_cost(st, sec)
    assert cost.dtype == torch.float64 and torch.equal(cost.cpu(), gd["cost"]), "association matrix is bit-exact (FP64)"
    assert torch.equal(tg.pairwise_iou(tg.state_footprint(st), tg.state_footprint(sec)).cpu(), gd["md_iou"])
    fa = gd["footprint"].double().cuda()
    lit = tg.md_iou(fa.unsqueeze(1).repeat(1, 100, 1), tg.state_footprint(sec).double().unsqueeze(0).repeat(120, 1, 1))
    assert torch.equal(lit.cpu(), gd["md_iou"])
    deg = torch.ones(1, 1, 4, dtype=torch.float64).cuda()
    assert torch.isnan(tg.md_iou(deg, deg)).all()                       # 0/0 -> NaN, no epsilon
    assert torch.equal(tg.space_nms(st, sc, 0.1).cpu(), gd["space_nms_0_1"])
    assert torch.equal(tg.space_nms(st, sc, 0.4).cpu(), gd["space_nms_0_4"])
    assert torch.equal(tg.im_nms(gd["corners"].cuda(), sc, 0.3).cpu(), gd["im_nms_0_3"])
    assert torch.equal(tg.im_nms(gd["corners"].cuda(), sc, 0.3, groups=torch.zeros(120).cuda()).cpu(), gd["im_nms_groups"])
    s = tg.self_iou(st)
    assert torch.equal(s.cpu(), s.cpu().t())


def test_full_size_tracking_frame_properties():
    """BASELINE config 5: 2000 objects; association matrix properties + NMS idempotence"""
    from geom3d_b200 import tracker_geometry as tg
    g = synth.gen(6)
    st, _ = synth.vehicle_states(2000, g)
    jit = st.clone()
    jit[:, :2] += torch.randn(2000, 2, generator=g) * torch.tensor([3.0, 0.5])
    cost = tg.association_cost(st.cuda(), jit.cuda())
    assert cost.shape == (2000, 2000)
    iou = 1.0 - cost
    assert bool(((iou >= 0) & (iou <= 1)).all())
    assert float(iou.diagonal().mean()) > 0.3 and torch.equal(tg.self_iou(st.cuda()).diagonal(), torch.ones(2000, dtype=torch.float64).cuda())
    sc = torch.rand(2000, generator=g).cuda()
    keep = tg.space_nms(st.cuda(), sc, 0.1)
    again = tg.space_nms(st.cuda()[keep], sc[keep], 0.1)
    assert torch.equal(again, torch.arange(keep.numel(), device="cuda"))


def test_full_size_homography_properties_10M():
    """BASELINE config 4: 10 M states over 18 cameras.  Too large for the oracle as a whole, so: (1) 4000 sampled rows
    (incl. the first and the last 64) of state_to_im against the oracle, (2) the state -> image -> state round trip of ALL
    rows (homography.py:554-604's own property), (3) the refined two-pass transform returns the heights it was given when
    the template is the box itself"""
    from geom3d_b200 import ops
    from oracle import homography_oracle as ho
    P, H = synth.camera_matrices(18)
    Pd, Hd = torch.from_numpy(P).cuda(), torch.from_numpy(H).cuda()
    g = synth.gen(44)
    d = 10_000_000
    st, cam = synth.vehicle_states(d, g)
    std, camd = st.cuda(), cam.cuda()
    im = ops.state_to_im(std, Pd, camd, wrapper=True)
    assert im.shape == (d, 8, 2) and im.dtype == torch.float64 and bool(torch.isfinite(im).all())
    rows = torch.cat((torch.arange(64), torch.randint(64, d - 64, (3872,), generator=g), torch.arange(d - 64, d)))
    Pl = torch.from_numpy(P)[cam[rows].long()]
    assert_close_rel(im[rows.cuda()].cpu(), ho.wrapper_state_to_im(st[rows], Pl[:, 0], Pl[:, 1]), TOL, "sampled rows vs oracle")
    del im
    # one correspondence per camera (Homography): the round trip holds for every row
    im1 = ops.state_to_im(std, Pd, camd, wrapper=False)
    back = ops.im_to_state(im1, std[:, 4].contiguous(), Hd, camd, wrapper=False)
    assert back.shape == (d, 6) and torch.equal(back[:, 5], std[:, 5])
    err = ((back[:, :4] - std[:, :4]).abs() / std[:, :4].abs().clamp_min(1.0)).amax()
    assert float(err) < 1e-4, f"round trip of x, y, l, w: worst relative error {float(err):.3e}"
    assert float((back[:, 4] - std[:, 4]).abs().amax()) < 1e-4
    n2 = 1_000_000
    ref_state, heights = ops.im_to_state_refined(im1[:n2], std[:n2, 4].contiguous(), Hd, Pd, camd[:n2], wrapper=False,
                                                 return_heights=True)
    assert float((heights.float() - std[:n2, 4]).abs().amax()) < 1e-3
    assert float(((ref_state[:, :4] - std[:n2, :4]).abs() / std[:n2, :4].abs().clamp_min(1.0)).amax()) < 1e-3
    del im1
    # the two-correspondence wrapper picks a side of the roadway by y > 60 (homography.py:840-856, on different points
    # in the two directions), so its round trip holds away from that line
    imw = ops.state_to_im(std[:n2], Pd, camd[:n2], wrapper=True)
    backw = ops.im_to_state(imw, std[:n2, 4].contiguous(), Hd, camd[:n2], wrapper=True)
    far = (std[:n2, 1] - 60.0).abs() > 12.0
    errw = ((backw[:, :4] - std[:n2, :4]).abs() / std[:n2, :4].abs().clamp_min(1.0))[far].amax()
    assert int(far.sum()) > n2 // 2 and float(errw) < 1e-4, f"wrapper round trip: {float(errw):.3e}"
