"""GPU parity: IoU / assignment / fused focal+regression+direction losses (forward and backward) through the C ABI,
against the reference's golden vectors and against the oracle on seeded inputs.
Bar: indices and IoU values bit-exact; losses and gradients within 1e-5 relative (BASELINE.json north_star)."""
import pytest
import torch

import synth
from conftest import assert_close_rel

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _mods():
    from geom3d_b200 import losses_impl, ops
    return ops, losses_impl


def _codes_from_oracle(info, annotations, three_d):
    """expected assignment codes (-2 ignore, -1 negative, >=0 original annotation row) from the oracle's per-image info"""
    col = 20 if three_d else 4
    out = []
    for j, (iou_max, iou_arg, pos, neg, _) in enumerate(info):
        valid = torch.nonzero(annotations[j][:, col] != -1).flatten()
        code = torch.full(iou_max.shape, -2, dtype=torch.int32)
        code[neg] = -1
        if valid.numel():
            code[pos] = valid[iou_arg[pos]].to(torch.int32)
        out.append(code)
    return torch.stack(out)


@pytest.mark.parametrize("name", ["loss3d", "loss2d"])
def test_calc_iou_golden(golden, name):
    ops, _ = _mods()
    from oracle import losses_oracle as lo
    gd = golden(name)
    ann = gd["annotations"][0]
    rows = lo.valid_rows(ann[:, :21] if name == "loss3d" else ann, name == "loss3d")
    b = rows[:, 16:20] if name == "loss3d" else rows[:, :4]
    got = ops.calc_iou(gd["anchors"][0].cuda(), b.contiguous().cuda()).cpu()
    assert torch.equal(got, gd["iou_matrix0"])


@pytest.mark.parametrize("name", ["loss3d", "loss2d"])
def test_assignment_golden(golden, name):
    ops, _ = _mods()
    gd = golden(name)
    iou_max, iou_arg, code, npos = ops.assign(gd["anchors"].cuda(), gd["annotations"].cuda())
    assert torch.equal(iou_max.cpu(), gd["iou_max"]), "IoU_max must be bit-exact"
    assert torch.equal(iou_arg.cpu(), gd["iou_argmax"]), "IoU_argmax must be bit-exact"
    assert torch.equal(npos.cpu().long(), (gd["iou_max"] >= 0.5).sum(1))
    assert int((code.cpu() >= 0).sum()) == int((gd["iou_max"] >= 0.5).sum())


@pytest.mark.parametrize("name", ["loss3d", "loss2d"])
def test_focal_loss_golden_forward_backward(golden, name):
    _, li = _mods()
    gd = golden(name)
    cls = gd["classification"].cuda().requires_grad_(True)
    reg = gd["regression"].cuda().requires_grad_(True)
    out = li.FocalLoss()(cls, reg, gd["anchors"].cuda(), gd["annotations"].cuda())
    assert len(out) == (3 if name == "loss3d" else 2)
    assert all(o.shape == (1,) and o.is_cuda for o in out)
    got = torch.cat([o.detach() for o in out]).cpu()
    assert_close_rel(got, gd["losses"], TOL, "losses")
    w = gd["grad_weights"]
    sum(o.sum() * float(w[i]) for i, o in enumerate(out)).backward()
    assert_close_rel(cls.grad.cpu(), gd["dcls"], TOL, "dcls")
    assert_close_rel(reg.grad.cpu(), gd["dreg"], TOL, "dreg")
    # zero exactly where the reference is zero (ignored anchors, clamped probabilities, non-positive regressions)
    assert torch.equal(cls.grad.cpu() == 0, gd["dcls"] == 0)
    assert torch.equal(reg.grad.cpu() == 0, gd["dreg"] == 0)


@pytest.mark.parametrize("three_d", [True, False])
@pytest.mark.parametrize("shape", [(96, 128, 3, 7, 2), (200, 168, 5, 40, 0), (64, 64, 1, 300, 3)])
def test_focal_loss_vs_oracle_seeded(three_d, shape):
    """seeded inputs at sizes the oracle finishes in seconds; padded rows, an empty image, G > one staging chunk"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    H, W, B, G, n_pad = shape
    g = synth.gen(1000 + H + G)
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
    ann = maker(B, G, H, W, g, n_pad=n_pad, empty_images=(1,) if B > 2 else (), **synth.TINY)
    cls, reg = synth.head_outputs(B, A, 8, 12 if three_d else 4, g)
    ref = lo.focal_loss(cls, reg, anc, ann)
    ref_losses, info = torch.cat([l for l in ref[:-1]]), ref[-1]
    fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc.cuda(), ann.cuda())
    n = 3 if three_d else 2
    assert_close_rel(fwd["losses"][:n].cpu(), ref_losses, TOL, "losses")
    assert torch.equal(fwd["assign"].cpu(), _codes_from_oracle(info, ann, three_d)), "assignment codes must be exact"
    npos = torch.stack([i[2].sum() for i in info]).float()
    assert torch.equal(fwd["per_image"][:, 3].cpu(), npos)
    iou_max, iou_arg, code, _ = ops.assign(anc.cuda(), ann.cuda())
    assert torch.equal(iou_max.cpu(), torch.stack([i[0] for i in info]))
    assert torch.equal(iou_arg.cpu(), torch.stack([i[1] for i in info]))
    assert torch.equal(code, fwd["assign"])


def test_generic_class_count_and_gradients_vs_autograd_oracle():
    """C != 8 takes the generic kernel path; gradients against torch.autograd on the oracle"""
    _, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(77)
    anc = synth.anchors(96, 96)
    A = anc.shape[1]
    ann = synth.gt_annotations_3d(2, 9, 96, 96, g, n_pad=1, num_classes=5, **synth.TINY)
    cls, reg = synth.head_outputs(2, A, 5, 12, g)
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)[:-1]
    (ref[0].sum() + 2 * ref[1].sum() + 3 * ref[2].sum()).backward()
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    out = li.FocalLoss()(c1, r1, anc.cuda(), ann.cuda())
    (out[0].sum() + 2 * out[1].sum() + 3 * out[2].sum()).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref).detach(), TOL, "losses")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")
    assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg")


@pytest.mark.parametrize("three_d", [True, False])
@pytest.mark.parametrize("upstream", [(1.0, 1.0, 1.0), (1.0, 0.25, 3.0), (0.5, 1.0, 1.0)])
def test_forward_written_gradients_vs_autograd_oracle(three_d, upstream):
    """the usual step - (cls + reg + vp).backward() - takes the classification gradient written during the forward pass;
    other upstream gradients must give the autograd result as well (device-side check, then recompute)"""
    _, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(31)
    anc = synth.anchors(120, 160)
    A = anc.shape[1]
    maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
    ann = maker(3, 12, 120, 160, g, n_pad=2, empty_images=(1,), **synth.TINY)
    cls, reg = synth.head_outputs(3, A, 8, 12 if three_d else 4, g)
    cls[0, :50] = torch.rand(50, 8, generator=g)            # probabilities beyond the series range and the clamp
    cls[0, 50:60, 0] = 1e-5
    cls[0, 60:70, 1] = 1.0 - 1e-6
    n = 3 if three_d else 2
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)[:-1]
    sum(upstream[i] * ref[i].sum() for i in range(n)).backward()
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    out = li.FocalLoss()(c1, r1, anc.cuda(), ann.cuda())
    sum(upstream[i] * out[i].sum() for i in range(n)).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref).detach(), TOL, "losses")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")
    assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg")
    assert torch.equal(c1.grad.cpu() == 0, c0.grad == 0)
    assert torch.equal(r1.grad.cpu() == 0, r0.grad == 0)


def test_backward_keeps_forward_written_dcls_only_when_upstream_matches():
    """the device-side check: matching upstream gradient -> dcls untouched (poisoned here to see it); other -> recomputed"""
    ops, _ = _mods()
    g = synth.gen(32)
    anc = synth.anchors(96, 96).cuda()
    ann = synth.gt_annotations_3d(2, 6, 96, 96, g, **synth.TINY).cuda()
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, g)
    cls, reg = cls.cuda(), reg.cuda()
    plain = ops.focal_loss_forward(cls, reg, anc, ann)
    want_c, want_r = ops.focal_loss_backward(plain, torch.tensor([1.0, 2.0, 3.0]).cuda())
    fwd = ops.focal_loss_forward(cls, reg, anc, ann, grad_cls_expected=1.0)
    assert torch.equal(fwd["losses"], plain["losses"]) and torch.equal(fwd["assign"], plain["assign"])
    assert_close_rel(fwd["dcls"].cpu(), want_c.cpu(), 1e-6, "forward-written dcls")   # two kernels, same formulas
    assert int((fwd["dreg"] != 0).sum()) == 0
    fwd["dcls"].fill_(7.0)
    got_c, got_r = ops.focal_loss_backward(fwd, torch.tensor([1.0, 2.0, 3.0]).cuda())
    assert bool((got_c == 7.0).all()), "upstream gradient matched: dcls must not be rewritten"
    assert torch.equal(got_r, want_r)
    got_c, got_r = ops.focal_loss_backward(fwd, torch.tensor([2.0, 2.0, 3.0]).cuda())
    assert_close_rel(got_c.cpu(), (2 * want_c).cpu(), TOL, "recomputed dcls")
    assert torch.equal(got_r, want_r)


def test_all_empty_batch_raises_and_optional_nan():
    _, li = _mods()
    anc = synth.anchors(64, 64).cuda()
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, synth.gen(3))
    ann = -torch.ones(2, 4, 27)
    with pytest.raises(RuntimeError):
        li.FocalLoss()(cls.cuda(), reg.cuda(), anc, ann.cuda())
    out = li.FocalLoss(check_empty=False)(cls.cuda(), reg.cuda(), anc, ann.cuda())
    assert torch.isnan(out[2]).all() and float(out[1]) == 0.0
    # the classification loss of empty images is the un-normalised negative-only sum (losses.py:58-70)
    from oracle import losses_oracle as lo
    z = torch.zeros(anc.shape[1], dtype=torch.bool)
    expect = torch.stack([lo.focal_classification_sum(cls[j], z, ~z, torch.zeros(anc.shape[1], dtype=torch.int64)) for j in range(2)]).mean()
    assert_close_rel(out[0].cpu(), expect.reshape(1), TOL, "empty-image cls loss")


def test_inputs_are_not_mutated_and_cpu_tensors_raise():
    ops, li = _mods()
    from geom3d_b200 import Geom3dError
    g = synth.gen(5)
    anc = synth.anchors(64, 64)
    ann = synth.gt_annotations_3d(2, 4, 64, 64, g, **synth.TINY)
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, g)
    dev = [t.cuda() for t in (cls, reg, anc, ann)]
    li.FocalLoss()(*dev)
    for a, b in zip(dev, (cls, reg, anc, ann)):
        assert torch.equal(a.cpu(), b)
    with pytest.raises(Geom3dError):
        ops.focal_loss_forward(cls, reg, anc, ann)     # CPU tensors: no fallback


def test_full_size_properties_1080p():
    """BASELINE config 2 shape (one 1080p image, 200 GT): size-independent properties instead of the oracle."""
    ops, _ = _mods()
    g = synth.gen(9)
    anc = synth.anchors(1080, 1920).cuda()
    A = anc.shape[1]
    assert A == 389205
    ann = synth.gt_annotations_3d(2, 200, 1080, 1920, g).cuda()
    cls, reg = synth.head_outputs(2, A, 8, 12, g)
    cls, reg = cls.cuda(), reg.cuda()
    fwd = ops.focal_loss_forward(cls, reg, anc, ann)
    iou_max, iou_arg, code, npos = ops.assign(anc, ann)
    assert torch.equal(code, fwd["assign"])
    assert torch.equal(npos.float(), fwd["per_image"][:, 3])
    # (1) the culled search equals the brute-force IoU matrix on a random sample of anchors
    sel = torch.randint(0, A, (4096,), generator=g).cuda()
    gt_box, _, _ = ops.gt_prepare(ann)
    for j in range(2):
        m = ops.calc_iou(anc[0][sel].contiguous(), gt_box[j].contiguous())
        mx, am = m.max(dim=1)
        assert torch.equal(mx, iou_max[j][sel]) and torch.equal(am, iou_arg[j][sel])
    # (2) determinism: a second run is bit-identical (fixed-order reduction, no float atomics)
    again = ops.focal_loss_forward(cls, reg, anc, ann)
    assert torch.equal(again["losses"], fwd["losses"]) and torch.equal(again["per_image"], fwd["per_image"])
    # (3) image order invariance of the per-image terms
    flip = ops.focal_loss_forward(cls.flip(0).contiguous(), reg.flip(0).contiguous(), anc, ann.flip(0).contiguous())
    assert torch.equal(flip["per_image"].flip(0), fwd["per_image"])
    assert bool(torch.isfinite(fwd["losses"]).all()) and int(npos.min()) > 0


def test_sharded_loss_single_rank_equals_plain_loss_and_gradients():
    """dist.sharded_focal_loss without a process group (world 1): same losses and gradients as the plain module"""
    _, li = _mods()
    from geom3d_b200 import dist as gdist
    g = synth.gen(41)
    anc = synth.anchors(96, 128).cuda()
    ann = synth.gt_annotations_3d(3, 7, 96, 128, g, n_pad=1, empty_images=(2,), **synth.TINY).cuda()
    cls, reg = synth.head_outputs(3, anc.shape[1], 8, 12, g)
    c0, r0 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    want = li.focal_loss(c0, r0, anc, ann)[0]
    w = torch.tensor([1.0, 2.0, 0.5]).cuda()
    (want * w).sum().backward()
    got = gdist.sharded_focal_loss(c1, r1, anc, ann)
    (got * w).sum().backward()
    assert_close_rel(got.cpu(), want.detach().cpu(), 1e-6, "losses")
    assert_close_rel(c1.grad.cpu(), c0.grad.cpu(), 1e-6, "dcls")
    assert_close_rel(r1.grad.cpu(), r0.grad.cpu(), 1e-6, "dreg")


def test_baseline_config_1_vs_oracle():
    """BASELINE.json configs[0], the reference's own CPU-runnable case: B = 2 at 540x960 (A = 97 965), 20 GT boxes per
    image - losses, assignment codes and gradients against the oracle at full size"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(0)
    anc = synth.anchors(540, 960)
    A = anc.shape[1]
    assert A == 97965
    ann = synth.gt_annotations_3d(2, 20, 540, 960, g)
    cls, reg = synth.head_outputs(2, A, 8, 12, g)
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)
    sum(l.sum() for l in ref[:-1]).backward()
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    out = li.FocalLoss()(c1, r1, anc.cuda(), ann.cuda())
    sum(o.sum() for o in out).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref[:-1]).detach(), TOL, "losses")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")
    assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg")
    fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc.cuda(), ann.cuda())
    assert torch.equal(fwd["assign"].cpu(), _codes_from_oracle(ref[-1], ann, True)), "assignment codes must be exact"


@pytest.mark.parametrize("seed", list(range(12)))
def test_loss_random_small_configurations(seed):
    """seeded sweep over odd shapes: anchors not a multiple of the 32-row chunks / 256-anchor tiles, batch not a multiple
    of the 4-image groups, 0..N padded rows, an empty image, annotation widths 21 / 27 (3D) and 5 (2D), C == 8 and != 8;
    losses, exact codes and gradients (unit and non-unit upstream) against the oracle"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(5000 + seed)
    r = lambda lo_, hi_: int(torch.randint(lo_, hi_ + 1, (1,), generator=g))   # noqa: E731
    three_d = seed % 3 != 2
    H, W = 32 * r(2, 5) + 8 * r(0, 3), 32 * r(2, 6) + 8 * r(0, 3)
    B, G, n_pad, C = r(1, 9), r(1, 40), r(0, 3), (8 if seed % 4 else 5)
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    empty = (r(0, B - 1),) if (B > 1 and seed % 2) else ()
    maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
    ann = maker(B, G, H, W, g, n_pad=n_pad, empty_images=empty, num_classes=C, **synth.TINY)
    if three_d and seed % 5 == 0:
        ann = ann[..., :21].contiguous()                      # the narrowest legal 3D annotation row
    cls, reg = synth.head_outputs(B, A, C, 12 if three_d else 4, g)
    up = (1.0, 1.0, 1.0) if seed % 2 == 0 else (0.5, 2.0, 1.5)
    n = 3 if three_d else 2
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)
    sum(up[i] * ref[i].sum() for i in range(n)).backward()
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    out = li.FocalLoss(check_empty=False)(c1, r1, anc.cuda(), ann.cuda())
    sum(up[i] * out[i].sum() for i in range(n)).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref[:-1]).detach(), TOL, f"losses (B={B} A={A} G={G} C={C})")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")
    # dreg per anchor row (1e-5 of the row's largest component): the synthetic heads produce direction vectors as short
    # as 1e-3, where the cosine gradient is a difference of terms ~1/|r| and the FP32 autograd of the reference is itself
    # 1.2e-5 of an element away from the FP64 value of the same formula (the kernel's cross-product form is closer)
    assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg", row_scale=True)
    fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc.cuda(), ann.cuda())
    codes = _codes_from_oracle(ref[-1], ann, three_d)
    assert torch.equal(fwd["assign"].cpu(), codes), "assignment codes must be exact"
    # ... and element-wise against the same formula evaluated in FP64 (images whose FP64 assignment is the FP32 one) at
    # 3e-5: FP32 rounding of the regression targets alone puts the reference's own FP32 autograd up to 1.7e-5 from it
    c2, r2 = cls.double().requires_grad_(True), reg.double().requires_grad_(True)
    ref64 = lo.focal_loss(c2, r2, anc.double(), ann.double())
    sum(up[i] * ref64[i].sum() for i in range(n)).backward()
    same = (_codes_from_oracle(ref64[-1], ann, three_d) == codes).all(dim=1)
    assert same.any()
    assert_close_rel(r1.grad.cpu()[same], r2.grad[same], 3 * TOL, "dreg vs FP64 formula")


def _tagged_anchors(H, W):
    """the [1,A,4] table as the model gets it: Anchors()(image) on the device, carrying the pyramid description"""
    from geom3d_b200.anchors_impl import Anchors
    return Anchors()(torch.zeros(1, 3, H, W, device="cuda"))


@pytest.mark.parametrize("seed", list(range(10)))
def test_gt_centric_assignment_equals_anchor_centric_and_oracle(seed, monkeypatch):
    """anchors from Anchors.forward carry the pyramid description -> GT-centric assignment (fill + window pairs +
    resolve).  Codes, losses and gradients must equal the anchor-centric kernel's (same table without the tag) and the
    oracle's, incl. ties between equal GT boxes (lower index wins), padded rows, an empty image, boxes that leave the
    image, and the 2D variant"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    monkeypatch.setenv("G3D_ASSIGN_GT_CENTRIC", "1")          # also with gradient buffers (default: forward-only calls)
    g = synth.gen(7000 + seed)
    r = lambda lo_, hi_: int(torch.randint(lo_, hi_ + 1, (1,), generator=g))   # noqa: E731
    three_d = seed % 3 != 2
    H, W = (540, 960) if seed == 0 else (32 * r(2, 6) + 8 * r(0, 3), 32 * r(2, 8) + 8 * r(0, 3))
    B, G, n_pad = r(1, 6), r(1, 60), r(0, 3)
    anc_t = _tagged_anchors(H, W)
    assert ops.anchor_pyramid_of(anc_t) is not None
    anc_plain = anc_t.clone()                                   # same values, no tag -> anchor-centric kernel
    assert ops.anchor_pyramid_of(anc_plain) is None
    A = anc_t.shape[1]
    assert torch.equal(anc_t.cpu(), synth.anchors(H, W))
    empty = (r(0, B - 1),) if (B > 1 and seed % 2) else ()
    maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
    kw = {} if H >= 540 else synth.TINY
    ann = maker(B, G, H, W, g, n_pad=n_pad, empty_images=empty, **kw)
    col = 16 if three_d else 0
    if G >= 4:                                                  # exact duplicates (ties) and an out-of-image box
        ann[0, 2, :] = ann[0, 0, :]
        ann[0, 3, col:col + 4] = torch.tensor([-40.0, -30.0, 25.0, 20.0])
    cls, reg = synth.head_outputs(B, A, 8, 12 if three_d else 4, g)
    before = dict(ops.STATS)
    f_gt = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_t, ann.cuda(), grad_cls_expected=1.0)
    f_an = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_plain, ann.cuda(), grad_cls_expected=1.0)
    assert f_gt["gt_centric"] and not f_an["gt_centric"]
    assert ops.STATS["gt_centric_calls"] == before["gt_centric_calls"] + 1
    assert torch.equal(f_gt["assign"], f_an["assign"]), "codes differ between the two assignment kernels"
    assert torch.equal(f_gt["per_image"][:, 3], f_an["per_image"][:, 3])
    assert torch.equal(f_gt["dreg"], f_an["dreg"]) and torch.equal(f_gt["dcls"], f_an["dcls"])
    assert_close_rel(f_gt["losses"].cpu(), f_an["losses"].cpu(), 1e-6, "losses")
    # default policy: GT-centric for forward-only calls, anchor-centric when gradient buffers are written
    monkeypatch.delenv("G3D_ASSIGN_GT_CENTRIC")
    f_fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_t, ann.cuda())
    assert f_fwd["gt_centric"] and torch.equal(f_fwd["assign"], f_an["assign"]) and torch.equal(f_fwd["losses"], f_gt["losses"])
    assert not ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_t, ann.cuda(), grad_cls_expected=1.0)["gt_centric"]
    monkeypatch.setenv("G3D_ASSIGN_GT_CENTRIC", "1")
    ref = lo.focal_loss(cls, reg, anc_t.cpu(), ann)
    assert torch.equal(f_gt["assign"].cpu(), _codes_from_oracle(ref[-1], ann, three_d)), "codes differ from the oracle"
    # through the autograd module (the path a training step takes)
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    before = dict(ops.STATS)
    out = li.FocalLoss(check_empty=False)(c1, r1, anc_t, ann.cuda())
    assert ops.STATS["gt_centric_calls"] == before["gt_centric_calls"] + 1
    sum(o.sum() for o in out).backward()
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref2 = lo.focal_loss(c0, r0, anc_t.cpu(), ann)
    n = 3 if three_d else 2
    sum(ref2[i].sum() for i in range(n)).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref2[:n]).detach(), TOL, "losses vs oracle")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")
    # (f_gt == f_an bit for bit above; against the FP32 autograd oracle the cancellation-prone rows need the 3e-5 of
    # test_loss_random_small_configurations' FP64 comparison)
    assert_close_rel(r1.grad.cpu(), r0.grad, 3 * TOL, "dreg", row_scale=True)


def test_gt_centric_assignment_falls_back(monkeypatch):
    """more than 256 annotation rows, or a table modified in place after Anchors produced it -> anchor-centric kernel"""
    ops, _ = _mods()
    g = synth.gen(7100)
    H, W = 96, 128
    anc = _tagged_anchors(H, W)
    A = anc.shape[1]
    cls, reg = synth.head_outputs(1, A, 8, 12, g)
    big = synth.gt_annotations_3d(1, 300, H, W, g, **synth.TINY)
    assert not ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc, big.cuda())["gt_centric"]
    small = synth.gt_annotations_3d(1, 20, H, W, g, **synth.TINY)
    assert ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc, small.cuda())["gt_centric"]
    monkeypatch.setenv("G3D_ASSIGN_GT_CENTRIC", "0")
    assert not ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc, small.cuda())["gt_centric"]
    monkeypatch.delenv("G3D_ASSIGN_GT_CENTRIC")
    moved = _tagged_anchors(H, W + 8)
    moved += 1.0                                                 # in-place edit bumps the tensor version
    assert ops.anchor_pyramid_of(moved) is None
