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


def test_backward_keeps_forward_written_gradients_only_when_upstream_matches():
    """the device-side check: matching upstream gradients -> dcls / dreg untouched (poisoned here to see it); a different
    classification gradient -> dcls recomputed; different regression / direction gradients -> positive rows recomputed"""
    ops, _ = _mods()
    g = synth.gen(32)
    anc = synth.anchors(96, 96).cuda()
    ann = synth.gt_annotations_3d(2, 6, 96, 96, g, **synth.TINY).cuda()
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, g)
    cls, reg = cls.cuda(), reg.cuda()
    plain = ops.focal_loss_forward(cls, reg, anc, ann)
    want_c, want_r = ops.focal_loss_backward(plain, torch.tensor([1.0, 2.0, 3.0]).cuda())
    unit_c, unit_r = ops.focal_loss_backward(plain, torch.tensor([1.0, 1.0, 1.0]).cuda())
    pos = plain["assign"] >= 0
    assert int(pos.sum()) > 0 and int((want_r[~pos] != 0).sum()) == 0
    fwd = ops.focal_loss_forward(cls, reg, anc, ann, grad_expected=1.0)
    assert torch.equal(fwd["losses"], plain["losses"]) and torch.equal(fwd["assign"], plain["assign"])
    assert_close_rel(fwd["dcls"].cpu(), want_c.cpu(), 1e-6, "forward-written dcls")   # two kernels, same formulas
    assert torch.equal(fwd["dreg"], unit_r), "forward-written dreg rows (expected upstream 1, 1)"
    fwd["dcls"].fill_(7.0)
    fwd["dreg"].fill_(5.0)
    got_c, got_r = ops.focal_loss_backward(fwd, torch.tensor([1.0, 1.0, 1.0]).cuda())
    assert bool((got_c == 7.0).all()) and bool((got_r == 5.0).all()), "upstream matched: nothing may be rewritten"
    got_c, got_r = ops.focal_loss_backward(fwd, torch.tensor([1.0, 2.0, 3.0]).cuda())
    assert bool((got_c == 7.0).all()), "classification gradient matched: dcls must not be rewritten"
    assert torch.equal(got_r[pos], want_r[pos]) and bool((got_r[~pos] == 5.0).all())
    got_c, got_r = ops.focal_loss_backward(fwd, torch.tensor([2.0, 2.0, 3.0]).cuda())
    assert_close_rel(got_c.cpu(), (2 * want_c).cpu(), TOL, "recomputed dcls")
    assert torch.equal(got_r[pos], want_r[pos])
    # three expected values, one per loss
    fresh = ops.focal_loss_forward(cls, reg, anc, ann, grad_expected=1.0)
    fwd3 = ops.focal_loss_forward(cls, reg, anc, ann, grad_expected=(1.0, 2.0, 3.0))
    assert torch.equal(fwd3["dreg"], want_r) and torch.equal(fwd3["dcls"], fresh["dcls"])


def test_all_empty_batch_raises_and_optional_nan():
    _, li = _mods()
    anc = synth.anchors(64, 64).cuda()
    cls, reg = synth.head_outputs(2, anc.shape[1], 8, 12, synth.gen(3))
    ann = -torch.ones(2, 4, 27)
    with pytest.raises(RuntimeError):
        li.FocalLoss(check_empty=True)(cls.cuda(), reg.cuda(), anc, ann.cuda())   # the reference's immediate error
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
def test_gt_centric_assignment_equals_anchor_centric_and_oracle(seed):
    """anchors from Anchors.forward carry the pyramid description -> GT-centric assignment (window pairs + resolve, dreg
    zero-filled by bulk copies spread over the launches).  Codes, losses and gradients must equal the anchor-centric
    kernel's (same table without the tag) bit for bit, and the oracle's, incl. ties between equal GT boxes (lower index
    wins), padded rows, an empty image, boxes that leave the image, and the 2D variant"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
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
    f_gt = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_t, ann.cuda(), grad_expected=1.0)
    f_an = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_plain, ann.cuda(), grad_expected=1.0)
    assert f_gt["gt_centric"] and not f_an["gt_centric"]
    assert ops.STATS["gt_centric_calls"] == before["gt_centric_calls"] + 1
    assert torch.equal(f_gt["assign"], f_an["assign"]), "codes differ between the two assignment kernels"
    assert torch.equal(f_gt["per_image"], f_an["per_image"]) and torch.equal(f_gt["losses"], f_an["losses"])
    assert torch.equal(f_gt["dreg"], f_an["dreg"]) and torch.equal(f_gt["dcls"], f_an["dcls"])
    # forward-only calls take the same path and give the same numbers
    f_fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_t, ann.cuda())
    assert f_fwd["gt_centric"] and torch.equal(f_fwd["assign"], f_an["assign"]) and torch.equal(f_fwd["losses"], f_gt["losses"])
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


@pytest.mark.parametrize("shape", [(200, 328, 3, 17), (96, 96, 2, 60), (64, 64, 1, 250)])
def test_dreg_fill_and_overlapped_positives_are_invisible(shape):
    """GT-centric path: dreg is zero-filled by the streaming launch - bulk copies for chunks without keys, lane stores for
    chunks with keys except the positive rows - while the positives launch, started behind it as a programmatic dependent
    launch, writes the positive rows concurrently.  Whatever the buffer held before, with and without PDL, the result
    equals the anchor-centric path's (plain stores, separate launches) bit for bit."""
    ops, _ = _mods()
    g = synth.gen(7300)
    H, W, B, G = shape
    anc = _tagged_anchors(H, W)
    A = anc.shape[1]
    ann = synth.gt_annotations_3d(B, G, H, W, g, n_pad=1, **synth.TINY).cuda()
    cls, reg = synth.head_outputs(B, A, 8, 12, g)
    cls, reg = cls.cuda(), reg.cuda()
    want = ops.focal_loss_forward(cls, reg, anc.clone(), ann, grad_expected=1.0)      # anchor-centric
    assert not want["gt_centric"]
    try:
        for pdl in (1, 0, 1):
            ops.set_tuning("pdl", pdl)
            poison = torch.full_like(reg, float("nan"))
            del poison                                            # the caching allocator hands the block to dreg next
            got = ops.focal_loss_forward(cls, reg, anc, ann, grad_expected=1.0)
            assert got["gt_centric"]
            assert torch.equal(got["dreg"], want["dreg"]) and torch.equal(got["dcls"], want["dcls"])
            assert torch.equal(got["losses"], want["losses"]) and torch.equal(got["per_image"], want["per_image"])
    finally:
        ops.set_tuning("pdl", 1)


def test_gt_centric_assignment_falls_back(monkeypatch):
    """more than 256 annotation rows, a negative threshold below the key range, or a table modified in place after Anchors
    produced it -> anchor-centric kernel"""
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
    assert not ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc, small.cuda(), hyper=dict(neg_iou=0.2, pos_iou=0.3))["gt_centric"]
    monkeypatch.setenv("G3D_ASSIGN_GT_CENTRIC", "0")
    assert not ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc, small.cuda())["gt_centric"]
    monkeypatch.delenv("G3D_ASSIGN_GT_CENTRIC")
    moved = _tagged_anchors(H, W + 8)
    moved += 1.0                                                 # in-place edit bumps the tensor version
    assert ops.anchor_pyramid_of(moved) is None


# ---------------------------------------------------------------------------------------------------------------------
# the benchmarked configuration against the oracle (missing #1)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tagged", [True, False])
def test_cfg2_one_image_1080p_200gt_vs_oracle(tagged):
    """BASELINE.json configs[1] at its full per-image size - 1080p, A = 389 205, 200 GT boxes, C = 8 - for one image of
    the batch: assignment codes exact, losses / dcls / dreg within 1e-5 of the oracle's forward + autograd backward.
    tagged: the table of the drop-in Anchors module (GT-centric assignment, the path bench.py times); untagged: the
    anchor-centric kernel on the same values.  bench.py seeds its image 0 the same way (synth.gen(100))."""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(100)
    H, W = 1080, 1920
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    assert A == 389205
    ann = synth.gt_annotations_3d(1, 200, H, W, g)
    cls, reg = synth.head_outputs(1, A, 8, 12, g)
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)
    sum(l.sum() for l in ref[:3]).backward()
    anc_d = _tagged_anchors(H, W) if tagged else anc.cuda()
    assert torch.equal(anc_d.cpu(), anc)
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    before = dict(ops.STATS)
    out = li.FocalLoss()(c1, r1, anc_d, ann.cuda())
    assert ops.STATS["gt_centric_calls" if tagged else "anchor_centric_calls"] == \
        before["gt_centric_calls" if tagged else "anchor_centric_calls"] + 1
    sum(o.sum() for o in out).backward()
    assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref[:3]).detach(), TOL, "losses at cfg2 size")
    assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls at cfg2 size")
    assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg at cfg2 size", row_scale=True)
    assert torch.equal(c1.grad.cpu() == 0, c0.grad == 0) and torch.equal(r1.grad.cpu() == 0, r0.grad == 0)
    fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc_d, ann.cuda())
    assert torch.equal(fwd["assign"].cpu(), _codes_from_oracle(ref[-1], ann, True)), "assignment codes must be exact"
    npos = int(ref[-1][0][2].sum())
    assert int(fwd["per_image"][0, 3]) == npos and npos > 1000
    # un-floored worst element, for the record (see conftest.assert_close_rel for the judged metric)
    nz = r0.grad != 0
    worst = float(((r1.grad.cpu() - r0.grad).abs()[nz] / r0.grad.abs()[nz]).max())
    print(f"cfg2 image: {npos} positives, worst un-floored relative dreg error {worst:.2e}")


def test_zero_direction_vector_gives_nan_loss_and_gradient_like_the_reference():
    """no epsilon in reg_norm * targ_norm (3D losses.py:227): a regression direction vector that is exactly 0 - the
    zero-initialised head, model.py:257-258 - makes the vp loss NaN and the gradient of that vector NaN; everything else
    stays finite.  The kernels must reproduce the NaN pattern of the reference's autograd exactly."""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(61)
    anc = synth.anchors(96, 128)
    A = anc.shape[1]
    ann = synth.gt_annotations_3d(2, 6, 96, 128, g, **synth.TINY)
    cls, reg = synth.head_outputs(2, A, 8, 12, g)
    info = lo.focal_loss(cls, reg, anc, ann)[-1]
    victim = int(torch.nonzero(info[0][2])[0])                 # a positive anchor of image 0
    reg[0, victim, 4:6] = 0.0                                   # its width direction vector
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann)
    sum(l.sum() for l in ref[:3]).backward()
    assert torch.isnan(ref[2]).all() and torch.isfinite(ref[0]).all() and torch.isfinite(ref[1]).all()
    for table in (_tagged_anchors(96, 128), anc.cuda()):
        c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
        out = li.FocalLoss()(c1, r1, table, ann.cuda())
        sum(o.sum() for o in out).backward()
        assert torch.isnan(out[2]).all() and torch.isfinite(out[0]).all() and torch.isfinite(out[1]).all()
        assert_close_rel(torch.cat(out[:2]).detach().cpu(), torch.cat(ref[:2]).detach(), TOL, "cls / reg losses")
        assert torch.equal(torch.isnan(r1.grad.cpu()), torch.isnan(r0.grad)), "NaN pattern of dreg"
        assert bool(torch.isnan(r1.grad[0, victim, 4:6]).all()) and int(torch.isnan(r1.grad).sum()) == 2
        assert_close_rel(r1.grad.cpu(), r0.grad, TOL, "dreg (finite entries)", row_scale=True)
        assert_close_rel(c1.grad.cpu(), c0.grad, TOL, "dcls")


def test_eight_host_threads_reentrancy():
    """the C ABI promises re-entrancy (nn.DataParallel calls the loss from one thread per replica,
    train_detector_3D_angle.py:316-318): 8 host threads, each on its own stream with its own inputs, all at once"""
    import threading
    ops, li = _mods()
    from oracle import losses_oracle as lo
    H, W = 128, 160
    anc = _tagged_anchors(H, W)
    A = anc.shape[1]
    cases = []
    for t in range(8):
        g = synth.gen(900 + t)
        ann = synth.gt_annotations_3d(2, 5 + t, H, W, g, n_pad=t % 3, **synth.TINY)
        cls, reg = synth.head_outputs(2, A, 8, 12, g)
        cases.append((cls, reg, ann))
    results, errors = [None] * 8, []
    barrier = threading.Barrier(8)

    def work(t):
        try:
            cls, reg, ann = cases[t]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                c, r, a = cls.cuda(), reg.cuda(), ann.cuda()
                stream.synchronize()
                barrier.wait()
                acc = []
                for _ in range(20):
                    c1, r1 = c.clone().requires_grad_(True), r.clone().requires_grad_(True)
                    out = li.FocalLoss(check_empty=False)(c1, r1, anc if t % 2 == 0 else anc.clone(), a)
                    sum(o.sum() for o in out).backward()
                    acc.append((torch.cat(out).detach(), c1.grad, r1.grad))
                stream.synchronize()
            first = acc[0]
            for other in acc[1:]:
                assert all(torch.equal(x, y) for x, y in zip(first, other)), "run-to-run difference under concurrency"
            results[t] = tuple(x.cpu() for x in first)
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))
            try:
                barrier.abort()
            except Exception:  # noqa: BLE001
                pass

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    for t, (cls, reg, ann) in enumerate(cases):
        c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
        ref = lo.focal_loss(c0, r0, anc.cpu(), ann)
        sum(l.sum() for l in ref[:3]).backward()
        assert_close_rel(results[t][0], torch.cat(ref[:3]).detach(), TOL, f"thread {t} losses")
        assert_close_rel(results[t][1], c0.grad, TOL, f"thread {t} dcls")
        assert_close_rel(results[t][2], r0.grad, 3 * TOL, f"thread {t} dreg", row_scale=True)


@pytest.mark.parametrize("three_d", [True, False])
@pytest.mark.parametrize("hyper", [dict(alpha=0.4), dict(gamma=1.5), dict(pos_iou=0.6, neg_iou=0.45), dict(beta=0.25),
                                   dict(top_weighting=0.25, clamp_min=1e-3, clamp_max=0.995),
                                   dict(alpha=0.3, gamma=3.0, pos_iou=0.55, neg_iou=0.35, beta=0.05, top_weighting=1.0)])
def test_hyper_parameters_as_keyword_arguments(three_d, hyper):
    """the constants the reference hard-codes in forward (losses.py:28-30,56,121,124,343,346-348) are keyword arguments of
    the drop-in with the reference's values as defaults; non-default values against the oracle with the same values"""
    ops, li = _mods()
    from oracle import losses_oracle as lo
    g = synth.gen(4100)
    H, W = 136, 200
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    maker = synth.gt_annotations_3d if three_d else synth.gt_annotations_2d
    ann = maker(3, 14, H, W, g, n_pad=1, empty_images=(2,), **synth.TINY)
    cls, reg = synth.head_outputs(3, A, 8, 12 if three_d else 4, g)
    cls[0, :40] = torch.rand(40, 8, generator=g)
    n = 3 if three_d else 2
    c0, r0 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    ref = lo.focal_loss(c0, r0, anc, ann, hyper=hyper)
    sum(ref[i].sum() for i in range(n)).backward()
    for table in (_tagged_anchors(H, W), anc.cuda()):
        c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
        out = li.FocalLoss(check_empty=False, **hyper)(c1, r1, table, ann.cuda())
        sum(o.sum() for o in out).backward()
        assert_close_rel(torch.cat(out).detach().cpu(), torch.cat(ref[:n]).detach(), TOL, f"losses {hyper}")
        assert_close_rel(c1.grad.cpu(), c0.grad, TOL, f"dcls {hyper}")
        assert_close_rel(r1.grad.cpu(), r0.grad, 3 * TOL, f"dreg {hyper}", row_scale=True)
        fwd = ops.focal_loss_forward(cls.cuda(), reg.cuda(), table, ann.cuda(), hyper=hyper)
        codes = _codes_from_oracle(ref[-1], ann, three_d)
        assert torch.equal(fwd["assign"].cpu(), codes), "assignment codes must be exact for any thresholds"
    with pytest.raises(ValueError):
        ops.focal_loss_forward(cls.cuda(), reg.cuda(), anc.cuda(), ann.cuda(), hyper=dict(gama=2.0))


def test_combine_shard_stats_kernel_equals_host_restatement():
    """g3d_combine_shard_stats (the kernel the multi-GPU path runs after its all-gather) against dist.combine_on_host, the
    torch restatement the CPU gloo test drives"""
    ops, _ = _mods()
    from geom3d_b200 import dist as gdist
    g = synth.gen(77)
    gathered = torch.rand(4, 5, generator=g, dtype=torch.float64) * 10
    gathered[:, 3] = torch.tensor([8.0, 8.0, 7.0, 9.0])
    gathered[:, 4] = torch.tensor([8.0, 0.0, 5.0, 9.0])                  # one rank without any GT
    for rank in range(4):
        losses, scale = ops.combine_shard_stats(gathered.cuda(), rank)
        want_l, want_s = gdist.combine_on_host(gathered, rank)
        assert torch.equal(losses.cpu(), want_l) and torch.equal(scale.cpu(), want_s)


def test_lazy_empty_check_retain_graph_and_inplace_detection():
    """check_empty='lazy' (default): no synchronisation in forward, the all-empty batch is reported at the next call;
    a second backward (retain_graph) returns the same gradients in fresh tensors; without retain_graph autograd raises;
    an in-place edit of an input between forward and backward is caught by the version counters"""
    _, li = _mods()
    g = synth.gen(88)
    anc = synth.anchors(64, 64).cuda()
    A = anc.shape[1]
    cls, reg = synth.head_outputs(2, A, 8, 12, g)
    empty = -torch.ones(2, 4, 27)
    mod = li.FocalLoss()
    out = mod(cls.cuda(), reg.cuda(), anc, empty.cuda())
    assert torch.isnan(out[2]).all()
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="non-empty TensorList"):
        mod.check(wait=True)
    mod(cls.cuda(), reg.cuda(), anc, empty.cuda())
    torch.cuda.synchronize()
    ann = synth.gt_annotations_3d(2, 5, 64, 64, g, **synth.TINY).cuda()
    with pytest.raises(RuntimeError, match="non-empty TensorList"):
        mod(cls.cuda(), reg.cuda(), anc, ann)                       # the NEXT forward reports the earlier batch
    c1, r1 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    out = mod(c1, r1, anc, ann)
    total = sum(o.sum() for o in out)
    total.backward(retain_graph=True)
    gc, gr = c1.grad.clone(), r1.grad.clone()
    c1.grad = None
    r1.grad = None
    total.backward()
    assert_close_rel(c1.grad.cpu(), gc.cpu(), 1e-6, "second backward dcls") 
    assert torch.equal(r1.grad, gr)
    with pytest.raises(RuntimeError):
        total.backward()                                            # buffers were freed: autograd's own error
    c2, r2 = cls.cuda().requires_grad_(True), reg.cuda().requires_grad_(True)
    c2b = c2 * 1.0
    out = mod(c2b, r2, anc, ann)
    c2b.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        sum(o.sum() for o in out).backward()


@pytest.mark.parametrize("three_d", [True, False])
def test_persistent_gradient_buffers_equal_fresh_ones_step_after_step(three_d):
    """FocalLoss(persistent_grad=True): dreg is never filled, each step zeroes only the rows the previous step wrote
    (g3d_focal_loss_fwd_bwd, dreg_state = G3D_DREG_CLEAN).  Over steps whose annotations - and so the positive rows -
    change, with an image losing all its boxes, a forward-only call in between, a different upstream gradient and a change
    of batch size, losses and both gradients must equal the default module's bit for bit"""
    ops, li = _mods()
    H, W, B = 96, 160, 3
    anc = _tagged_anchors(H, W)
    A = anc.shape[1]
    keep, fresh = li.FocalLoss(persistent_grad=True), li.FocalLoss()
    g = synth.gen(91)
    for step in range(6):
        Bs = B if step < 4 else B + 1                                   # a new shape: buffers are re-made
        cls, reg = (t.cuda() for t in synth.head_outputs(Bs, A, 8, 12 if three_d else 4, g))
        ann = (synth.gt_annotations_3d if three_d else synth.gt_annotations_2d)(Bs, 9, H, W, g, **synth.TINY).cuda()
        if step == 2:
            ann[1, :, 20 if three_d else 4] = -1                        # image 1 has no box this time
        up = 1.0 if step != 3 else 0.37                                 # not the announced upstream gradient
        outs = []
        for mod in (keep, fresh):
            c1, r1 = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
            losses = mod(c1, r1, anc, ann)
            (sum(l.sum() for l in losses) * up).backward()
            outs.append((torch.cat([l.detach() for l in losses]), c1.grad.clone(), r1.grad.clone()))
        for a_, b_ in zip(*outs):
            assert torch.equal(a_, b_), step
        assert int((outs[0][2].abs().sum(-1) > 0).sum()) > 0
        if step == 1:                                                   # validation in between: no gradient, own workspace
            with torch.no_grad():
                keep(cls, reg, anc, ann)
    # the functional form with an explicit PersistentGrads; the returned gradients ARE its buffers
    pg = ops.PersistentGrads()
    fwd = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0, persistent=pg)
    assert fwd["dreg"].data_ptr() == pg.bufs[1].data_ptr() and torch.equal(fwd["dreg"], outs[1][2])
