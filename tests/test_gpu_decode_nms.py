"""GPU parity: decode / clip (bit-exact for the add-mul 3D decode, 1e-6 for the exp of the 2D one), score ladder,
candidate compaction, NMS keep-lists (bit-exact vs torchvision as pinned in the golden vectors) and the three
post-processing return contracts of the models."""
import numpy as np
import pytest
import torch

import synth
from conftest import assert_close_rel

pytestmark = pytest.mark.gpu


def _mods():
    from geom3d_b200 import ops, postprocess
    return ops, postprocess


def test_decode_golden(golden):
    ops, pp = _mods()
    gd = golden("decode")
    anc = gd["anchors"].cuda()
    assert torch.equal(pp.BBoxTransform3D()(anc, gd["regression"].cuda()).cpu(), gd["decoded3d"]), "3D decode is bit-exact"
    d2 = pp.BBoxTransform2D()(anc, gd["deltas"].cuda())
    assert_close_rel(d2.cpu(), gd["decoded2d"], 1e-6, "2D decode (expf)")
    h, w = gd["image_hw"].tolist()
    img = torch.zeros(3, 3, h, w)
    ref_in = gd["decoded2d"].cuda()
    out = pp.ClipBoxes()(ref_in, img)
    assert out.data_ptr() == ref_in.data_ptr(), "ClipBoxes works in place"
    assert torch.equal(out.cpu(), gd["clipped2d"])
    fused = pp.BBoxTransform2D()(anc, gd["deltas"].cuda(), clip_wh=(w, h))
    assert_close_rel(fused.cpu(), gd["clipped2d"], 1e-6, "fused decode+clip")
    # non-contiguous view keeps the in-place contract
    wide = torch.zeros(3, anc.shape[1], 6, device="cuda")
    wide[..., :4] = gd["decoded2d"].cuda()
    pp.ClipBoxes()(wide[..., :4], img)
    assert torch.equal(wide[..., :4].cpu(), gd["clipped2d"])


def test_decode3d_ragged_sizes_vs_oracle():
    ops, _ = _mods()
    from oracle import decode_oracle
    g = synth.gen(3)
    for A, B in [(1, 1), (127, 2), (128, 1), (129, 3), (1000, 5)]:
        anc = torch.rand(1, A, 4, generator=g) * 100
        anc[..., 2:] += anc[..., :2] + 1
        reg = torch.randn(B, A, 12, generator=g)
        assert torch.equal(ops.decode3d(anc.cuda(), reg.cuda()).cpu(), decode_oracle.decode3d(anc, reg)), (A, B)
    assert ops.decode3d(torch.zeros(1, 0, 4).cuda(), torch.zeros(2, 0, 12).cuda()).shape == (2, 0, 20)


def test_nms_golden(golden):
    ops, pp = _mods()
    gd = golden("nms")
    b, s = gd["boxes"].cuda(), gd["scores"].cuda()
    for thr in (0.5, 0.3, 0.1, 0.8, 0.2):
        got = pp.nms(b, s, thr)
        assert got.dtype == torch.int64 and got.is_cuda
        assert torch.equal(got.cpu(), gd[f"keep_{thr}"]), f"keep-list at thr {thr}"
    assert torch.equal(pp.batched_nms(b, s, gd["idxs"].cuda(), 0.5).cpu(), gd["keep_batched_0.5"])
    assert torch.equal(pp.nms(gd["eq_boxes"].cuda(), gd["eq_scores"].cuda(), 0.5).cpu(), gd["eq_keep_0.5"])
    empty = pp.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.5)
    assert empty.shape == (0,) and empty.dtype == torch.int64
    assert pp.batched_nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0).long().cuda(), 0.5).numel() == 0


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 200, 255, 256, 1000, 2048, 2049, 3000, 4096, 4097, 5000, 9000, 20000])
def test_nms_sizes_vs_oracle(N):
    """block edges of the 64-wide greedy pass, the bit-matrix path (>= 256) with its rank sort / strip scan (<= 2048 one
    removed-set word per lane, <= 4096 two) and the shuffle scan above, the global sort (> 16384)"""
    ops, _ = _mods()
    from oracle import nms_oracle
    b, s = synth.clustered_boxes(N, synth.gen(N))
    s = (s * 50).round() / 50 if N <= 5000 else s        # ties for the smaller cases (rank sort: index order on ties)
    for thr in (0.5, 0.2):
        assert torch.equal(ops.nms(b.cuda(), s.cuda(), thr).cpu(), nms_oracle.nms(b, s, thr)), (N, thr)


def test_nms_segmented_matches_per_segment():
    ops, _ = _mods()
    from oracle import nms_oracle
    g = synth.gen(8)
    lens = [0, 5, 64, 700, 1, 0, 333, 2048]
    boxes, scores = synth.clustered_boxes(sum(lens), g)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    keep, cnt = ops.nms_segmented(boxes.cuda(), scores.cuda(), offs.cuda(), max(lens), 0.5)
    keep, cnt = keep.cpu(), cnt.cpu()
    for i, n in enumerate(lens):
        o = int(offs[i])
        exp = nms_oracle.nms(boxes[o:o + n], scores[o:o + n], 0.5)
        assert int(cnt[i]) == exp.numel()
        assert torch.equal(keep[o:o + exp.numel()], exp), f"segment {i}"


@pytest.mark.parametrize("cap", [100, 2000, 3000, 5000])
def test_nms_single_segment_with_device_side_length(cap):
    """one segment whose bounds live on the device (capacity = buffer size, length read by the kernels): what a captured
    launch sequence replays for a varying number of boxes; bit-matrix path (cap >= 256) and greedy path (cap < 256)"""
    ops, _ = _mods()
    from oracle import nms_oracle
    b, s = synth.clustered_boxes(cap, synth.gen(cap + 1))
    bd, sd = b.cuda(), s.cuda()
    seg = torch.zeros(2, dtype=torch.int32, device="cuda")
    for lo, hi in ((0, cap), (0, cap // 2 + 7), (0, 0), (0, 1), (0, min(cap, 257)), (cap // 3, cap - 5), (0, 64)):
        seg.copy_(torch.tensor([lo, hi], dtype=torch.int32))
        exp = nms_oracle.nms(b[lo:hi], s[lo:hi], 0.3)
        for relative in (True, False):
            keep, cnt = ops.nms_segmented(bd, sd, seg, cap, 0.3, 0, relative)
            assert int(cnt[0]) == exp.numel(), (lo, hi)
            assert torch.equal(keep[lo:lo + exp.numel()].cpu(), exp + (0 if relative else lo)), (lo, hi, relative)


def test_threshold_ladder_and_compaction_vs_oracle():
    ops, _ = _mods()
    from oracle import nms_oracle
    g = synth.gen(4)
    B, A, C = 2, 30000, 8
    cls = torch.rand(B, A, C, generator=g) * 0.3
    cls[1, :, 3] *= 0.01                      # a vector whose ladder stops at a low rung
    cls[0, :, 5] = 0.0                        # nothing above the first rung
    rung, count, thr = ops.threshold_ladder(cls.cuda(), B, C, A, A * C, 1e-25, 10000)
    idx, cnt = ops.filter_compact(cls.cuda(), B, C, A, A * C, thr, 10000)
    seg, cs, _, src = ops.gather_candidates(cls.cuda(), B, C, A, A * C, idx, cnt, 10000)
    seg, cs, src = seg.cpu(), cs.cpu(), src.cpu()
    for b in range(B):
        for c in range(C):
            mask, last = nms_oracle.ladder_threshold(cls[b, :, c], 1e-25)
            s = b * C + c
            assert float(thr[s]) == float(last), (b, c)
            assert int(count[s]) == int(mask.sum()) == int(cnt[s])
            o = int(seg[s])
            exp_idx = torch.nonzero(mask).flatten()
            assert torch.equal(src[o:o + exp_idx.numel()].long(), exp_idx)       # ascending = boolean-mask order
            assert torch.equal(cs[o:o + exp_idx.numel()], cls[b, :, c][mask])
    # flat vector, MULTI_FRAME start
    flat = torch.rand(50000, generator=g)
    rung, count, thr = ops.threshold_ladder(flat.cuda(), 1, 1, 50000, 50000, 1e-7, 10000)
    mask, last = nms_oracle.ladder_threshold(flat, 1e-7)
    assert float(thr[0]) == float(last) and int(count[0]) == int(mask.sum())


def test_rowmax_first_index_on_ties():
    ops, _ = _mods()
    x = torch.tensor([[0.1, 0.7, 0.7, 0.2], [0.5, 0.5, 0.5, 0.5], [0.0, 0.0, 0.0, 0.9]])
    s, a = ops.rowmax(x.cuda())
    es, ea = x.max(dim=1)
    assert torch.equal(s.cpu(), es) and torch.equal(a.cpu(), ea)


def test_postprocess_3d_golden(golden):
    _, pp = _mods()
    gd = golden("post3d")
    h, w = gd["image_hw"].tolist()
    anc = synth.anchors(h, w).cuda()
    post = pp.PostProcess3D()
    cls, reg = gd["classification"].cuda(), gd["regression"].cuda()
    s, c, b = post(cls[:1], reg[:1], anc)
    assert torch.equal(s.cpu(), gd["scores"]) and torch.equal(c.cpu(), gd["classes"]) and torch.equal(b.cpu(), gd["boxes"])
    assert c.dtype == torch.int64
    s, c, b, im = post(cls, reg, anc, MULTI_FRAME=True)
    assert torch.equal(s.cpu(), gd["mf_scores"]) and torch.equal(c.cpu(), gd["mf_classes"])
    assert torch.equal(b.cpu(), gd["mf_boxes"]) and torch.equal(im.cpu(), gd["mf_im"])
    bl, cl = post(cls, reg, anc, LOCALIZE=True)
    assert np.array_equal(synth.digest(bl), gd["loc_boxes_digest"].numpy()) and cl is cls


def test_postprocess_3d_dense_ladder_golden(golden):
    _, pp = _mods()
    gd = golden("post3d")
    h, w = gd["dense_hw"].tolist()
    cls, reg = synth.dense_detection_inputs(int(gd["dense_seed"][0]), h, w)
    s, c, b = pp.PostProcess3D()(cls.cuda(), reg.cuda(), synth.anchors(h, w).cuda())
    assert np.array_equal(np.bincount(c.cpu().numpy(), minlength=8), gd["dense_count"].numpy())
    assert torch.equal(s[:64].cpu(), gd["dense_head_scores"]) and torch.equal(b[:64].cpu(), gd["dense_head_boxes"])
    assert np.array_equal(synth.digest(s), gd["dense_scores_digest"].numpy())
    assert np.array_equal(synth.digest(b), gd["dense_boxes_digest"].numpy())


def test_postprocess_2d_golden(golden):
    _, pp = _mods()
    gd = golden("post2d")
    h, w = gd["image_hw"].tolist()
    s, c, b = pp.PostProcess2D()(gd["classification"].cuda(), gd["regression"].cuda(), synth.anchors(h, w).cuda(),
                                 torch.zeros(1, 3, h, w))
    # the 2D decode goes through expf: boxes agree to 1e-6, and on this fixture no IoU sits within that of the threshold
    assert torch.equal(s.cpu(), gd["scores"]) and torch.equal(c.cpu(), gd["classes"])
    assert_close_rel(b.cpu(), gd["boxes"], 1e-6, "2D boxes")


def test_batched_detection_equals_per_image():
    """cfg 3 shape in miniature: B images in one call == the same images one by one (per-image, per-class NMS)"""
    _, pp = _mods()
    g = synth.gen(12)
    H, W, B = 128, 160, 5
    anc = synth.anchors(H, W).cuda()
    A = anc.shape[1]
    cls = synth.detection_scores(B, A, 8, g, objects=10, per_object=9).cuda()
    reg = torch.randn(B, A, 4, generator=g).cuda() * 0.5
    boxes = pp.BBoxTransform2D()(anc, reg, clip_wh=(W, H))
    s, c, b, im = pp.detect_per_class(cls, boxes, score_threshold=0.05)
    for j in range(B):
        sj, cj, bj, _ = pp.detect_per_class(cls[j:j + 1], boxes[j:j + 1], score_threshold=0.05)
        m = im == j
        assert torch.equal(s[m], sj) and torch.equal(c[m], cj) and torch.equal(b[m], bj)


def test_full_size_decode_nms_properties():
    """BASELINE config 3 shapes (1080p, ~5k pre-NMS boxes per image): idempotence and order properties"""
    ops, pp = _mods()
    g = synth.gen(33)
    anc = synth.anchors(1080, 1920).cuda()
    A = anc.shape[1]
    B = 4
    cls = synth.detection_scores(B, A, 8, g).cuda()
    reg = torch.randn(B, A, 12, generator=g).cuda() * 0.1
    reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]).cuda() + torch.randn(B, A, 4, generator=g).cuda() * 0.05
    boxes = pp.BBoxTransform3D()(anc, reg)
    s, c, b, im = pp.detect_per_class(cls, boxes, box_col=16, score_threshold=0.05)
    assert s.numel() > 0 and bool((s > 0.05).all())
    # descending scores inside every (image, class) group, groups in image-major / class order
    key = im * 8 + c
    assert bool((key[1:] >= key[:-1]).all())
    same = key[1:] == key[:-1]
    assert bool((s[1:][same] <= s[:-1][same]).all())
    # idempotence: NMS of the kept boxes of a group keeps all of them, in the same order
    for grp in key.unique()[:6]:
        m = key == grp
        k2 = ops.nms(b[m][:, 16:20].contiguous(), s[m].contiguous(), 0.5)
        assert torch.equal(k2, torch.arange(int(m.sum()), device="cuda"))


@pytest.mark.parametrize("three_d", [True, False])
def test_fused_detection_tail_equals_decode_then_detect(three_d):
    """filter-before-decode must give, bit for bit, what decode -> detect_per_class gives - batch of
    images, fixed threshold and adaptive ladder, and an image without any detection"""
    _, pp = _mods()
    g = synth.gen(21)
    H, W, B = 128, 160, 4
    anc = synth.anchors(H, W).cuda()
    A = anc.shape[1]
    cls = synth.detection_scores(B, A, 8, g, objects=12, per_object=9).cuda()
    cls[2] = 0.01                                                     # nothing above the fixed threshold in image 2
    if three_d:
        reg = torch.randn(B, A, 12, generator=g).cuda() * 0.1
        reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]).cuda() + torch.randn(B, A, 4, generator=g).cuda() * 0.05
        boxes, col, extra = pp.BBoxTransform3D()(anc, reg), 16, {}
    else:
        reg = torch.randn(B, A, 4, generator=g).cuda() * 0.5
        t = pp.BBoxTransform2D()
        boxes, col = t(anc, reg, clip_wh=(W, H)), 0
        mean, std = t._host_params()
        extra = dict(mean=mean, std=std, clip_wh=(W, H))
    for kw in (dict(score_threshold=0.05), dict(ladder_start=1e-25, keep_max=300, cap=512)):
        want = pp.detect_per_class(cls, boxes, box_col=col, **kw)
        got = pp.detect_per_class_fused(cls, reg, anc, **kw, **extra)
        assert want[0].numel() > 0
        for w_, g_ in zip(want, got):
            assert w_.dtype == g_.dtype and torch.equal(w_, g_)
    none = pp.detect_per_class_fused(cls[2:3], reg[2:3], anc, score_threshold=0.05, **extra)
    assert none[0].numel() == 0 and none[2].shape == (0, 20 if three_d else 4)


@pytest.mark.parametrize("three_d", [True, False])
def test_postprocess_batch_is_flattened_like_the_reference(three_d):
    """B > 1 through the drop-in modules: the reference squeezes the batch dimension and boolean-masks [B,A] tensors, so a
    batch behaves like one image with B*A anchors (3D model.py:365-395, 2D retinanet/model.py:287-309) - every image's
    detections are returned, none is dropped"""
    _, pp = _mods()
    from oracle import decode_oracle, nms_oracle
    g = synth.gen(77)
    H, W, B = 96, 128, 3
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    cls = synth.detection_scores(B, A, 8, g, objects=8, per_object=9)
    if three_d:
        reg = torch.randn(B, A, 12, generator=g) * 0.1
        reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]) + torch.randn(B, A, 4, generator=g) * 0.05
        want = nms_oracle.detect_3d(cls, decode_oracle.decode3d(anc, reg))
        got = pp.PostProcess3D()(cls.cuda(), reg.cuda(), anc.cuda())
    else:
        reg = torch.randn(B, A, 4, generator=g) * 0.5
        boxes = decode_oracle.clip(decode_oracle.decode2d(anc, reg), H, W)
        want = nms_oracle.detect_2d(cls, boxes)
        got = pp.PostProcess2D()(cls.cuda(), reg.cuda(), anc.cuda(), torch.zeros(B, 3, H, W))
    assert want[0].numel() > 0
    one = (pp.PostProcess3D()(cls[:1].cuda(), reg[:1].cuda(), anc.cuda()) if three_d else
           pp.PostProcess2D()(cls[:1].cuda(), reg[:1].cuda(), anc.cuda(), torch.zeros(1, 3, H, W)))
    assert got[0].numel() > one[0].numel(), "the images after the first must not be dropped"
    assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
    if three_d:
        assert torch.equal(got[2].cpu(), want[2])
    else:
        assert_close_rel(got[2].cpu(), want[2], 1e-6, "2D boxes")


@pytest.mark.parametrize("three_d", [True, False])
def test_cfg3_one_image_1080p_vs_oracle(three_d):
    """BASELINE.json configs[2] at its full per-image size - 1080p, A = 389 205, ~5 000 boxes above 0.05, NMS 0.5 - for
    one image (bench.py's image 0): the fused tail (filter -> decode the candidates -> per-class NMS -> assemble) against
    the oracle's decode + torchvision-semantics NMS: identical scores / classes, boxes bit-exact (3D) or 1e-6 (2D expf)"""
    _, pp = _mods()
    from oracle import decode_oracle, nms_oracle
    g = synth.gen(7)
    H, W = 1080, 1920
    anc = synth.anchors(H, W)
    A = anc.shape[1]
    assert A == 389205
    cls = synth.detection_scores(1, A, 8, g)                       # 200 objects x 25 anchors scoring U(0.05, 1)
    n_cand = int((cls > 0.05).sum())
    assert 4000 < n_cand < 6000
    if three_d:
        reg = torch.randn(1, A, 12, generator=g) * 0.1
        reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]) + torch.randn(1, A, 4, generator=g) * 0.05
        dec = decode_oracle.decode3d(anc, reg)
        got = pp.detect_per_class_fused(cls.cuda(), reg.cuda(), anc.cuda(), score_threshold=0.05)
        col = 16
    else:
        reg = torch.randn(1, A, 4, generator=g) * 0.5
        dec = decode_oracle.clip(decode_oracle.decode2d(anc, reg), H, W)
        t = pp.BBoxTransform2D()
        mean, std = t._host_params()
        got = pp.detect_per_class_fused(cls.cuda(), reg.cuda(), anc.cuda(), score_threshold=0.05, mean=mean, std=std,
                                        clip_wh=(W, H))
        col = 0
    S, K, X = [], [], []
    for c in range(8):                                             # retinanet/model.py:287-309 with the oracle's nms
        m = cls[0, :, c] > 0.05
        if int(m.sum()) == 0:
            continue
        sc, bx = cls[0, m, c], dec[0][m]
        keep = nms_oracle.nms(bx[:, col:col + 4].contiguous(), sc, 0.5)
        S.append(sc[keep]); K.append(torch.full((keep.numel(),), c, dtype=torch.int64)); X.append(bx[keep])
    want = (torch.cat(S), torch.cat(K), torch.cat(X))
    assert 0 < want[0].numel() < n_cand, "NMS must suppress something on this workload"
    assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
    if three_d:
        assert torch.equal(got[2].cpu(), want[2])
    else:
        assert_close_rel(got[2].cpu(), want[2], 1e-6, "2D boxes")
    assert int(got[3].max()) == 0


def _tail_both_ways(cls, reg, anc, thr_score, iou, mean, std):
    """(short-path result, general-chain result, segments the short path left over) of one detection tail"""
    from geom3d_b200 import ops
    B, A, C = cls.shape
    thr = torch.full((B * C,), thr_score, dtype=torch.float32, device="cuda")
    cap = min(16384, A)
    res = []
    for short in (True, False):
        t = ops.detect_tail(cls, B, C, A, A * C, thr, cap, anc, reg, iou, mean, std, None, short=short)
        summ = t["summary"].cpu()
        out = ops.assemble_detections(t["keep"], t["keep_count"], t["seg_offsets"], t["cand_scores"], t["cand_src"], B, C, A,
                                      anc, reg, mean, std, None, out_offsets=t["out_offsets"], K=int(summ[0]))
        res.append((out, int(summ[2]), t))
    return res[0][0], res[1][0], res[0][1], res[0][2]


def _boxes_as_anchors(boxes):
    """[N,4] boxes -> (anchors [1,N,4], zero 2D regression [1,N,4], mean, std): the 2D decode of a zero delta is the box"""
    N = boxes.shape[0]
    return (boxes.reshape(1, N, 4).contiguous().cuda(), torch.zeros(1, N, 4, device="cuda"),
            np.zeros(4, dtype=np.float32), np.ones(4, dtype=np.float32))


@pytest.mark.parametrize("iou", [0.0, 0.1, 0.25, 0.3, 0.5, 0.75, 0.95])
def test_short_tail_equals_general_chain_on_adversarial_boxes(iou):
    """g3d_detect_tail_short finds the suppressing pairs through a grid over the box centres (a superset argument, see
    detect_tail.cu) - its detections must equal the general chain's (all-pairs greedy NMS) bit for bit on every box set:
    clusters, nested and extreme aspect ratios, exact duplicates, score ties, empty / inverted / NaN boxes, one spot, anisotropic extents"""
    g = synth.gen(int(iou * 100) + 3)
    sets = []
    b, _ = synth.clustered_boxes(3000, g)
    sets.append(b)
    c = torch.rand(2500, 2, generator=g) * 300                             # sizes over six decades, heavy nesting
    wh = 10 ** (torch.rand(2500, 2, generator=g) * 6 - 3)
    sets.append(torch.cat((c - wh / 2, c + wh / 2), dim=1))
    d = synth.clustered_boxes(400, g, objects=40)[0]                       # exact duplicates + near duplicates
    sets.append(torch.cat((d, d, d + 1e-4, d[:200])))
    e = synth.clustered_boxes(2000, g)[0]                                  # malformed rows sprinkled in
    e[::17, 2] = e[::17, 0]                                                # zero width
    e[5::23, [0, 2]] = e[5::23, [2, 0]]                                    # inverted
    e[7::29, 1] = float("nan")
    e[11::31, 3] = float("inf")                                            # the 2D decode turns it into inf - inf = NaN
    sets.append(e)
    sets.append(torch.tensor([[10., 10., 20., 20.]]).repeat(700, 1) + torch.rand(700, 4, generator=g) * 0.5)   # one spot
    f = synth.clustered_boxes(1500, g, extent=1e6, jitter=3e3)[0] * torch.tensor([1.0, 1e-3, 1.0, 1e-3])      # anisotropic
    sets.append(f)
    for k, boxes in enumerate(sets):
        N = boxes.shape[0]
        C = 4
        cls = torch.rand(1, N, C, generator=g)
        if k == 2:
            cls = (cls * 8).floor() / 8                                     # many equal scores
        hot = torch.randint(0, C, (N,), generator=g)
        cls[0, torch.arange(N), hot] += 1.0                                 # ~N/4 candidates per class above the cut ...
        cls = cls.cuda().contiguous()
        anc, reg, mean, std = _boxes_as_anchors(boxes.float())
        short, general, left, t = _tail_both_ways(cls, reg, anc, 1.0, iou, mean, std)
        assert int(t["count"].max()) <= 1024
        if left == 0:
            assert general[0].numel() > 0
            for a_, b_ in zip(short, general):
                assert a_.dtype == b_.dtype and a_.shape == b_.shape, (k, iou)
                if a_.is_floating_point():                                  # NaN box rows survive NMS: compare the bits
                    a_, b_ = a_.view(torch.int32), b_.view(torch.int32)
                assert torch.equal(a_, b_), (k, iou)
        else:
            assert k in (2, 4) or iou < 0.25, (k, iou, left)                # only the pair-heavy sets may overflow the pool


def test_short_tail_leaves_long_odd_and_pair_heavy_segments_to_the_general_chain():
    from geom3d_b200 import postprocess as pp
    g = synth.gen(77)
    # (a) a class with more than 1024 candidates, (b) boxes with sides below 1e-10, (c) 700 boxes on one spot (245 k
    # suppressing pairs)
    long_b = synth.clustered_boxes(3000, g)[0]
    tiny = synth.clustered_boxes(500, g)[0]
    tiny[::50] = torch.tensor([1e-11, 1e-11, 1.1e-11, 1.1e-11])        # sides of 1e-12
    spot = torch.tensor([[10., 10., 20., 20.]]).repeat(700, 1) + torch.rand(700, 4, generator=g) * 0.01
    for boxes, C in ((long_b, 2), (tiny, 1), (spot, 1)):
        N = boxes.shape[0]
        cls = (torch.rand(1, N, C, generator=g) + 1.0).cuda()
        anc, reg, mean, std = _boxes_as_anchors(boxes.float())
        short, general, left, t = _tail_both_ways(cls, reg, anc, 1.0, 0.5, mean, std)
        assert left >= 1 and int((t["keep_count"] < 0).sum()) == left
        # the public entry notices and repeats with the general chain - twice, to exercise the "stay general" hint
        for _ in range(2):
            got = pp.detect_per_class_fused(cls, reg, anc, score_threshold=1.0, mean=mean, std=std)
            for a_, b_ in zip(got, general):
                assert torch.equal(a_, b_)
    # ... and a shape goes back to the short path once its counts are short again
    key = (torch.cuda.current_device(), 1)
    pp._TAIL_GENERAL[key] = 1
    few = synth.clustered_boxes(300, g)[0]
    cls = (torch.rand(1, 300, 1, generator=g) + 1.0).cuda()
    anc, reg, mean, std = _boxes_as_anchors(few.float())
    pp.detect_per_class_fused(cls, reg, anc, score_threshold=1.0, mean=mean, std=std)
    assert pp._TAIL_GENERAL[key] == 0


def test_decode2d_is_bit_identical_to_the_reference_arithmetic_on_cuda():
    """retinanet/utils.py:102-126 evaluated by torch ON THE GPU (where the reference's model runs it): torch's CUDA `exp`
    is the same libdevice expf the kernel calls, every other step is one IEEE operation in the same order - so the kernel's
    boxes are bit-identical there.  (Against torch's CPU exp - the golden vectors - they differ by <= 1 ulp, which is why
    those tests use 1e-6 and why an NMS keep list could only differ if an IoU sat within that distance of the threshold.)"""
    _, pp = _mods()
    g = synth.gen(5)
    anc = synth.anchors(200, 328).cuda()
    A = anc.shape[1]
    reg = (torch.randn(3, A, 4, generator=g) * 0.7).cuda()
    t = pp.BBoxTransform2D()
    got = t(anc, reg)
    mean, std = t.mean.cuda(), t.std.cuda()
    boxes = anc
    widths = boxes[:, :, 2] - boxes[:, :, 0]
    heights = boxes[:, :, 3] - boxes[:, :, 1]
    ctr_x = boxes[:, :, 0] + 0.5 * widths
    ctr_y = boxes[:, :, 1] + 0.5 * heights
    dx = reg[:, :, 0] * std[0] + mean[0]
    dy = reg[:, :, 1] * std[1] + mean[1]
    dw = reg[:, :, 2] * std[2] + mean[2]
    dh = reg[:, :, 3] * std[3] + mean[3]
    pcx, pcy = ctr_x + dx * widths, ctr_y + dy * heights
    pw, ph = torch.exp(dw) * widths, torch.exp(dh) * heights
    want = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)
    assert torch.equal(got, want)


def test_detection_tail_from_four_host_threads():
    """the tail keeps its intermediates, its pinned summary block and its hints per (device, stream, host thread): four
    threads running different batches at the same time (DataParallel-style) each get their own, correct, detections"""
    import threading
    _, pp = _mods()
    H, W = 128, 160
    anc = synth.anchors(H, W).cuda()
    A = anc.shape[1]
    jobs = []
    for k in range(4):
        g = synth.gen(900 + k)
        cls = synth.detection_scores(2 + k % 2, A, 8, g, objects=10 + 3 * k, per_object=9).cuda()
        reg = (torch.randn(cls.shape[0], A, 12, generator=g) * 0.1).cuda()
        reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]).cuda() + torch.randn(cls.shape[0], A, 4, generator=g).cuda() * 0.05
        jobs.append((cls, reg, pp.detect_per_class_fused(cls, reg, anc, score_threshold=0.05)))
    torch.cuda.synchronize()
    errors = []

    def work(k):
        try:
            cls, reg, want = jobs[k]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(25):
                    got = pp.detect_per_class_fused(cls, reg, anc, score_threshold=0.05)
                    stream.synchronize()
                    for a_, b_ in zip(got, want):
                        assert torch.equal(a_, b_)
        except Exception as e:   # noqa: BLE001
            errors.append((k, repr(e)))
    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_tail_summary_by_copy_and_event_equals_mapped_polling(monkeypatch):
    """G3D_TAIL_MAPPED=0: the summary travels by asynchronous copy + event instead of mapped pinned memory - same detections"""
    _, pp = _mods()
    g = synth.gen(31)
    H, W = 128, 160
    anc = synth.anchors(H, W).cuda()
    A = anc.shape[1]
    cls = synth.detection_scores(3, A, 8, g, objects=12, per_object=9).cuda()
    reg = (torch.randn(3, A, 12, generator=g) * 0.1).cuda()
    reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]).cuda() + torch.randn(3, A, 4, generator=g).cuda() * 0.05
    want = pp.detect_per_class_fused(cls, reg, anc, score_threshold=0.05)
    monkeypatch.setenv("G3D_TAIL_MAPPED", "0")
    got = pp.detect_per_class_fused(cls, reg, anc, score_threshold=0.05)
    assert want[0].numel() > 0
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)
