"""CPU: host-side logic that needs no GPU - anchor layout, ladder rungs, matrix banks, camera-name handling, shard
ranges, and the world_size-2 gloo run of the loss reduction (dist.gather_shard_stats + combine) against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_anchor_counts_and_layout():
    from geom3d_b200.anchors_impl import Anchors, anchors_for_image
    assert anchors_for_image(540, 960).shape == (97965, 4)
    assert anchors_for_image(1080, 1920).shape == (389205, 4)
    assert anchors_for_image(112, 112).shape == (2394, 4)
    a = anchors_for_image(64, 64)
    # level 3 first: cell (0,0), 9 shapes centred at (4,4), ratio-major
    assert np.allclose((a[:9, 0] + a[:9, 2]) / 2, 4.0) and np.allclose((a[:9, 1] + a[:9, 3]) / 2, 4.0)
    w, h = a[:9, 2] - a[:9, 0], a[:9, 3] - a[:9, 1]
    assert np.allclose(h / w, np.repeat([0.5, 1, 2], 3), rtol=1e-6)
    assert np.allclose((a[9:18, 0] + a[9:18, 2]) / 2, 12.0)          # next cell along x
    from geom3d_b200._lib import Geom3dError
    with pytest.raises(Geom3dError):                                  # the module has no CPU path
        Anchors()(torch.zeros(2, 3, 64, 64))


def test_ladder_rungs():
    from geom3d_b200.ops import ladder_rungs
    r = ladder_rungs(1e-25)
    assert r.dtype == np.float32 and r[0] == np.float32(1e-25) and np.isinf(r[-1]) and np.all(np.diff(r[:-1]) > 0)
    t, k = 1e-25, 0
    while k < 40:
        assert r[k] == np.float32(t)
        t *= 10 ** .2
        k += 1
    assert len(ladder_rungs(1e-7)) < len(r) <= 512


def test_shard_range_partitions():
    from geom3d_b200.dist import shard_range
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_matrix_bank_and_camera_names():
    from geom3d_b200.homography_impl import Homography, Homography_Wrapper, _bank_for, _camera_arg
    P, H = synth.camera_matrices(4)
    hg1, hg2 = Homography(), Homography()
    for i, n in enumerate(synth.CAMERAS[:4]):
        hg1.add_correspondence_matrices(n, H[i, 0], P[i, 0])
        hg2.add_correspondence_matrices(n, H[i, 1], P[i, 1])
    assert hg1.default_correspondence == "p1c1"
    assert np.allclose(hg1.correspondence["p1c2"]["H_inv"] @ hg1.correspondence["p1c2"]["H"], np.eye(3), atol=1e-9)
    wr = Homography_Wrapper(hg1, hg2)
    bank = _bank_for(wr, hg1, hg2)
    assert bank.P_host.shape == (4, 2, 3, 4) and np.array_equal(bank.P_host[2, 1], P[2, 1])
    assert _bank_for(wr, hg1, hg2) is bank                              # cached until a matrix changes
    hg2.add_correspondence_matrices("p1c3", H[2, 0], P[2, 0])
    assert _bank_for(wr, hg1, hg2) is not bank
    cpu = torch.device("cpu")
    assert _camera_arg(bank, "p1c3", "p1c1", 5, cpu) == 2 and _camera_arg(bank, None, "p1c2", 5, cpu) == 1
    idx = _camera_arg(bank, ["p1c4", "p1c1", "p1c4"], None, 3, cpu)
    assert idx.dtype == torch.uint8 and idx.tolist() == [3, 0, 3]
    with pytest.raises(KeyError):
        _camera_arg(bank, "nope", None, 1, cpu)
    with pytest.raises(ValueError):
        _camera_arg(bank, ["p1c1"], None, 2, cpu)
    assert hg1.guess_heights(["semi", "sedan", "unicorn"]).tolist() == [12.0, 4.0, 5.0]
    # pickle round trip keeps the reference's attribute layout
    import pickle
    back = pickle.loads(pickle.dumps(hg1))
    assert set(back.correspondence) == set(hg1.correspondence) and back.class_heights["van"] == 6


def _shard_stats_of(per_image, gt_count):
    """float64[5] exactly as the kernel's finalize step writes shard_stats (focal_loss.cu: finalize_image)"""
    pi = per_image.double()
    ne = gt_count > 0
    return torch.stack((pi[:, 0].sum(), pi[:, 1].sum(), (pi[:, 2] * ne).sum(),
                        torch.tensor(float(per_image.shape[0]), dtype=torch.float64), ne.sum().double()))


def _gloo_worker(rank, world, port, per_image, gt_count, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the host functions the GPU path itself runs (dist._ShardedFocalLossFn.forward): the all-gather of the 5 shard
    # statistics, then the combination (on the GPU: g3d_combine_shard_stats; here its torch restatement, which the -m gpu
    # tests compare with the kernel)
    from geom3d_b200.dist import combine_on_host, gather_shard_stats, shard_range
    lo, hi = shard_range(per_image.shape[0], rank, world)
    gathered = gather_shard_stats(_shard_stats_of(per_image[lo:hi], gt_count[lo:hi]))
    losses, scale = combine_on_host(gathered, rank)
    out_q.put((rank, losses.tolist(), gathered.sum(0).tolist(), scale.tolist(), hi - lo, int((gt_count[lo:hi] > 0).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_reduction_gloo_world2():
    """two CPU ranks each reduce the per-image terms of their image shard; the combined means equal the oracle's
    whole-batch losses (the reference's semantics on one device), including the vp mean over non-empty images only"""
    from oracle import losses_oracle as lo
    g = synth.gen(21)
    anc = synth.anchors(64, 64)
    B = 5
    ann = synth.gt_annotations_3d(B, 5, 64, 64, g, n_pad=1, empty_images=(0, 3), **synth.TINY)
    cls, reg = synth.head_outputs(B, anc.shape[1], 8, 12, g)
    ref = lo.focal_loss(cls, reg, anc, ann)
    # per-image terms exactly as the kernel reports them: (cls_j, reg_j, vp_j, num_pos_j)
    rows = []
    for j in range(B):
        one = lo.focal_loss(cls[j:j + 1], reg[j:j + 1], anc, ann[j:j + 1]) if j not in (0, 3) else None
        if one is None:
            z = torch.zeros(anc.shape[1], dtype=torch.bool)
            c = lo.focal_classification_sum(cls[j], z, ~z, torch.zeros(anc.shape[1], dtype=torch.int64))
            rows.append(torch.stack((c, torch.tensor(0.0), torch.tensor(0.0), torch.tensor(0.0))))
        else:
            rows.append(torch.stack((one[0][0], one[1][0], one[2][0], one[3][0][2].sum().float())))
    per_image = torch.stack(rows)
    gt_count = torch.tensor([(ann[j][:, 20] != -1).sum() for j in range(B)], dtype=torch.int32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, per_image, gt_count, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.cat(ref[:3])
    for rank, losses, total, scale, b_local, ne_local in results:
        assert torch.allclose(torch.tensor(losses), expect, rtol=1e-6), (rank, losses, expect)
        assert total[3] == B and total[4] == 3
        # d(global mean)/d(local mean): B_l / B_g for cls and reg, NE_l / NE_g for vp
        assert scale == pytest.approx([b_local / B, b_local / B, ne_local / 3], rel=1e-6)
    assert results[0][1] == results[1][1], "every rank must hold bit-identical global losses"


def test_anchor_pyramid_tag_is_host_metadata_with_version_check(monkeypatch):
    """ops.tag_anchor_pyramid / anchor_pyramid_of: the description FocalLoss' GT-centric assignment needs travels as an
    attribute of the anchor table; a copy, an in-place edit or G3D_ASSIGN_GT_CENTRIC=0 drops it"""
    from geom3d_b200 import ops
    from geom3d_b200.anchors_impl import _level_shapes, _RATIOS, _SCALES, anchors_for_image
    H, W = 75, 133
    levels = (3, 4, 5, 6, 7)
    rows = [(H + 2 ** x - 1) // 2 ** x for x in levels]
    cols = [(W + 2 ** x - 1) // 2 ** x for x in levels]
    shapes = np.stack([_level_shapes(2 ** (x + 2), _RATIOS, _SCALES) for x in levels])
    table = torch.from_numpy(anchors_for_image(H, W)).unsqueeze(0)
    assert ops.anchor_pyramid_of(table) is None
    ops.tag_anchor_pyramid(table, rows, cols, [2.0 ** x for x in levels], shapes)
    desc = ops.anchor_pyramid_of(table)
    assert desc is not None and desc.dtype == np.float64 and desc[0] == 5 and desc[1] == 9
    assert desc.size == 2 + 3 * 5 + 2 * 5 * 9
    lv = desc[2:17].reshape(5, 3)
    assert int((lv[:, 0] * lv[:, 1]).sum()) * 9 == table.shape[1] and lv[0, 2] == 8.0
    wh = desc[17:].reshape(5, 9, 2)
    a = table[0].numpy()
    assert np.allclose(wh[0, :, 0], a[:9, 2] - a[:9, 0], rtol=1e-6) and np.allclose(wh[0, :, 1], a[:9, 3] - a[:9, 1], rtol=1e-6)
    assert ops.anchor_pyramid_of(table.clone()) is None                     # a copy is just a tensor
    monkeypatch.setenv("G3D_ASSIGN_GT_CENTRIC", "0")
    assert ops.anchor_pyramid_of(table) is None
    monkeypatch.delenv("G3D_ASSIGN_GT_CENTRIC")
    table += 0.5                                                             # in-place edit: version moves on
    assert ops.anchor_pyramid_of(table) is None
