"""GPU parity of device-side anchor generation: `Anchors` on a CUDA image runs `g3d_generate_anchors`;
the table must equal the reference's numpy table bit for bit (golden tables / sha256 from the unmodified reference, and
the oracle for non-default pyramids)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_anchor_kernel_equals_reference_tables():
    from geom3d_b200.anchors_impl import Anchors
    gd = np.load(os.path.join(GOLDEN, "anchors.npz"))
    shapes = [tuple(map(int, k[7:].split("x"))) for k in gd.files if k.startswith("sha256_")]
    mod = Anchors()
    for h, w in shapes:
        img = torch.zeros(1, 3, h, w, device="cuda")
        t = mod(img)
        assert t.is_cuda and t.dtype == torch.float32 and t.shape == (1, int(gd[f"count_{h}x{w}"]), 4)
        assert mod(img) is t                                                        # resident: generated once
        table = t[0].cpu().numpy()
        assert hashlib.sha256(table.tobytes()).digest() == gd[f"sha256_{h}x{w}"].tobytes(), (h, w)
        if f"anchors_{h}x{w}" in gd.files:
            assert np.array_equal(table, gd[f"anchors_{h}x{w}"])


def test_anchor_cache_is_bounded_and_never_hands_out_an_edited_table():
    from geom3d_b200 import ops
    from geom3d_b200.anchors_impl import Anchors
    mod = Anchors(max_cached=2)
    img = torch.zeros(1, 3, 64, 96, device="cuda")
    t = mod(img)
    pristine = t.clone()
    t[0, :, 0].clamp_(min=0)                          # what an in-place ClipBoxes does to its argument
    t2 = mod(img)
    assert t2 is not t and torch.equal(t2, pristine) and ops.anchor_pyramid_of(t2) is not None
    assert mod(img) is t2
    for hw in ((32, 32), (48, 48), (80, 80)):
        mod(torch.zeros(1, 3, *hw, device="cuda"))
    assert len(mod._cache) == 2 and (64, 96, "cuda:0") not in mod._cache
    assert torch.equal(mod(img), pristine)


def test_anchor_kernel_custom_pyramid_vs_oracle():
    from geom3d_b200 import ops
    from geom3d_b200.anchors_impl import Anchors
    from oracle import anchors_oracle as ao
    ratios, scales, levels = (0.3, 1.0, 1.7, 3.1), (1.0, 2 ** 0.5), (2, 4, 5, 8)
    mod = Anchors()
    mod.pyramid_levels = list(levels)
    mod.strides = [2 ** x for x in levels]
    mod.sizes = [2 ** (x + 2) for x in levels]
    mod.ratios, mod.scales = np.array(ratios), np.array(scales)
    for h, w in ((97, 211), (3, 5), (256, 256)):
        got = mod(torch.zeros(1, 1, h, w, device="cuda"))[0].cpu().numpy()
        assert np.array_equal(got, ao.anchors(h, w, levels, ratios, scales)), (h, w)
    # a level whose feature map is empty contributes nothing; zero anchors overall is legal
    z = ops.generate_anchors(np.zeros((2, 3, 4)), [8.0, 16.0], [0, 0], [4, 0], "cuda")
    assert z.shape == (0, 4)
    mixed = ops.generate_anchors(np.arange(24, dtype=np.float64).reshape(2, 3, 4), [8.0, 16.0], [0, 2], [4, 1], "cuda")
    assert mixed.shape == (6, 4) and mixed[0].tolist() == [12.0 + 8, 13.0 + 8, 14.0 + 8, 15.0 + 8]
    assert mixed[3].tolist() == [12.0 + 8, 13.0 + 24, 14.0 + 8, 15.0 + 24]


def test_anchor_kernel_argument_errors():
    from geom3d_b200 import ops
    from geom3d_b200._lib import Geom3dError
    with pytest.raises(Geom3dError):
        ops.generate_anchors(np.zeros((9, 12, 4)), [1.0] * 9, [1] * 9, [1] * 9, "cuda")         # more than 8 levels
    with pytest.raises(Geom3dError):
        ops.generate_anchors(np.zeros((8, 13, 4)), [1.0] * 8, [1] * 8, [1] * 8, "cuda")         # 104 level shapes
    with pytest.raises(Geom3dError):
        ops.generate_anchors(np.zeros((1, 9, 4)), [8.0], [1], [1], "cpu")                       # no CPU fallback
