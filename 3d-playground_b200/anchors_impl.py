"""Anchor generation with the layout the hot path consumes (anchors.py:6-129, identical in both retinanet copies).

Order: pyramid level 3..7 -> cell row-major (y outer, x inner) -> 9 shapes (ratio-major: ratios {0.5,1,2} x scales
{2^0, 2^(1/3), 2^(2/3)}).  Values are formed in float64 and cast to float32 exactly as the reference's numpy code does.

For an image on a CUDA device the table is written by one kernel (`g3d_generate_anchors`, csrc/anchors.cu, SURVEY
§8f-2) from the 45 base-shape doubles, once per (image shape, device), and stays resident; the reference rebuilds it in
numpy and copies 6.2 MB host->device on every forward at 1080p.  `anchors_for_image` is the host-side generator (numpy,
what the reference itself runs) used by the tests and the synthetic-input builders; the module itself takes CUDA images only.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

_PYRAMID_LEVELS = (3, 4, 5, 6, 7)
_RATIOS = np.array([0.5, 1, 2])
_SCALES = np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)])


def _level_shapes(size, ratios, scales):
    """[len(ratios)*len(scales), 4] float64 boxes centred on the origin, ratio-major."""
    side = size * np.tile(scales, len(ratios))              # square side before the ratio correction
    ratio = np.repeat(ratios, len(scales))
    w = np.sqrt((side * side) / ratio)
    h = w * ratio
    return np.stack((0.0 - w * 0.5, 0.0 - h * 0.5, w - w * 0.5, h - h * 0.5), axis=1)


def anchors_for_image(height, width, pyramid_levels=_PYRAMID_LEVELS, ratios=_RATIOS, scales=_SCALES, strides=None,
                      sizes=None):
    """float32 [A,4] anchors for an image of height x width."""
    out = []
    for k, lvl in enumerate(pyramid_levels):
        stride = 2 ** lvl if strides is None else strides[k]
        size = 2 ** (lvl + 2) if sizes is None else sizes[k]
        rows, cols = (height + 2 ** lvl - 1) // 2 ** lvl, (width + 2 ** lvl - 1) // 2 ** lvl     # anchors.py:25
        cx = (np.arange(cols) + 0.5) * stride
        cy = (np.arange(rows) + 0.5) * stride
        gx, gy = np.meshgrid(cx, cy)                            # [rows, cols], x fastest
        centres = np.stack((gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()), axis=1)     # [K,4]
        shapes = _level_shapes(size, ratios, scales)            # [9,4]
        out.append((centres[:, None, :] + shapes[None, :, :]).reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.float32)


class Anchors(nn.Module):
    """forward(image[B,C,H,W]) -> float32 [1,A,4] on the image's (CUDA) device.

    The table of an (image shape, device) is generated once by g3d_generate_anchors and kept (the reference rebuilds and
    re-uploads it on every forward).  The cached tensor is handed out as is - it carries the pyramid description that lets
    FocalLoss run its GT-centric assignment - but its version counter is remembered: a caller that modified it in place
    (an in-place ClipBoxes, say) gets a freshly generated table the next time, never the edited one.  At most
    `max_cached` shapes are kept (least recently used goes first).  CPU images raise: like every op of this package the
    module has no CPU fallback (`anchors_for_image` below is the host-side generator the tests and the synthetic-input
    builders use)."""

    def __init__(self, pyramid_levels=None, strides=None, sizes=None, ratios=None, scales=None, max_cached=8):
        super().__init__()
        self.pyramid_levels = list(_PYRAMID_LEVELS) if pyramid_levels is None else pyramid_levels
        self.strides = [2 ** x for x in self.pyramid_levels] if strides is None else strides
        self.sizes = [2 ** (x + 2) for x in self.pyramid_levels] if sizes is None else sizes
        self.ratios = _RATIOS if ratios is None else ratios
        self.scales = _SCALES if scales is None else scales
        self.max_cached = max(1, int(max_cached))
        self._cache = OrderedDict()       # (h, w, device) -> (table, version when generated)

    def _generate(self, h, w, device):
        from . import ops
        ratios, scales = np.asarray(self.ratios), np.asarray(self.scales)
        strides = [float(s) for s in self.strides]
        rows = [(h + 2 ** x - 1) // (2 ** x) for x in self.pyramid_levels]
        cols = [(w + 2 ** x - 1) // (2 ** x) for x in self.pyramid_levels]
        shapes = np.stack([_level_shapes(size, ratios, scales) for size in self.sizes])
        table = ops.generate_anchors(shapes, strides, rows, cols, device).unsqueeze(0)
        # host metadata: lets FocalLoss run its GT-centric assignment on this table (ops.anchor_pyramid_of)
        return ops.tag_anchor_pyramid(table, rows, cols, strides, shapes)

    def forward(self, image):
        if not image.is_cuda:
            from ._lib import Geom3dError
            raise Geom3dError("Anchors runs on CUDA images only (no CPU fallback): got an image on "
                              f"{image.device}; anchors_impl.anchors_for_image is the host-side generator")
        h, w = int(image.shape[2]), int(image.shape[3])
        key = (h, w, str(image.device))
        hit = self._cache.get(key)
        if hit is not None and hit[0]._version == hit[1]:
            self._cache.move_to_end(key)
            return hit[0]
        table = self._generate(h, w, image.device)
        self._cache[key] = (table, table._version)
        self._cache.move_to_end(key)
        while len(self._cache) > self.max_cached:
            self._cache.popitem(last=False)
        return table
