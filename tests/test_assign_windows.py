"""CPU: the window bound the GT-centric assignment kernel relies on (3d-playground_b200/csrc/focal_loss.cu,
assign_pairs_kernel).  The kernel only evaluates, per GT row and per (level, shape) of the anchor pyramid, the cells
whose anchor centre lies in [g1 + i_min - a/2, g2 - i_min + a/2] with i_min from `inter >= t/(1+t) (Aa + Ag)`,
t = 0.385.  Restated here in float32 exactly as the kernel computes it (worst-case 3e-7 relative error injected for the
approximate reciprocal) and checked against the oracle's IoU matrix: every pair with IoU >= 0.4 - the only pairs that
can change an assignment code - must fall inside its window; the windows should also stay small."""
import math

import numpy as np
import pytest
import torch

import synth
from geom3d_b200.anchors_impl import _RATIOS, _SCALES, _level_shapes, anchors_for_image
from oracle import losses_oracle as lo

f = np.float32
LEVELS = (3, 4, 5, 6, 7)


def _window_candidates(H, W, box, recip_err):
    gx1, gy1, gx2, gy2 = [f(v) for v in box]
    gw, gh = f(gx2 - gx1), f(gy2 - gy1)
    if not (gw > 0 and gh > 0):
        return set()
    Ag = f(gw * gh)
    kq, eps = f(f(0.385) / f(1.385)), f(0.05)
    out, first = set(), 0
    for lvl in LEVELS:
        stride = 2 ** lvl
        inv = f(1.0 / stride)
        rows, cols = (H + stride - 1) // stride, (W + stride - 1) // stride
        shapes = _level_shapes(2 ** (lvl + 2), _RATIOS, _SCALES)
        for s in range(9):
            aw, ah = f(shapes[s, 2] - shapes[s, 0]), f(shapes[s, 3] - shapes[s, 1])
            imin = f(kq * f(f(aw * ah) + Ag))
            mw, mh = min(aw, gw), min(ah, gh)
            if f(mw * mh) < imin:
                continue
            iw_min = f(f(0.9999) * f(f(imin / mh) * f(1 + recip_err)))
            ih_min = f(f(0.9999) * f(f(imin / mw) * f(1 + recip_err)))
            lo_x, hi_x = f(f(f(gx1 + iw_min) - f(f(0.5) * aw)) - eps), f(f(f(gx2 - iw_min) + f(f(0.5) * aw)) + eps)
            lo_y, hi_y = f(f(f(gy1 + ih_min) - f(f(0.5) * ah)) - eps), f(f(f(gy2 - ih_min) + f(f(0.5) * ah)) + eps)
            c0, c1 = max(0, math.ceil(f(f(lo_x * inv) - f(0.5)))), min(cols - 1, math.floor(f(f(hi_x * inv) - f(0.5))))
            r0, r1 = max(0, math.ceil(f(f(lo_y * inv) - f(0.5)))), min(rows - 1, math.floor(f(f(hi_y * inv) - f(0.5))))
            for r in range(r0, r1 + 1):
                for c in range(c0, c1 + 1):
                    out.add(first + (r * cols + c) * 9 + s)
        first += rows * cols * 9
    return out


@pytest.mark.parametrize("H,W,G,tiny", [(540, 960, 40, False), (200, 168, 40, True), (96, 128, 30, True), (75, 133, 30, True)])
def test_windows_contain_every_pair_that_can_matter(H, W, G, tiny):
    g = synth.gen(H + W)
    anc = torch.from_numpy(anchors_for_image(H, W))
    ann = synth.gt_annotations_3d(1, G, H, W, g, **(synth.TINY if tiny else {}))
    boxes = ann[0][:, 16:20].clone()
    boxes[0] = torch.tensor([-20.0, -15.0, 30.0, 28.0])          # leaves the image
    boxes[1] = torch.tensor([10.0, 10.0, 10.0, 40.0])            # degenerate: can overlap nothing
    iou = lo.calc_iou(anc, boxes)
    need_total = cand_total = 0
    for gi in range(G):
        need = set(torch.nonzero(iou[:, gi] >= 0.4).flatten().tolist())
        for err in (-3e-7, 3e-7):
            cand = _window_candidates(H, W, boxes[gi], err)
            assert need <= cand, (H, W, gi, sorted(need - cand)[:5])
        need_total += len(need)
        cand_total += len(cand)
    assert need_total > 0 and cand_total < 2.5 * need_total + 50 * G, (need_total, cand_total)
