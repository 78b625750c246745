"""CPU: the window bound the short-segment detection tail relies on (3d-playground_b200/csrc/detect_tail.cu,
tail_short_kernel, phase 3).  The kernel tests a pair of boxes only if the centre of one lies in the window
    |cx - cx'| <= win * w + eps,  |cy - cy'| <= win * h + eps,   win = 1.01 (1 - 0.99 t) / (0.99 t),  eps = 1e-5 max|coord|
around the centre of the other, (w, h) being the sides of the box that owns the window, and relies on the window of EITHER
box of a pair catching it (the pair is enumerated from the earlier grid cell, whichever box that is).  Restated here in
float32 as the kernel computes it and checked against the float32 IoU of torchvision's formula: every pair with
IoU > t must lie inside BOTH windows, for clustered, nested, extreme-aspect and tiny / huge (within the kernel's 1e-10 ..
1e15 range) boxes.  Also: the cell index is monotone in the coordinate, which is what turns the window into a cell range."""
import numpy as np
import pytest
import torch

import synth

f = np.float32


def _iou_f32(b):
    """pairwise IoU in float32, operation by operation as nms.cu / torchvision do it"""
    x1, y1, x2, y2 = (b[:, k].astype(f) for k in range(4))
    area = ((x2 - x1).astype(f) * (y2 - y1).astype(f)).astype(f)
    w = (np.minimum(x2[:, None], x2[None]) - np.maximum(x1[:, None], x1[None])).astype(f)
    h = (np.minimum(y2[:, None], y2[None]) - np.maximum(y1[:, None], y1[None])).astype(f)
    inter = (np.maximum(w, f(0)) * np.maximum(h, f(0))).astype(f)
    union = ((area[:, None] + area[None]).astype(f) - inter).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(f)
    return np.where((w > 0) & (h > 0), iou, f(0))


def _box_sets(g):
    sets = [synth.clustered_boxes(600, g)[0].numpy()]
    c = torch.rand(500, 2, generator=g) * 300
    wh = 10 ** (torch.rand(500, 2, generator=g) * 6 - 3)                      # six decades, heavy nesting
    sets.append(torch.cat((c - wh / 2, c + wh / 2), dim=1).numpy())
    d = synth.clustered_boxes(200, g, objects=20)[0]
    sets.append(torch.cat((d, d + 1e-4, d * 1.001)).numpy())                   # near duplicates
    sets.append((synth.clustered_boxes(400, g, extent=1e6, jitter=3e3)[0] * torch.tensor([1.0, 1e-3, 1.0, 1e-3])).numpy())
    sets.append((synth.clustered_boxes(300, g)[0] * 1e-7).numpy())            # sides ~1e-6 .. 1e-5
    sets.append((synth.clustered_boxes(300, g)[0] * 1e10 + 3e12).numpy())      # coordinates ~1e13
    return sets


@pytest.mark.parametrize("t", [0.25, 0.3, 0.5, 0.7, 0.9, 0.99])
def test_every_suppressing_pair_lies_in_both_windows(t):
    g = synth.gen(int(t * 1000))
    thr = f(t)
    t2 = f(f(0.99) * thr)
    win = f(f(1.01) * f(f(1) - t2) / t2)
    worst, pairs = 0.0, 0
    for b in _box_sets(g):
        b = b.astype(f)
        w, h = (b[:, 2] - b[:, 0]).astype(f), (b[:, 3] - b[:, 1]).astype(f)
        ok = (w >= f(1e-10)) & (h >= f(1e-10)) & (w <= f(1e15)) & (h <= f(1e15)) & (np.abs(b[:, 0]) <= f(1e15)) & (np.abs(b[:, 1]) <= f(1e15))
        b, w, h = b[ok], w[ok], h[ok]
        cx, cy = (b[:, 0] + f(0.5) * w).astype(f), (b[:, 1] + f(0.5) * h).astype(f)
        eps = f(f(1e-5) * np.abs(b).max())
        rx, ry = (win * w + eps).astype(f), (win * h + eps).astype(f)
        iou = _iou_f32(b)
        hit = np.triu(iou > thr, 1)
        i, j = np.nonzero(hit)
        pairs += len(i)
        for a_, b_ in ((i, j), (j, i)):                                        # the window of either box catches the other
            inx = (cx[b_] >= (cx[a_] - rx[a_]).astype(f)) & (cx[b_] <= (cx[a_] + rx[a_]).astype(f))
            iny = (cy[b_] >= (cy[a_] - ry[a_]).astype(f)) & (cy[b_] <= (cy[a_] + ry[a_]).astype(f))
            assert bool((inx & iny).all()), (t, int((~(inx & iny)).sum()))
            if len(a_):
                worst = max(worst, float((np.abs(cx[b_] - cx[a_]) / rx[a_]).max()), float((np.abs(cy[b_] - cy[a_]) / ry[a_]).max()))
    assert pairs > 100, pairs
    assert worst < 1.0        # and not by luck: the farthest partner sits well inside for t >= 0.5
    if t >= 0.5:
        assert worst < 0.75, worst


def test_cell_index_is_monotone_in_the_coordinate():
    """cell_coord(v) = clamp(int((v - lo) * scale), 0, 15): float subtraction, multiplication and truncation are all
    monotone, so every centre in [a, b] falls into a cell in [cell(a), cell(b)] - no margin needed for the cell range"""
    rng = np.random.RandomState(3)
    for lo, hi in ((0.0, 1920.0), (-3e5, 7e5), (1e-6, 3e-6), (3e12, 3.5e12)):
        lo, hi = f(lo), f(hi)
        scale = f(f(16) / f(hi - lo))
        v = np.sort(np.concatenate([rng.uniform(lo - (hi - lo), hi + (hi - lo), 20000), [lo, hi]]).astype(f))
        with np.errstate(invalid="ignore", over="ignore"):
            cell = np.clip(((v - lo).astype(f) * scale).astype(f).astype(np.int64), 0, 15)
        assert bool((np.diff(cell) >= 0).all())
