"""ORACLE (test infrastructure, not product code): CPU restatement of Anchors.forward of retinanet/anchors.py (the 2D
and the 3D copy of the file are identical), as explicit per-anchor loops in float64.

    feature-map sizes   anchors.py:23-25     ceil(H / 2^level), ceil(W / 2^level), levels 3..7
    base shapes         anchors.py:42-74     side = size * scale; w = sqrt(side^2 / ratio); h = w * ratio; box around 0
    shifts              anchors.py:109-129   ((col|row) + 0.5) * stride; cell row-major, shapes innermost
    dtype               anchors.py:29,33,38  float64 sums (np.append promotes), one cast to float32 at the end

Parity pin: tests/test_oracle_golden.py checks this against tests/golden/anchors.npz, the output of the UNMODIFIED
reference class on zero images (three shapes stored completely, three as sha256 of the float32 bytes).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
import math

import numpy as np

LEVELS = (3, 4, 5, 6, 7)
RATIOS = (0.5, 1.0, 2.0)
SCALES = (2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0))


def base_shapes(size, ratios=RATIOS, scales=SCALES):
    """[len(ratios)*len(scales)][4] python floats; ratio-major (anchors.py:57-72)."""
    rows = []
    for ratio in ratios:
        for scale in scales:
            side = size * scale
            w = math.sqrt((side * side) / ratio)
            h = w * ratio
            rows.append((0.0 - w * 0.5, 0.0 - h * 0.5, w - w * 0.5, h - h * 0.5))
    return rows


def anchors(height, width, levels=LEVELS, ratios=RATIOS, scales=SCALES):
    """float32 [A,4]"""
    out = []
    for lvl in levels:
        stride, size = 2 ** lvl, 2 ** (lvl + 2)
        n_rows, n_cols = (height + stride - 1) // stride, (width + stride - 1) // stride
        shapes = np.array(base_shapes(size, ratios, scales), dtype=np.float64)                 # [S,4]
        level = np.empty((n_rows, n_cols, shapes.shape[0], 4), dtype=np.float64)
        for r in range(n_rows):
            sy = (r + 0.5) * stride
            for c in range(n_cols):
                sx = (c + 0.5) * stride
                level[r, c] = shapes + np.array([sx, sy, sx, sy])
        out.append(level.reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.float32)
