import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """dict of torch tensors from tests/golden/<name>.npz"""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over all elements, in float64"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() / b.abs().clamp(min=1e-30)).max()) if a.numel() else 0.0


WORST_RELATIVE = []     # (worst un-floored relative error, tolerance, what, row_scale, elements above tol un-floored, elements)


def assert_close_rel(a, b, tol, what="", row_scale=False):
    """|a-b| <= tol * max(|b|, scale), scale = mean |b| over the non-zero entries: element-wise relative error, except
    that entries far below the tensor's typical magnitude (sums with cancellation) are judged against that magnitude.
    row_scale=True judges every element against the largest |b| of its last-dimension row instead of its own |b|
    (norm-wise error per anchor row): the right measure for a gradient row whose components are differences of much
    larger terms -- there the FP32 reference itself sits ~1e-5 of an element away from the exact value."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.numel() == 0:
        return
    nz = b[(b != 0) & ~torch.isnan(b)]
    scale = float(nz.abs().mean()) if nz.numel() else 0.0
    mag = b.abs()
    if row_scale and b.dim() > 1:
        mag = torch.nan_to_num(mag, nan=0.0).amax(dim=-1, keepdim=True).expand_as(b)
    bound = tol * torch.maximum(mag, torch.full_like(b, scale))
    # un-floored figure for the report at the end of the run: the worst plain element-wise relative error |a-b| / |b|
    # (no magnitude floor, no row scale) - what north_star's "1e-5 relative" means literally
    finite = (b != 0) & ~torch.isnan(b) & ~torch.isnan(a)
    if bool(finite.any()):
        rel = ((a - b).abs()[finite] / b.abs()[finite])
        WORST_RELATIVE.append((float(rel.max()), tol, what or "?", bool(row_scale), int((rel > tol).sum()), int(finite.sum())))
    bad = (a - b).abs() > bound
    nan_mismatch = torch.isnan(a) != torch.isnan(b)
    bad = (bad & ~torch.isnan(b)) | nan_mismatch
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())} of {a.numel()} elements differ by more than {tol:g} relative; "
                                 f"worst |d|={float((a - b).abs()[~torch.isnan(a - b)].max()):.3e}")


def pytest_terminal_summary(terminalreporter):
    """tolerance report: for every assert_close_rel comparison of the run, the worst plain relative error (no magnitude
    floor, no row scaling) next to the tolerance the test applied with its floor - so the floors hide nothing"""
    if not WORST_RELATIVE:
        return
    tr = terminalreporter
    tr.write_sep("-", "assert_close_rel: worst un-floored element-wise relative errors")
    agg = {}
    for worst, tol, what, row_scale, above, n in WORST_RELATIVE:
        k = (what, tol, row_scale)
        w, a_, n_, c = agg.get(k, (0.0, 0, 0, 0))
        agg[k] = (max(w, worst), a_ + above, n_ + n, c + 1)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0] / kv[0][1])[:12]
    for (what, tol, row_scale), (worst, above, n, calls) in rows:
        tr.write_line(f"  {what[:60]:60s} tol {tol:g}{' (row-scaled)' if row_scale else ''}: worst {worst:.3e}; "
                      f"{above} of {n} elements above tol without the floor ({calls} comparisons)")
