"""Tensor-level wrappers over the libgeom3d C ABI (include/geom3d.h).

PyTorch is used for device memory and streams only: every function validates its tensors, allocates outputs with
torch on the inputs' device, and enqueues the CUDA kernels on torch's current stream of that device.  Nothing here
computes on the CPU and nothing falls back: a CPU tensor raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib
from ._lib import VARIANT_2D, VARIANT_3D, Geom3dError, check


# ------------------------------------------------------------------------------------------------------- helpers
def _need_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise Geom3dError("geom3d ops run on CUDA tensors only (no CPU fallback): got a tensor on "
                              f"{t.device}; move it with .cuda()")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise Geom3dError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _prep(t, dtype, align=16):
    """contiguous tensor of `dtype` whose data pointer is `align`-byte aligned (detached: raw kernels see no autograd)"""
    t = t.detach()
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _idx(dev):
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _workspace(nbytes, dev):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def sm_count(device=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return check(_lib.lib().g3d_sm_count(_idx(dev)), "g3d_sm_count")


# -------------------------------------------------------------------------------------------- a1 / a2: IoU, assignment
def calc_iou(a, b):
    """losses.py:5-22.  a[A,4], b[G,4] -> float32 [A,G], bit-identical to the eager reference."""
    dev = _need_cuda(a, b)
    a, b = _prep(a, torch.float32), _prep(b, torch.float32)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != 4 or b.shape[1] != 4:
        raise ValueError("calc_iou expects a[A,4] and b[G,4]")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_calc_iou(_p(a), a.shape[0], _p(b), b.shape[0], _p(out), _idx(dev), _stream(dev)), "g3d_calc_iou")
    return out


def _variant_of(annotations, regressions=None):
    w = annotations.shape[-1]
    if regressions is not None:
        r = regressions.shape[-1]
        if r == 12:
            return VARIANT_3D
        if r == 4:
            return VARIANT_2D
        raise ValueError(f"regression width {r}: expected 12 (3D directional) or 4 (2D)")
    return VARIANT_3D if w >= 21 else VARIANT_2D


def gt_prepare(annotations, variant=None):
    """Row filter (class != -1) + 2D assignment box per GT row.  Returns gt_box[B,G,4], gt_row[B,G], gt_count[B]."""
    dev = _need_cuda(annotations)
    ann = _prep(annotations, torch.float32)
    if ann.dim() != 3:
        raise ValueError("annotations must be [B,G,W]")
    variant = _variant_of(ann) if variant is None else variant
    B, G, W = ann.shape
    gt_box = torch.empty((B, G, 4), dtype=torch.float32, device=dev)
    gt_row = torch.empty((B, G), dtype=torch.int32, device=dev)
    gt_count = torch.empty((B,), dtype=torch.int32, device=dev)
    check(_lib.lib().g3d_gt_prepare(_p(ann), B, G, W, variant, _p(gt_box), _p(gt_row), _p(gt_count), _idx(dev),
                                    _stream(dev)), "g3d_gt_prepare")
    return gt_box, gt_row, gt_count


def assign(anchors, annotations, variant=None):
    """IoU max / argmax / assignment code per (image, anchor).  anchors[A,4] or [1,A,4]; annotations[B,G,W].

    Returns (iou_max f32[B,A], iou_argmax i64[B,A] (index into the filtered GT rows, as torch.max gives),
             assign i32[B,A] (-2 ignore, -1 negative, >=0 original annotation row), num_pos i32[B])."""
    dev = _need_cuda(anchors, annotations)
    anc = _prep(anchors, torch.float32).reshape(-1, 4)
    gt_box, gt_row, gt_count = gt_prepare(annotations, variant)
    B, G = gt_row.shape
    A = anc.shape[0]
    iou_max = torch.empty((B, A), dtype=torch.float32, device=dev)
    iou_arg = torch.empty((B, A), dtype=torch.int64, device=dev)
    code = torch.empty((B, A), dtype=torch.int32, device=dev)
    num_pos = torch.empty((B,), dtype=torch.int32, device=dev)
    check(_lib.lib().g3d_assign(_p(anc), A, _p(gt_box), _p(gt_row), _p(gt_count), B, G, _p(iou_max), _p(iou_arg),
                                _p(code), _p(num_pos), _idx(dev), _stream(dev)), "g3d_assign")
    return iou_max, iou_arg, code, num_pos


# ------------------------------------------------------------------------------------------------ a2-a6: fused loss
STATS = {"gt_centric_calls": 0, "anchor_centric_calls": 0}   # which assignment path focal_loss_forward took (for tests)
_STATS_LOCK = threading.Lock()

# the reference's constants (losses.py:28-30 alpha / gamma / top_weighting, :56 clamp, :121,:124 thresholds, :346-348 beta);
# order = include/geom3d.h hyper_host
HYPER_DEFAULTS = (("alpha", 0.25), ("gamma", 2.0), ("pos_iou", 0.5), ("neg_iou", 0.4), ("beta", 1.0 / 9.0),
                  ("top_weighting", 0.5), ("clamp_min", 1e-4), ("clamp_max", 1.0 - 1e-4))


def hyper_array(hyper):
    """dict of overrides (keys of HYPER_DEFAULTS) -> ctypes float[8], or None for the defaults"""
    if not hyper:
        return None
    unknown = set(hyper) - {k for k, _ in HYPER_DEFAULTS}
    if unknown:
        raise ValueError(f"unknown loss hyper-parameter(s) {sorted(unknown)}; known: {[k for k, _ in HYPER_DEFAULTS]}")
    return (ctypes.c_float * len(HYPER_DEFAULTS))(*[float(hyper.get(k, d)) for k, d in HYPER_DEFAULTS])


def set_tuning(name, value):
    """process-wide tuning knob of the loss path (g3d_set_tuning): pdl, force_anchor_centric"""
    check(_lib.lib().g3d_set_tuning(_lib.TUNE_KEYS[name], int(value)), "g3d_set_tuning")


def _tuning_from_env():
    """G3D_TUNE="pdl=0,force_anchor_centric=1" (benchmark sweeps); read once, at import"""
    spec = os.environ.get("G3D_TUNE", "")
    for item in filter(None, (x.strip() for x in spec.split(","))):
        k, v = item.split("=")
        set_tuning(k.strip(), int(v))


def anchor_pyramid_of(anchors):
    """the pyramid description `Anchors.forward` attaches to its table (None for any other anchor tensor, for a table
    modified in place since, or with G3D_ASSIGN_GT_CENTRIC=0): float64 array {L, S, L x (rows, cols, stride), L x S x
    (width, height)} - see g3d_focal_loss_fwd_bwd's pyramid_host"""
    tag = getattr(anchors, "_g3d_pyramid", None)
    if tag is None or os.environ.get("G3D_ASSIGN_GT_CENTRIC", "") == "0":
        return None
    desc, version = tag
    return desc if anchors._version == version else None


def tag_anchor_pyramid(table, rows, cols, strides, level_shapes):
    """attach the pyramid description to an anchor table produced by Anchors.forward (host metadata only)"""
    shp = np.asarray(level_shapes, dtype=np.float64)                       # [L,S,4]
    L, S = shp.shape[0], shp.shape[1]
    desc = np.concatenate(([float(L), float(S)],
                           np.stack((np.asarray(rows, dtype=np.float64), np.asarray(cols, dtype=np.float64),
                                     np.asarray(strides, dtype=np.float64)), axis=1).reshape(-1),
                           np.stack((shp[..., 2] - shp[..., 0], shp[..., 3] - shp[..., 1]), axis=2).reshape(-1)))
    table._g3d_pyramid = (np.ascontiguousarray(desc), table._version)
    return table


def _expected3(grad_expected):
    if grad_expected is None:
        return None
    if isinstance(grad_expected, (int, float)):
        grad_expected = (grad_expected,) * 3
    g = [float(x) for x in grad_expected]
    if len(g) != 3:
        raise ValueError("grad_expected: a float or three floats (cls, reg, vp)")
    return (ctypes.c_float * 3)(*g)


def focal_loss_forward(classifications, regressions, anchors, annotations, want_assign=True, grad_expected=None,
                       trace_events=None, want_shard_stats=False, hyper=None, grad_cls_expected=None, persistent=None):
    """FocalLoss forward.  Returns dict(losses f32[4], per_image f32[B,4], gt_count i32[B], assign i32[B,A] if
    want_assign, plus the prepared contiguous inputs and the workspace the backward needs).

    grad_expected (host float or 3 floats, e.g. 1.0): ALSO write, in the same launches, the complete gradients for those
    upstream gradients of (cls, reg, vp) - dict keys "dcls", "dreg", "grad_expected"; focal_loss_backward then confirms on
    the device that the upstream gradients are those and recomputes only what differs.  (grad_cls_expected: older name.)
    hyper: dict of loss hyper-parameter overrides (HYPER_DEFAULTS).  trace_events: up to 6 torch.cuda.Event(enable_timing=
    True) recorded before the first launch and after each of the five launches (see include/geom3d.h).
    persistent (PersistentGrads, with grad_expected): the gradients are written into ITS buffers, kept from step to step -
    dreg is then not filled at all, only the rows the previous step wrote are zeroed again (g3d_focal_loss_fwd_bwd,
    dreg_state = G3D_DREG_CLEAN): 43 % fewer bytes per step.  The returned gradients are those buffers: valid until the next
    forward with the same PersistentGrads."""
    if grad_expected is None and grad_cls_expected is not None:
        grad_expected = grad_cls_expected
    dev = _need_cuda(classifications, regressions, anchors, annotations)
    cls = _prep(classifications, torch.float32, 32)
    reg = _prep(regressions, torch.float32)
    # GT-centric assignment when the anchors are tagged as the regular pyramid (Anchors.forward); G3D_ASSIGN_GT_CENTRIC=0
    # disables it
    pyr = anchor_pyramid_of(anchors)
    if pyr is not None and (annotations.shape[1] > 256 or annotations.shape[1] < 1 or int(pyr[0]) > 8 or int(pyr[1]) > 16
                            or classifications.shape[-1] > 250 or (hyper and float(hyper.get("neg_iou", 0.4)) < 0.3)):
        pyr = None                                     # the library would fall back as well; keep the counter honest
    pyr_p = pyr.ctypes.data_as(ctypes.c_void_p) if pyr is not None else ctypes.c_void_p(0)
    with _STATS_LOCK:
        STATS["gt_centric_calls" if pyr is not None else "anchor_centric_calls"] += 1
    anc = _prep(anchors, torch.float32).reshape(-1, 4)
    ann = _prep(annotations, torch.float32)
    if cls.dim() != 3 or reg.dim() != 3 or ann.dim() != 3:
        raise ValueError("expected classifications[B,A,C], regressions[B,A,R], annotations[B,G,W]")
    B, A, C = cls.shape
    R = reg.shape[2]
    if reg.shape[0] != B or reg.shape[1] != A or anc.shape[0] != A or ann.shape[0] != B:
        raise ValueError(f"shape mismatch: cls {tuple(cls.shape)}, reg {tuple(reg.shape)}, anchors {tuple(anc.shape)}, "
                         f"annotations {tuple(ann.shape)}")
    variant = _variant_of(ann, reg)
    G, W = ann.shape[1], ann.shape[2]
    L = _lib.lib()
    wbytes = L.g3d_focal_workspace_bytes(B, A, G)
    ge = _expected3(grad_expected)
    grads = ge is not None
    keep = persistent.buffers(cls, reg, G, wbytes) if (persistent is not None and grads) else None
    ws = keep[2] if keep else _workspace(wbytes, dev)
    losses = torch.empty((4,), dtype=torch.float32, device=dev)
    per_image = torch.empty((B, 4), dtype=torch.float32, device=dev)
    code = torch.empty((B, A), dtype=torch.int32, device=dev) if want_assign else None
    gt_count = torch.empty((B,), dtype=torch.int32, device=dev)
    hy = hyper_array(hyper)
    out = dict(losses=losses, per_image=per_image, gt_count=gt_count, cls=cls, reg=reg, anchors=anc, ann=ann,
               variant=variant, workspace=ws, gt_centric=pyr is not None, hyper=hy, pyramid=pyr)
    if want_assign:
        out["assign"] = code
    stats = torch.empty((5,), dtype=torch.float64, device=dev) if want_shard_stats else None
    out["shard_stats"] = stats
    if keep:
        dcls, dreg = keep[0], keep[1]
    else:
        dcls = torch.empty_like(cls) if grads else None
        dreg = torch.empty_like(reg) if grads else None
    ev, n_ev = None, 0
    if trace_events is not None:
        for e in trace_events:          # torch creates the CUDA event lazily, on its first record
            if not e.cuda_event:
                e.record(torch.cuda.current_stream(dev))
        n_ev = len(trace_events)
        ev = (ctypes.c_void_p * n_ev)(*[e.cuda_event for e in trace_events])
    check(L.g3d_focal_loss_fwd_bwd(_p(cls), _p(reg), _p(anc), _p(ann), B, A, C, R, G, W, variant, hy, ge, _p(losses),
                                   _p(per_image), _p(code), _p(gt_count), _p(stats), _p(dcls), _p(dreg), _p(ws),
                                   ws.numel(), pyr_p, ev, n_ev, DREG_CLEAN if keep else DREG_UNDEFINED, _idx(dev),
                                   _stream(dev)), "g3d_focal_loss_fwd_bwd")
    if grads:
        out.update(dcls=dcls, dreg=dreg, grad_expected=ge)
    return out


DREG_UNDEFINED, DREG_CLEAN = 0, 1      # include/geom3d.h: G3D_DREG_*


class PersistentGrads:
    """Gradient buffers and loss workspace of one training loop, kept from step to step (FocalLoss(persistent_grad=True)).

    The regression gradient is zero except on the ~1 % positive rows, yet autograd wants it dense: writing its zeros is
    43 % of the bytes a loss step moves.  With a buffer that survives the step only the rows the previous step wrote need
    zeroing again.  The price is ownership: the tensors handed to autograd ARE these buffers - read them (the regression
    head's backward does) before the next forward, do not keep or modify them.  Buffers are re-made (dreg and workspace
    zeroed) whenever shape, device or annotation width change."""

    def __init__(self):
        self.key, self.bufs = None, None

    def buffers(self, cls, reg, G, wbytes):
        key = (tuple(cls.shape), tuple(reg.shape), int(G), cls.device, int(wbytes))
        if key != self.key:
            self.bufs = None                            # free the old ones first
            self.bufs = (torch.empty_like(cls), torch.zeros_like(reg),
                         torch.zeros(max(int(wbytes), 256), dtype=torch.uint8, device=cls.device))
            self.key = key
        return self.bufs


def combine_shard_stats(gathered, rank):
    """gathered f64[world,5] (every rank's shard_stats) -> (losses f32[3] global means, scale f32[3] for the backward)."""
    dev = _need_cuda(gathered)
    gathered = _prep(gathered, torch.float64)
    world = gathered.numel() // 5
    losses = torch.empty((3,), dtype=torch.float32, device=dev)
    scale = torch.empty((3,), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_combine_shard_stats(_p(gathered), world, int(rank), _p(losses), _p(scale), _idx(dev), _stream(dev)),
          "g3d_combine_shard_stats")
    return losses, scale


def exchange_shard_stats(stats, peer_ptrs_dev, world, rank):
    """stats f64[5] of this rank -> (losses f32[3] global means, scale f32[3]) through the peers' exchange buffers
    (g3d_exchange_shard_stats: stores over NVLink + epoch flags, no collective call).  peer_ptrs_dev: int, device address of
    the array of the `world` buffer pointers."""
    dev = _need_cuda(stats)
    stats = _prep(stats, torch.float64)
    losses = torch.empty((3,), dtype=torch.float32, device=dev)
    scale = torch.empty((3,), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_exchange_shard_stats(_p(stats), ctypes.c_void_p(int(peer_ptrs_dev)), int(world), int(rank), _p(losses),
                                              _p(scale), _idx(dev), _stream(dev)), "g3d_exchange_shard_stats")
    return losses, scale


def focal_loss_backward(fwd, grad_out, grad_scale=None, take=False):
    """Backward of focal_loss_forward.  grad_out f32[3] (device).  Returns (dcls[B,A,C], dreg[B,A,R]).

    If the forward already wrote the gradients for `grad_expected`, the kernels compare grad_out (* grad_scale) with it on
    the device: what agrees stays, what differs is recomputed (no host synchronisation either way).  take=True hands the
    forward's buffers over to the caller (they leave `fwd`: autograd can then adopt them as .grad without a copy, and a
    second backward through the same forward computes into fresh buffers)."""
    cls, reg, anc, ann = fwd["cls"], fwd["reg"], fwd["anchors"], fwd["ann"]
    dev = cls.device
    B, A, C = cls.shape
    R = reg.shape[2]
    G, W = ann.shape[1], ann.shape[2]
    g = _prep(grad_out, torch.float32)
    if g.numel() != 3:
        raise ValueError("grad_out must have 3 elements (cls, reg, vp)")
    gs = _prep(grad_scale, torch.float32) if grad_scale is not None else None
    if fwd.get("dcls") is not None:
        dcls, dreg, have, ge = fwd["dcls"], fwd["dreg"], 1, fwd["grad_expected"]
        if take:
            del fwd["dcls"], fwd["dreg"]
    else:
        dcls, dreg, have, ge = torch.empty_like(cls), torch.empty_like(reg), 0, None
    ws, pyr = fwd["workspace"], fwd.get("pyramid")
    pyr_p = pyr.ctypes.data_as(ctypes.c_void_p) if pyr is not None else ctypes.c_void_p(0)
    check(_lib.lib().g3d_focal_loss_bwd(_p(cls), _p(reg), _p(anc), _p(ann), B, A, C, R, G, W, fwd["variant"], fwd.get("hyper"),
                                        _p(g), _p(gs), have, ge, _p(ws), ws.numel(), pyr_p, _p(dcls), _p(dreg), _idx(dev),
                                        _stream(dev)), "g3d_focal_loss_bwd")
    return dcls, dreg


# ------------------------------------------------------------------------------------------------ a7-a9: decode / clip
def decode3d(anchors, regression):
    """3D BBoxTransform: anchors[1,A,4] or [A,4], regression[B,A,12] -> [B,A,20] (utils.py:102-149)."""
    dev = _need_cuda(anchors, regression)
    anc = _prep(anchors, torch.float32).reshape(-1, 4)
    reg = _prep(regression, torch.float32)
    if reg.dim() != 3 or reg.shape[2] != 12 or reg.shape[1] != anc.shape[0]:
        raise ValueError(f"expected regression[B,A,12] with A == {anc.shape[0]} anchors, got {tuple(reg.shape)}")
    B, A, _ = reg.shape
    out = torch.empty((B, A, 20), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_decode3d(_p(anc), _p(reg), B, A, _p(out), _idx(dev), _stream(dev)), "g3d_decode3d")
    return out


def decode2d(anchors, deltas, mean, std, clip_wh=None):
    """2D BBoxTransform (+ optional fused ClipBoxes): anchors[Ba,A,4], deltas[B,A,4] -> [B,A,4] (retinanet/utils.py:102-126)."""
    dev = _need_cuda(anchors, deltas)
    anc = _prep(anchors, torch.float32)
    dl = _prep(deltas, torch.float32)
    if anc.dim() == 2:
        anc = anc.unsqueeze(0)
    if dl.dim() != 3 or dl.shape[2] != 4 or anc.shape[1] != dl.shape[1] or anc.shape[2] != 4:
        raise ValueError(f"expected anchors[Ba,A,4] and deltas[B,A,4], got {tuple(anc.shape)} and {tuple(dl.shape)}")
    B, A, _ = dl.shape
    mean_h = (ctypes.c_float * 4)(*[float(x) for x in mean])
    std_h = (ctypes.c_float * 4)(*[float(x) for x in std])
    out = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    clip = 0 if clip_wh is None else 1
    cw, ch = (0.0, 0.0) if clip_wh is None else (float(clip_wh[0]), float(clip_wh[1]))
    check(_lib.lib().g3d_decode2d(_p(anc), anc.shape[0], _p(dl), B, A, mean_h, std_h, clip, cw, ch, _p(out), _idx(dev),
                                  _stream(dev)), "g3d_decode2d")
    return out


def clip_boxes_(boxes, width, height):
    """ClipBoxes in place on a contiguous float32 tensor [..., K>=4] (retinanet/utils.py:134-144)."""
    dev = _need_cuda(boxes)
    if boxes.dtype != torch.float32 or not boxes.is_contiguous():
        raise Geom3dError("clip_boxes_ works in place: needs a contiguous float32 tensor")
    K = boxes.shape[-1]
    N = boxes.numel() // K if K else 0
    check(_lib.lib().g3d_clip_boxes(_p(boxes), N, K, float(width), float(height), _idx(dev), _stream(dev)),
          "g3d_clip_boxes")
    return boxes


# ------------------------------------------------------------------------------------------------ a10: score filter
def rowmax(classification):
    """scores, classes = torch.max(classification, dim=-1) for [..., C] (3D model.py:320)."""
    dev = _need_cuda(classification)
    cls = _prep(classification, torch.float32)
    C = cls.shape[-1]
    rows = cls.numel() // C
    smax = torch.empty(cls.shape[:-1], dtype=torch.float32, device=dev)
    amax = torch.empty(cls.shape[:-1], dtype=torch.int64, device=dev)
    check(_lib.lib().g3d_rowmax(_p(cls), rows, C, _p(smax), _p(amax), _idx(dev), _stream(dev)), "g3d_rowmax")
    return smax, amax


_RUNG_CACHE = {}


def ladder_rungs(start, factor=10 ** .2):
    """float32-cast rungs of the reference's `threshold *= 10**.2` loop (3D model.py:368-374), up to +inf."""
    key = (float(start), float(factor))
    if key not in _RUNG_CACHE:
        t, out = float(start), []
        while True:
            with np.errstate(over="ignore"):          # the last rung overflows to +inf on purpose
                f = np.float32(t)
            out.append(f)
            if np.isinf(f) or len(out) >= 512:
                break
            t *= factor
        _RUNG_CACHE[key] = np.ascontiguousarray(np.array(out, dtype=np.float32))
    return _RUNG_CACHE[key]


def threshold_ladder(scores, outer, inner, N, outer_pitch, start, keep_max=10000):
    """Adaptive threshold per score vector.  Returns (rung i32[S], count i32[S], thr f32[S]) on the device."""
    dev = _need_cuda(scores)
    if scores.dtype != torch.float32 or not scores.is_contiguous():
        raise Geom3dError("threshold_ladder needs a contiguous float32 score tensor")
    rungs = ladder_rungs(start)
    S = outer * inner
    L = _lib.lib()
    ws = _workspace(L.g3d_ladder_workspace_bytes(S, len(rungs)), dev)
    rung = torch.empty((S,), dtype=torch.int32, device=dev)
    count = torch.empty((S,), dtype=torch.int32, device=dev)
    thr = torch.empty((S,), dtype=torch.float32, device=dev)
    check(L.g3d_threshold_ladder(_p(scores), outer, inner, N, outer_pitch, rungs.ctypes.data_as(ctypes.c_void_p),
                                 len(rungs), keep_max, _p(rung), _p(count), _p(thr), _p(ws), ws.numel(), _idx(dev),
                                 _stream(dev)), "g3d_threshold_ladder")
    return rung, count, thr


def filter_compact(scores, outer, inner, N, outer_pitch, thr, cap):
    """Indices n with score > thr[s] per vector.  Returns (idx i32[S,cap] (arrival order), count i32[S])."""
    dev = _need_cuda(scores, thr)
    if scores.dtype != torch.float32 or not scores.is_contiguous():
        raise Geom3dError("filter_compact needs a contiguous float32 score tensor")
    S = outer * inner
    thr = _prep(thr, torch.float32)
    idx = torch.empty((S, cap), dtype=torch.int32, device=dev)
    count = torch.empty((S,), dtype=torch.int32, device=dev)
    check(_lib.lib().g3d_filter_compact(_p(scores), outer, inner, N, outer_pitch, _p(thr), cap, _p(idx), _p(count),
                                        _idx(dev), _stream(dev)), "g3d_filter_compact")
    return idx, count


def gather_candidates(scores, outer, inner, N, outer_pitch, idx, count, cap, boxes=None, box_col=0):
    """Pack filter_compact's candidates (ascending index per vector).  Returns (seg_offsets i32[S+1], cand_scores f32[T],
    cand_boxes f32[T,4]|None, cand_src i32[T]) with T = S*cap rows allocated (only seg_offsets[S] are valid)."""
    dev = _need_cuda(scores, idx, count, boxes)
    S = outer * inner
    T = S * cap
    seg_offsets = torch.empty((S + 1,), dtype=torch.int32, device=dev)
    cand_scores = torch.empty((T,), dtype=torch.float32, device=dev)
    cand_src = torch.empty((T,), dtype=torch.int32, device=dev)
    cand_boxes = None
    stride = 4
    if boxes is not None:
        if boxes.dtype != torch.float32 or not boxes.is_contiguous():
            raise Geom3dError("gather_candidates needs contiguous float32 boxes")
        stride = boxes.shape[-1]
        cand_boxes = torch.empty((T, 4), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_gather_candidates(_p(scores), outer, inner, N, outer_pitch, _p(boxes), stride, box_col, _p(idx),
                                           _p(count), cap, _p(seg_offsets), _p(cand_scores), _p(cand_boxes), _p(cand_src),
                                           _idx(dev), _stream(dev)), "g3d_gather_candidates")
    return seg_offsets, cand_scores, cand_boxes, cand_src


def _decode_args(anchors, regression, mean, std, clip_wh):
    """common argument block of the decode-on-the-fly entry points"""
    anc = _prep(anchors, torch.float32)
    if anc.dim() == 2:
        anc = anc.unsqueeze(0)
    reg = _prep(regression, torch.float32)
    if reg.dim() != 3 or reg.shape[2] not in (4, 12) or anc.dim() != 3 or anc.shape[1] != reg.shape[1] or anc.shape[2] != 4:
        raise ValueError(f"expected anchors[1|B,A,4] and regression[B,A,12|4], got {tuple(anc.shape)} and {tuple(reg.shape)}")
    variant = VARIANT_3D if reg.shape[2] == 12 else VARIANT_2D
    mean_h = (ctypes.c_float * 4)(*[float(x) for x in mean]) if mean is not None else None
    std_h = (ctypes.c_float * 4)(*[float(x) for x in std]) if std is not None else None
    if variant == VARIANT_2D and (mean_h is None or std_h is None):
        raise ValueError("the 2D decode needs mean and std")
    clip = 0 if clip_wh is None else 1
    cw, ch = (0.0, 0.0) if clip_wh is None else (float(clip_wh[0]), float(clip_wh[1]))
    return anc, reg, variant, mean_h, std_h, clip, cw, ch


def gather_candidates_decoded(scores, outer, inner, N, outer_pitch, idx, count, cap, anchors, regression, mean=None,
                              std=None, clip_wh=None):
    """gather_candidates whose NMS boxes are decoded on the fly from (anchors, regression) - no decoded tensor needed.
    Returns (seg_offsets i32[S+1], cand_scores f32[T], cand_boxes f32[T,4], cand_src i32[T])."""
    dev = _need_cuda(scores, idx, count, anchors, regression)
    anc, reg, variant, mean_h, std_h, clip, cw, ch = _decode_args(anchors, regression, mean, std, clip_wh)
    S = outer * inner
    T = S * cap
    seg_offsets = torch.empty((S + 1,), dtype=torch.int32, device=dev)
    cand_scores = torch.empty((T,), dtype=torch.float32, device=dev)
    cand_src = torch.empty((T,), dtype=torch.int32, device=dev)
    cand_boxes = torch.empty((T, 4), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_gather_candidates_decoded(_p(scores), outer, inner, N, outer_pitch, _p(anc), anc.shape[0], _p(reg),
                                                   variant, mean_h, std_h, clip, cw, ch, _p(idx), _p(count), cap,
                                                   _p(seg_offsets), _p(cand_scores), _p(cand_boxes), _p(cand_src),
                                                   _idx(dev), _stream(dev)), "g3d_gather_candidates_decoded")
    return seg_offsets, cand_scores, cand_boxes, cand_src


class TailPlan(dict):
    """Buffers of one detection tail - count, keep_count, seg_offsets, out_offsets, summary, cand_src, cand_scores, keep
    (dict entries) and the workspace - for S = outer*inner segments of capacity cap on one device, with their ctypes
    pointers made once.  A caller that runs the tail every frame keeps the plan: launch() and assemble() are then one
    ctypes call each with no allocation besides the four result tensors (postprocess._tail_run)."""

    def __init__(self, outer, inner, N, cap, dev):
        super().__init__()
        S, T = outer * inner, outer * inner * cap
        L = _lib.lib()
        self.shape = (outer, inner, N, cap, dev)
        # one int32 block for the small arrays and the candidate tables, one int64 keep list, one workspace
        small = torch.empty((4 * S + 6 + 2 * T + 64,), dtype=torch.int32, device=dev)
        o = 0
        self["count"] = small[o:o + S]; o += S
        self["keep_count"] = small[o:o + S]; o += S
        self["seg_offsets"] = small[o:o + S + 1]; o += S + 1
        self["out_offsets"] = small[o:o + S + 1]; o += S + 1
        self["summary"] = small[o:o + 4]; o += 4
        o = (o + 3) & ~3                                        # 16-byte alignment of the two tables
        self["cand_src"] = small[o:o + T]; o += T
        o = (o + 3) & ~3
        self["cand_scores"] = small[o:o + T].view(torch.float32)
        self["keep"] = torch.empty((T,), dtype=torch.int64, device=dev)
        self.ws = _workspace(L.g3d_detect_tail_workspace_bytes(S, cap), dev)
        self.ptr = {k: _p(v) for k, v in self.items()}
        self.ws_ptr, self.ws_bytes, self.dev_idx = _p(self.ws), self.ws.numel(), _idx(dev)

    def launch(self, short, scores, outer_pitch, thr, decode, iou_threshold, stream, summary_ptr=None):
        """the tail up to summary; `decode` = _decode_args(...) of the anchors / regression this tail decodes from;
        summary_ptr: where the 4 summary integers go instead of self["summary"] (mapped pinned host memory, see the header)"""
        outer, inner, N, cap, _ = self.shape
        anc, reg, variant, mean_h, std_h, clip, cw, ch = decode
        q = self.ptr
        L = _lib.lib()
        fn = L.g3d_detect_tail_short if short else L.g3d_detect_tail
        check(fn(_p(scores), outer, inner, N, outer_pitch, _p(thr), cap, _p(anc), anc.shape[0], _p(reg), variant,
                 mean_h, std_h, clip, cw, ch, float(iou_threshold), q["count"], q["seg_offsets"], q["cand_scores"],
                 q["cand_src"], q["keep"], q["keep_count"], q["out_offsets"], q["summary"] if summary_ptr is None else summary_ptr,
                 self.ws_ptr, self.ws_bytes, self.dev_idx, stream), "g3d_detect_tail_short" if short else "g3d_detect_tail")

    def assemble(self, rows, decode, stream):
        """(scores f32[rows], classes i64[rows], boxes f32[rows,20|4], image i64[rows]) of the kept candidates"""
        outer, inner, N, _, dev = self.shape
        anc, reg, variant, mean_h, std_h, clip, cw, ch = decode
        q = self.ptr
        scores = torch.empty((rows,), dtype=torch.float32, device=dev)
        classes = torch.empty((rows,), dtype=torch.int64, device=dev)
        image = torch.empty((rows,), dtype=torch.int64, device=dev)
        boxes = torch.empty((rows, 20 if variant == VARIANT_3D else 4), dtype=torch.float32, device=dev)
        if rows:
            check(_lib.lib().g3d_assemble_detections(q["keep"], q["keep_count"], q["seg_offsets"], q["out_offsets"],
                                                     q["cand_scores"], q["cand_src"], outer, inner, N, _p(anc), anc.shape[0],
                                                     _p(reg), variant, mean_h, std_h, clip, cw, ch, _p(scores), _p(classes),
                                                     _p(boxes), _p(image), rows, self.dev_idx, stream),
                  "g3d_assemble_detections")
        return scores, classes, boxes, image


def detect_tail(scores, outer, inner, N, outer_pitch, thr, cap, anchors, regression, iou_threshold, mean=None, std=None,
                clip_wh=None, short=False, reuse=None):
    """filter_compact -> gather_candidates_decoded -> nms_segmented(relative=False) -> detection_offsets in ONE library
    call (g3d_detect_tail: the ~10 short launches are issued from C++ back to back).  Returns a TailPlan - a dict with
    count, seg_offsets, cand_scores, cand_src, keep, keep_count, out_offsets and summary i32[4] = (detections, max count,
    segments not processed, 0).  short=True: g3d_detect_tail_short - one launch for the whole chain, for segments of at
    most 1024 candidates; summary[2] > 0 tells the caller to repeat with short=False.  `reuse`: the plan of an earlier
    call with the same sizes on the same stream whose results are no longer needed - its buffers are written again."""
    dev = _need_cuda(scores, thr, anchors, regression)
    if scores.dtype != torch.float32 or not scores.is_contiguous():
        raise Geom3dError("detect_tail needs a contiguous float32 score tensor")
    decode = _decode_args(anchors, regression, mean, std, clip_wh)
    thr = _prep(thr, torch.float32)
    plan = reuse if reuse is not None and reuse.shape == (outer, inner, N, cap, dev) else TailPlan(outer, inner, N, cap, dev)
    plan.launch(short, scores, outer_pitch, thr, decode, iou_threshold, _stream(dev))
    return plan


def detection_offsets(keep_count):
    """out_offsets i32[S+1] = exclusive scan of the per-segment keep counts (out_offsets[S] = number of detections)."""
    dev = _need_cuda(keep_count)
    S = keep_count.numel()
    out_offsets = torch.empty((S + 1,), dtype=torch.int32, device=dev)
    check(_lib.lib().g3d_exclusive_scan_i32(_p(keep_count), S, _p(out_offsets), _idx(dev), _stream(dev)),
          "g3d_exclusive_scan_i32")
    return out_offsets


def assemble_detections(keep, keep_count, seg_offsets, cand_scores, cand_src, outer, inner, N, anchors, regression,
                        mean=None, std=None, clip_wh=None, out_offsets=None, K=None):
    """Final (scores f32[K], classes i64[K], boxes f32[K,20|4], image_index i64[K]) from the per-segment keep lists of
    nms_segmented(relative=False): one scan, ONE 4-byte device->host read (K), one assembly launch.  A caller that
    pipelines several batches passes out_offsets (detection_offsets) and K once it has read it; K may also be an estimate
    made BEFORE the count is known - rows beyond it are simply not written (the caller checks and repeats if short)."""
    dev = _need_cuda(keep, keep_count, seg_offsets, cand_scores, cand_src, anchors, regression)
    anc, reg, variant, mean_h, std_h, clip, cw, ch = _decode_args(anchors, regression, mean, std, clip_wh)
    S = outer * inner
    L = _lib.lib()
    if out_offsets is None:
        out_offsets = detection_offsets(keep_count)
    if K is None:
        K = int(out_offsets[S].item())
    cols = 20 if variant == VARIANT_3D else 4
    scores = torch.empty((K,), dtype=torch.float32, device=dev)
    classes = torch.empty((K,), dtype=torch.int64, device=dev)
    image = torch.empty((K,), dtype=torch.int64, device=dev)
    boxes = torch.empty((K, cols), dtype=torch.float32, device=dev)
    if K:
        check(L.g3d_assemble_detections(_p(keep), _p(keep_count), _p(seg_offsets), _p(out_offsets), _p(cand_scores),
                                        _p(cand_src), outer, inner, N, _p(anc), anc.shape[0], _p(reg), variant, mean_h,
                                        std_h, clip, cw, ch, _p(scores), _p(classes), _p(boxes), _p(image), K, _idx(dev),
                                        _stream(dev)), "g3d_assemble_detections")
    return scores, classes, boxes, image


# ------------------------------------------------------------------------------------------------ a11: NMS
def nms_segmented(boxes, scores, seg_offsets, max_seg_len, iou_threshold, box_col=0, relative=True):
    """Greedy NMS on S independent segments.  boxes[N,K] (box columns box_col..box_col+3), scores[N],
    seg_offsets i32[S+1] on the device.  Returns (keep i64[N] - segment s's kept indices start at seg_offsets[s] -,
    keep_count i32[S])."""
    dev = _need_cuda(boxes, scores, seg_offsets)
    bx = _prep(boxes, torch.float32)
    sc = _prep(scores, torch.float32)
    so = _prep(seg_offsets, torch.int32) if seg_offsets is not None else None
    if bx.dim() != 2 or bx.shape[1] < box_col + 4 or sc.dim() != 1 or sc.shape[0] != bx.shape[0]:
        raise ValueError(f"expected boxes[N,>=4] and scores[N], got {tuple(bx.shape)} and {tuple(sc.shape)}")
    N, K = bx.shape
    S = so.numel() - 1 if so is not None else 1
    keep = torch.empty((N,), dtype=torch.int64, device=dev)
    keep_count = torch.empty((max(S, 0),), dtype=torch.int32, device=dev)
    L = _lib.lib()
    max_seg_len = int(min(max_seg_len, N))
    ws = _workspace(L.g3d_nms_workspace_bytes(N, S, max_seg_len), dev)
    check(L.g3d_nms_segmented(_p(bx), K, box_col, _p(sc), N, _p(so), S, max_seg_len, float(iou_threshold),
                              1 if relative else 0, _p(keep), _p(keep_count), _p(ws), ws.numel(), _idx(dev),
                              _stream(dev)), "g3d_nms_segmented")
    return keep, keep_count


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms drop-in: boxes[N,4], scores[N] -> int64 kept indices in descending score order."""
    dev = _need_cuda(boxes, scores)
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError(f"boxes should be a 2d tensor of shape [N,4], got {tuple(boxes.shape)}")
    if scores.dim() != 1 or scores.shape[0] != boxes.shape[0]:
        raise ValueError("boxes and scores should have the same number of elements in dimension 0")
    N = boxes.shape[0]
    if N == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    keep, cnt = nms_segmented(boxes, scores, None, N, iou_threshold, 0, True)   # one segment [0, N)
    return keep[: int(cnt.item())]


# ------------------------------------------------------------------------------------------------ a12-a19: homography
def _cam_args(cam, dev, d):
    """cam: None/int -> constant camera; uint8 tensor[d] -> per-object."""
    if cam is None:
        return None, 0
    if isinstance(cam, int):
        return None, cam
    cam = cam.to(device=dev, dtype=torch.uint8).contiguous()
    if cam.numel() != d:
        raise ValueError(f"camera index tensor has {cam.numel()} entries for {d} objects")
    return cam, 0


def _pts_f(pts):
    """float32 stays float32, everything else is promoted to float64 (the reference calls .double() on it anyway)"""
    return _prep(pts, torch.float32 if pts.dtype == torch.float32 else torch.float64)


def state_to_space(states):
    """homography.py:305-320.  states[d,>=6] -> float32 [d,8,3]."""
    dev = _need_cuda(states)
    st = _prep(states, torch.float32)
    if st.dim() != 2 or st.shape[1] < 6:
        raise ValueError(f"states must be [d,>=6], got {tuple(st.shape)}")
    d, S = st.shape
    out = torch.empty((d, 8, 3), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_state_to_space(_p(st), d, S, _p(out), _idx(dev), _stream(dev)), "g3d_state_to_space")
    return out


def space_to_im(pts, P, cam=None, wrapper=False):
    """homography.py:438-476.  pts[d,m,3], P f64[ncam,2,3,4] (device) -> float64 [d,m,2]."""
    dev = _need_cuda(pts, P)
    p = _pts_f(pts)
    if p.dim() != 3 or p.shape[2] != 3:
        raise ValueError(f"points must be [d,m,3], got {tuple(p.shape)}")
    d, m, _ = p.shape
    cam_t, cam_c = _cam_args(cam, dev, d)
    out = torch.empty((d, m, 2), dtype=torch.float64, device=dev)
    check(_lib.lib().g3d_space_to_im(_p(p), int(p.dtype == torch.float64), d, m, _p(P), P.shape[0], _p(cam_t), cam_c,
                                     int(wrapper), _p(out), _idx(dev), _stream(dev)), "g3d_space_to_im")
    return out


def state_to_im(states, P, cam=None, wrapper=False, all_cams=False, out_dtype=torch.float64):
    """homography.py:479-488 fused.  states[d,>=6] -> [d,8,2] (or [d,ncam,8,2] with all_cams) float64 (float32 opt-in)."""
    dev = _need_cuda(states, P)
    st = _prep(states, torch.float32)
    if st.dim() != 2 or st.shape[1] < 6:
        raise ValueError(f"states must be [d,>=6], got {tuple(st.shape)}")
    if out_dtype not in (torch.float64, torch.float32):
        raise ValueError("out_dtype must be float64 or float32")
    d, S = st.shape
    ncam = P.shape[0]
    cam_t, cam_c = _cam_args(cam, dev, d)
    shape = (d, ncam, 8, 2) if all_cams else (d, 8, 2)
    out = torch.empty(shape, dtype=out_dtype, device=dev)
    check(_lib.lib().g3d_state_to_im(_p(st), d, S, _p(P), ncam, _p(cam_t), cam_c, int(wrapper), int(all_cams), _p(out),
                                     int(out_dtype == torch.float32), _idx(dev), _stream(dev)), "g3d_state_to_im")
    return out


def _pts_heights(pts, heights):
    p = _pts_f(pts)
    if p.dim() != 3 or p.shape[1] != 8 or p.shape[2] != 2:
        raise ValueError(f"points must be [d,8,2], got {tuple(p.shape)}")
    h = _prep(heights, p.dtype).reshape(-1)
    if h.numel() != p.shape[0]:
        raise ValueError(f"heights has {h.numel()} entries for {p.shape[0]} objects")
    return p, h


def im_to_space(pts, heights, H, cam=None, wrapper=False):
    """homography.py:388-435.  pts[d,8,2], heights[d], H f64[ncam,2,3,3] -> float64 [d,8,3]."""
    dev = _need_cuda(pts, heights, H)
    p, h = _pts_heights(pts, heights)
    d = p.shape[0]
    cam_t, cam_c = _cam_args(cam, dev, d)
    out = torch.empty((d, 8, 3), dtype=torch.float64, device=dev)
    check(_lib.lib().g3d_im_to_space(_p(p), _p(h), int(p.dtype == torch.float64), d, _p(H), H.shape[0], _p(cam_t), cam_c,
                                     int(wrapper), _p(out), _idx(dev), _stream(dev)), "g3d_im_to_space")
    return out


def space_to_state(pts):
    """homography.py:274-303.  pts[d,8,3] -> float32 [d,6]."""
    dev = _need_cuda(pts)
    p = _pts_f(pts)
    if p.dim() != 3 or p.shape[1] != 8 or p.shape[2] != 3:
        raise ValueError(f"points must be [d,8,3], got {tuple(p.shape)}")
    d = p.shape[0]
    out = torch.empty((d, 6), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_space_to_state(_p(p), int(p.dtype == torch.float64), d, _p(out), _idx(dev), _stream(dev)),
          "g3d_space_to_state")
    return out


def im_to_state(pts, heights, H, cam=None, wrapper=False):
    """homography.py:491-500 fused.  -> float32 [d,6]."""
    dev = _need_cuda(pts, heights, H)
    p, h = _pts_heights(pts, heights)
    d = p.shape[0]
    cam_t, cam_c = _cam_args(cam, dev, d)
    out = torch.empty((d, 6), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_im_to_state(_p(p), _p(h), int(p.dtype == torch.float64), d, _p(H), H.shape[0], _p(cam_t), cam_c,
                                     int(wrapper), _p(out), _idx(dev), _stream(dev)), "g3d_im_to_state")
    return out


def im_to_state_refined(pts, heights, H, P, cam=None, wrapper=False, return_heights=False):
    """The trackers' two-pass idiom (MC3D_crop_tracker.py:364-370) in one kernel.  -> float32 [d,6] (, f64 heights)."""
    dev = _need_cuda(pts, heights, H, P)
    p, h = _pts_heights(pts, heights)
    d = p.shape[0]
    cam_t, cam_c = _cam_args(cam, dev, d)
    out = torch.empty((d, 6), dtype=torch.float32, device=dev)
    hout = torch.empty((d,), dtype=torch.float64, device=dev) if return_heights else None
    check(_lib.lib().g3d_im_to_state_refined(_p(p), _p(h), int(p.dtype == torch.float64), d, _p(H), _p(P), H.shape[0],
                                             _p(cam_t), cam_c, int(wrapper), _p(out), _p(hout), _idx(dev), _stream(dev)),
          "g3d_im_to_state_refined")
    return (out, hout) if return_heights else out


def height_from_template(template_boxes, template_space_heights, boxes):
    """homography.py:519-551 with torch's type promotion.  -> [d] float64 if any input is, else float32."""
    dev = _need_cuda(template_boxes, template_space_heights, boxes)
    tb, th, bx = _pts_f(template_boxes), _pts_f(template_space_heights).reshape(-1), _pts_f(boxes)
    d = tb.shape[0]
    if tuple(tb.shape[1:]) != (8, 2) or tuple(bx.shape) != tuple(tb.shape) or th.numel() != d:
        raise ValueError("expected template_boxes[d,8,2], template_space_heights[d], boxes[d,8,2]")
    any64 = torch.float64 in (tb.dtype, th.dtype, bx.dtype)
    out = torch.empty((d,), dtype=torch.float64 if any64 else torch.float32, device=dev)
    check(_lib.lib().g3d_height_from_template(_p(tb), int(tb.dtype == torch.float64), _p(th), int(th.dtype == torch.float64),
                                              _p(bx), int(bx.dtype == torch.float64), d, _p(out), _idx(dev), _stream(dev)),
          "g3d_height_from_template")
    return out


def state_footprint(states):
    """Road-plane footprint (xmin,ymin,xmax,ymax) of each state (MC3D_crop_tracker.py:625-632). -> float32 [d,4]."""
    dev = _need_cuda(states)
    st = _prep(states, torch.float32)
    if st.dim() != 2 or st.shape[1] < 6:
        raise ValueError(f"states must be [d,>=6], got {tuple(st.shape)}")
    d, S = st.shape
    out = torch.empty((d, 4), dtype=torch.float32, device=dev)
    check(_lib.lib().g3d_state_footprint(_p(st), d, S, _p(out), _idx(dev), _stream(dev)), "g3d_state_footprint")
    return out


def corners_to_box(pts):
    """min/max box of 8 image corners (MC3D_crop_tracker.py:602-607).  pts[d,8,2] -> [d,4] same dtype."""
    dev = _need_cuda(pts)
    p = _pts_f(pts)
    if p.dim() != 3 or p.shape[1] != 8 or p.shape[2] != 2:
        raise ValueError(f"points must be [d,8,2], got {tuple(p.shape)}")
    d = p.shape[0]
    out = torch.empty((d, 4), dtype=p.dtype, device=dev)
    check(_lib.lib().g3d_corners_to_box(_p(p), int(p.dtype == torch.float64), d, _p(out), _idx(dev), _stream(dev)),
          "g3d_corners_to_box")
    return out


# ------------------------------------------------------------------------------------------------ a20: md_iou
def pairwise_iou(first, second, eps=0.0, one_minus=False):
    """out[i,j] = IoU(first[i], second[j]) in float64 (MC3D_crop_tracker.py:1030-1049 on un-broadcast inputs)."""
    dev = _need_cuda(first, second)
    a, b = _prep(first, torch.float32), _prep(second, torch.float32)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != 4 or b.shape[1] != 4:
        raise ValueError("pairwise_iou expects first[n,4] and second[m,4]")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=dev)
    check(_lib.lib().g3d_pairwise_iou_f64(_p(a), a.shape[0], _p(b), b.shape[0], float(eps), int(one_minus), _p(out),
                                          _idx(dev), _stream(dev)), "g3d_pairwise_iou_f64")
    return out


def md_iou(a, b):
    """Literal md_iou: a, b [..., 4] float64 of equal shape -> IoU [...] float64."""
    dev = _need_cuda(a, b)
    if a.shape != b.shape or a.shape[-1] != 4:
        raise ValueError("md_iou expects two [...,4] tensors of equal shape")
    a64, b64 = _prep(a, torch.float64), _prep(b, torch.float64)
    out = torch.empty(a.shape[:-1], dtype=torch.float64, device=dev)
    check(_lib.lib().g3d_md_iou(_p(a64), _p(b64), out.numel(), _p(out), _idx(dev), _stream(dev)), "g3d_md_iou")
    return out


# ------------------------------------------------------------------------------------------------ §8(f)-4: Kalman filter
def _host_f32(t, n):
    a = np.ascontiguousarray(np.asarray(t.detach().cpu() if isinstance(t, torch.Tensor) else t, dtype=np.float32).reshape(-1))
    if a.size != n:
        raise ValueError(f"expected {n} coefficients, got {a.size}")
    return a


def kf_predict_(X, P, D, dt, F, Q, dt_default, T=None):
    """Torch_KF.predict in place (util_track/kf.py:292-336).  X f32[n,S], P f32[n,S,S], D f32[n] (direction), T f64[n]|None
    (device, contiguous); dt: python float or a float64 device tensor [n]; F, Q: [S,S] (any tensor / array, read on the host)."""
    dev = _need_cuda(X, P, D, T)
    if X.dtype != torch.float32 or P.dtype != torch.float32 or not X.is_contiguous() or not P.is_contiguous():
        raise Geom3dError("kf_predict_ works in place: X and P must be contiguous float32")
    n, S = X.shape
    Fh, Qh = _host_f32(F, S * S), _host_f32(Q, S * S)
    Dd = _prep(D, torch.float32) if D is not None else None
    if isinstance(dt, torch.Tensor):
        dtd, dts = _prep(dt.to(dev), torch.float64), 0.0
        if dtd.numel() != n:
            raise ValueError("dt tensor must have one entry per object")
    else:
        dtd, dts = None, float(dt)
    if T is not None and (T.dtype != torch.float64 or not T.is_contiguous()):
        raise Geom3dError("T must be a contiguous float64 tensor")
    check(_lib.lib().g3d_kf_predict(_p(X), _p(P), _p(Dd), _p(dtd), dts, float(dt_default), _p(T), n, S,
                                    Fh.ctypes.data_as(ctypes.c_void_p), Qh.ctypes.data_as(ctypes.c_void_p), _idx(dev),
                                    _stream(dev)), "g3d_kf_predict")
    return X, P


def kf_update_(X, P, rows, z, H, R, mu_R=None):
    """Torch_KF.update in place (util_track/kf.py:339-403) for the objects `rows` (int64 device tensor, distinct) with
    measurements z[m,M] (promoted to float64 like the reference)."""
    dev = _need_cuda(X, P, rows, z)
    if X.dtype != torch.float32 or P.dtype != torch.float32 or not X.is_contiguous() or not P.is_contiguous():
        raise Geom3dError("kf_update_ works in place: X and P must be contiguous float32")
    n, S = X.shape
    zz = _prep(z, torch.float64)
    m, M = zz.shape
    rr = _prep(rows, torch.int64)
    if rr.numel() != m:
        raise ValueError("one row index per measurement")
    Hh, Rh = _host_f32(H, M * S), _host_f32(R, M * M)
    muh = _host_f32(mu_R, M) if mu_R is not None else None
    check(_lib.lib().g3d_kf_update(_p(X), _p(P), _p(rr), _p(zz), m, S, M, Hh.ctypes.data_as(ctypes.c_void_p),
                                   Rh.ctypes.data_as(ctypes.c_void_p),
                                   muh.ctypes.data_as(ctypes.c_void_p) if muh is not None else None, _idx(dev),
                                   _stream(dev)), "g3d_kf_update")
    return X, P


def generate_anchors(level_shapes, strides, rows, cols, device):
    """Anchors.forward on the device (anchors.py:21-40): float32 [A,4].  level_shapes: float64 [L,S,4] base boxes per level
    (numpy, from generate_anchors); strides [L]; rows/cols [L]: feature-map sizes per level."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise Geom3dError(f"geom3d ops run on CUDA devices only (no CPU fallback): got {dev}")
    shp = np.ascontiguousarray(np.asarray(level_shapes, dtype=np.float64))
    if shp.ndim != 3 or shp.shape[2] != 4:
        raise ValueError("level_shapes must be [L,S,4]")
    L, S = int(shp.shape[0]), int(shp.shape[1])
    st = np.ascontiguousarray(np.asarray(strides, dtype=np.float64).reshape(-1))
    rr = np.ascontiguousarray(np.asarray(rows, dtype=np.int64).reshape(-1))
    cc = np.ascontiguousarray(np.asarray(cols, dtype=np.int64).reshape(-1))
    if not (st.size == rr.size == cc.size == L):
        raise ValueError("strides, rows and cols need one entry per level")
    A = int((rr * cc).sum()) * S
    out = torch.empty((A, 4), dtype=torch.float32, device=dev)
    vp = ctypes.c_void_p
    check(_lib.lib().g3d_generate_anchors(shp.ctypes.data_as(vp), st.ctypes.data_as(vp), rr.ctypes.data_as(vp),
                                          cc.ctypes.data_as(vp), L, S, _p(out), A, _idx(dev), _stream(dev)),
          "g3d_generate_anchors")
    return out


def cross_camera_pairs(footprints, cams, threshold):
    """estimate_ts_bias's pair mining (MC3D_crop_tracker.py:277-289): int64 [K,2] pairs (i, j), i < j, different cameras,
    float64 IoU of the float32 footprints[d,4] > threshold, in the reference's loop order.  One 4-byte read sizes K."""
    dev = _need_cuda(footprints, cams)
    fp = _prep(footprints, torch.float32)
    cm = _prep(cams, torch.int32).reshape(-1)
    if fp.dim() != 2 or fp.shape[1] != 4 or cm.numel() != fp.shape[0]:
        raise ValueError(f"expected footprints[d,4] and cams[d], got {tuple(fp.shape)} and {tuple(cm.shape)}")
    d = fp.shape[0]
    if d < 2:
        return torch.empty((0, 2), dtype=torch.int64, device=dev)
    L = _lib.lib()
    count = torch.empty((d,), dtype=torch.int32, device=dev)
    offsets = torch.empty((d + 1,), dtype=torch.int32, device=dev)
    null = ctypes.c_void_p(0)
    check(L.g3d_cross_camera_pairs(_p(fp), _p(cm), d, float(threshold), _p(count), null, null, 0, _idx(dev), _stream(dev)),
          "g3d_cross_camera_pairs")
    check(L.g3d_exclusive_scan_i32(_p(count), d, _p(offsets), _idx(dev), _stream(dev)), "g3d_exclusive_scan_i32")
    K = int(offsets[d].item())
    pairs = torch.empty((K, 2), dtype=torch.int32, device=dev)
    if K:
        check(L.g3d_cross_camera_pairs(_p(fp), _p(cm), d, float(threshold), null, _p(offsets), _p(pairs), K, _idx(dev),
                                       _stream(dev)), "g3d_cross_camera_pairs")
    return pairs.long()


if os.environ.get("G3D_TUNE"):
    _tuning_from_env()
