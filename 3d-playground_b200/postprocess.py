"""Host side of the decode / clip / score-filter / NMS drop-ins over the CUDA kernels.

Reference call sites:
  3D copy  pytorch_retinanet_detector_directional/retinanet/utils.py:82-167 (BBoxTransform 12->20, ClipBoxes),
           .../model.py:19-57 (batched_nms), :311-344 (MULTI_FRAME), :346-397 (default / LOCALIZE)
  2D copy  retinanet/utils.py:82-144, retinanet/model.py:270-311
  NMS      torchvision.ops.nms at all the sites listed in include/geom3d.h (a11)
"""
import ctypes
import os
import threading
import time

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import Geom3dError

KEEP_MAX = 10000          # `keep = 10000` of 3D model.py:322,367
LADDER_START_SINGLE = 1e-25   # 3D model.py:369
LADDER_START_MULTI = 1e-7     # 3D model.py:324
SCORE_THRESHOLD_2D = 0.05     # retinanet/model.py:289
NMS_IOU = 0.5                 # retinanet/model.py:297, 3D model.py:336,383


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms(boxes[N,4], scores[N], iou_threshold) -> int64[K] (descending score)."""
    return ops.nms(boxes, scores, iou_threshold)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """3D model.py:19-57, literally: boxes are shifted by idx * (max_coordinate + 1) in float32 (the reference's
    arithmetic, so that rounding of the shifted coordinates and cross-group overlaps of negative coordinates are
    reproduced), then one NMS over all boxes."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    boxes_for_nms = boxes + offsets[:, None]
    return ops.nms(boxes_for_nms, scores, iou_threshold)


class BBoxTransform3D(nn.Module):
    """pytorch_retinanet_detector_directional/retinanet/utils.py:82-149.  mean/std are kept as plain attributes as in
    the reference (they are dead there: the 12->20 decode never reads them)."""

    def __init__(self, mean=None, std=None):
        super().__init__()
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.mean = torch.zeros(4, dtype=torch.float32, device=dev) if mean is None else mean
        self.std = (0.1 * torch.ones([10], device=dev)) if std is None else std

    def forward(self, boxes, regression):
        return ops.decode3d(boxes, regression)


class BBoxTransform2D(nn.Module):
    """retinanet/utils.py:82-126.  mean / std are plain tensor attributes (default 0 and [.1,.1,.2,.2])."""

    def __init__(self, mean=None, std=None):
        super().__init__()
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.mean = torch.zeros(4, dtype=torch.float32, device=dev) if mean is None else mean
        self.std = torch.tensor([0.1, 0.1, 0.2, 0.2], dtype=torch.float32, device=dev) if std is None else std
        self._host = None

    def _host_params(self):
        key = (id(self.mean), getattr(self.mean, "_version", 0), id(self.std), getattr(self.std, "_version", 0))
        if self._host is None or self._host[0] != key:
            self._host = (key, [float(x) for x in self.mean.detach().cpu().reshape(-1)[:4]],
                          [float(x) for x in self.std.detach().cpu().reshape(-1)[:4]])
        return self._host[1], self._host[2]

    def forward(self, boxes, deltas, clip_wh=None):
        mean, std = self._host_params()
        return ops.decode2d(boxes, deltas, mean, std, clip_wh)


class ClipBoxes(nn.Module):
    """retinanet/utils.py:129-144 (and the unused 3D copy utils.py:152-167): in place, uses only img.shape."""

    def __init__(self, width=None, height=None):
        super().__init__()

    def forward(self, boxes, img):
        _, _, height, width = img.shape
        if boxes.is_contiguous() and boxes.dtype == torch.float32:
            ops.clip_boxes_(boxes, width, height)
        else:  # keep the in-place contract for views: clip a packed copy and write it back
            tmp = boxes.to(torch.float32).contiguous()
            ops.clip_boxes_(tmp, width, height)
            boxes.copy_(tmp)
        return boxes


# ---------------------------------------------------------------------------------------------- detection front end
def _segment_candidates(classification, boxes, box_col, thr, cap):
    """classification[B,A,C] f32, boxes[B,A,K], thr f32[B*C] (device).  Returns packed candidates per (image,class)."""
    B, A, C = classification.shape
    idx, count = ops.filter_compact(classification, B, C, A, A * C, thr, cap)
    seg_offsets, cand_scores, cand_boxes, cand_src = ops.gather_candidates(
        classification, B, C, A, A * C, idx, count, cap, boxes=boxes, box_col=box_col)
    return count, seg_offsets, cand_scores, cand_boxes, cand_src


def detect_per_class(classification, boxes, box_col=0, score_threshold=None, ladder_start=None, keep_max=KEEP_MAX,
                     iou_threshold=NMS_IOU, cap=None):
    """Per-(image, class) score filter + NMS for a whole batch in a handful of launches.

    classification[B,A,C] (post-sigmoid), boxes[B,A,K] decoded boxes whose NMS box is columns box_col..box_col+3.
    Exactly one of score_threshold (fixed `scores > thr`, retinanet/model.py:289) or ladder_start (adaptive ladder,
    3D model.py:368-374) must be given.
    Returns (scores f32[K], classes i64[K], boxes f32[K,Kcols], image_index i64[K]) ordered image-major, then class,
    then descending score: for B == 1 that is the reference's concatenation order (model.py:390-395)."""
    dev = classification.device
    if (score_threshold is None) == (ladder_start is None):
        raise ValueError("give exactly one of score_threshold / ladder_start")
    cls = ops._prep(classification, torch.float32)
    bx = ops._prep(boxes, torch.float32)
    B, A, C = cls.shape
    if ladder_start is not None:
        _, _, thr = ops.threshold_ladder(cls, B, C, A, A * C, ladder_start, keep_max)
        cap = min(keep_max, 16384) if cap is None else cap
    else:
        thr = torch.full((B * C,), float(np.float32(score_threshold)), dtype=torch.float32, device=dev)
        cap = 16384 if cap is None else cap
    cap = int(min(cap, max(A, 1)))
    count, seg_offsets, cand_scores, cand_boxes, cand_src = _segment_candidates(cls, bx, box_col, thr, cap)
    keep, keep_count = ops.nms_segmented(cand_boxes, cand_scores, seg_offsets, cap, iou_threshold, 0, relative=False)
    # one device->host read for the variable-length result (the reference synchronises ~100 times per image here)
    host = torch.stack((count, keep_count, seg_offsets[:-1])).cpu()
    counts, kcs, offs = host[0].tolist(), host[1].tolist(), host[2].tolist()
    if max(counts, default=0) > cap:
        raise Geom3dError(f"a (image, class) segment has {max(counts)} candidates above the score threshold but the "
                          f"candidate capacity is {cap}; pass a larger `cap` (<= 16384) or raise the threshold")
    pieces = [keep[o:o + k] for o, k in zip(offs, kcs) if k > 0]
    if not pieces:
        K = bx.shape[-1]
        return (torch.empty(0, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                torch.empty((0, K), device=dev), torch.empty(0, dtype=torch.int64, device=dev))
    sel = torch.cat(pieces)
    seg_of = torch.repeat_interleave(torch.arange(B * C, device=dev), torch.tensor(kcs, device=dev))
    image_index = seg_of // C
    classes = seg_of % C
    scores = cand_scores[sel]
    rows = image_index * A + cand_src[sel].to(torch.int64)
    out_boxes = bx.reshape(B * A, -1)[rows]
    return scores, classes, out_boxes, image_index


_TAIL_HOST = {}
_TAIL_ROWS = {}        # (device, segments) -> rows the speculative output buffers get (grows with what was seen)
_TAIL_GENERAL = {}     # (device, segments) -> calls that still go straight to the general chain (see _tail_run)
_SHORT_CAP = 1024      # longest segment g3d_detect_tail_short takes (detect_tail.cu: kShortCap)
_TAIL_BUFS = {}        # (device, stream, thread) -> intermediates of the last tail (rewritten by the next one on that stream)
_TAIL_THR = {}         # (device, segments, threshold) -> the constant threshold vector


def _tail_run(cls, regression, anchors, score_threshold, ladder_start, keep_max, iou_threshold, cap, mean, std, clip_wh):
    """The detection tail with ONE host synchronisation, at the very end.  Everything that needs no host decision is one
    library call; the assembly launch follows immediately into output buffers sized from an estimate (twice the largest
    result seen for this shape), so the GPU does not idle while the host learns the number of detections; only then the
    integers the host needs (detections, largest candidate count, segments left over) are read, through pinned memory.
    A result larger than the estimate is assembled again at its exact size.  (The call returns as soon as the counts are
    known; the assembly launch may still be running - on the caller's stream, like any other asynchronous torch op.)

    The library call is g3d_detect_tail_short - one launch for gather + decode + sort + NMS - as long as the segments are
    short (<= 1024 candidates; what scores > 0.05 leaves per class of an image); if it reports segments it did not take,
    the general chain (g3d_detect_tail) runs instead, and keeps running for this shape while the counts stay long (the
    threshold ladder of the 3D model keeps up to 10000 candidates per class).  Both give the same detections."""
    dev = ops._need_cuda(cls, regression, anchors)
    B, A, C = cls.shape
    if ladder_start is not None:
        _, _, thr = ops.threshold_ladder(cls, B, C, A, A * C, ladder_start, keep_max)
    else:
        tkey = (dev.index, B * C, float(np.float32(score_threshold)))
        thr = _TAIL_THR.get(tkey)
        if thr is None:
            if len(_TAIL_THR) > 64:
                _TAIL_THR.clear()
            thr = _TAIL_THR[tkey] = torch.full((B * C,), tkey[2], dtype=torch.float32, device=dev)
    rows_key = (dev.index, B * C)
    guess = _TAIL_ROWS.get(rows_key, 4096 * B)
    stream = torch.cuda.current_stream(dev)
    bkey = (dev.index, stream.cuda_stream, threading.get_ident())
    if bkey not in _TAIL_HOST:       # one pinned block (+ event) per stream and host thread, reused: each call ends with its own wait
        if len(_TAIL_HOST) >= 64:
            _TAIL_HOST.clear()
        pinned = torch.empty(4, dtype=torch.int32).pin_memory()
        _TAIL_HOST[bkey] = (pinned, pinned.numpy(), ctypes.c_void_p(pinned.data_ptr()), torch.cuda.Event())
    host, host_np, host_ptr, done = _TAIL_HOST[bkey]
    # The kernels write the summary straight into the pinned block (mapped host memory, entry 3 last behind a system
    # fence) and the host polls it: no copy, no event, and the host moves on while the assembly launch is still running.
    # G3D_TAIL_MAPPED=0: device buffer + asynchronous copy + event instead.
    mapped = os.environ.get("G3D_TAIL_MAPPED", "1") != "0"

    decode = ops._decode_args(anchors, regression, mean, std, clip_wh)     # validated once, shared by both launches
    plan = _TAIL_BUFS.get(bkey)
    if plan is None or plan.shape != (B, C, A, cap, dev):
        if len(_TAIL_BUFS) >= 8:
            _TAIL_BUFS.clear()
        # the intermediates (candidate tables, keep lists: S * cap entries each) are only read by the assembly launch
        # of the same call, on this stream: the next tail on the stream overwrites them
        plan = _TAIL_BUFS[bkey] = ops.TailPlan(B, C, A, cap, dev)
    sptr = ops._stream(dev)

    def run(short):
        if mapped:
            host_np[3] = -1
            plan.launch(short, cls, A * C, thr, decode, iou_threshold, sptr, summary_ptr=host_ptr)
            out = plan.assemble(guess, decode, sptr)
            spins, t_first = 0, None
            while host_np[3] < 0:
                spins += 1
                if not spins & 255:
                    time.sleep(0)                      # let other host threads run
                    t_first = t_first or time.monotonic()
                    if time.monotonic() - t_first > 20.0:        # a failed launch never writes the flag
                        torch.cuda.synchronize(dev)              # (raises the CUDA error, if there is one)
                        if host_np[3] < 0:
                            raise Geom3dError("detection tail: the summary was never written")
            K, most, left = int(host_np[0]), int(host_np[1]), int(host_np[2])
        else:
            plan.launch(short, cls, A * C, thr, decode, iou_threshold, sptr)
            out = plan.assemble(guess, decode, sptr)
            host.copy_(plan["summary"], non_blocking=True)
            done.record(stream)
            done.synchronize()
            K, most, left, _ = host.tolist()
        if most > cap:
            raise Geom3dError(f"a (image, class) segment has {most} candidates above the score threshold but the candidate "
                              f"capacity is {cap}; pass a larger `cap` (<= 16384) or raise the threshold")
        if left == 0 and K > guess:
            out = plan.assemble(K, decode, sptr)
        return out, K, most, left

    general = _TAIL_GENERAL.get(rows_key, 0)
    short = iou_threshold >= 0 and general == 0
    out, K, most, left = run(short)
    if short and left:
        out, K, most, _ = run(False)
        _TAIL_GENERAL[rows_key] = 16
    elif not short:
        _TAIL_GENERAL[rows_key] = general - 1 if (general > 0 and most <= _SHORT_CAP) else 16
    _TAIL_ROWS[rows_key] = max(guess, 2 * K)
    return tuple(x[:K] for x in out)


def detect_per_class_fused(classification, regression, anchors, score_threshold=None, ladder_start=None, keep_max=KEEP_MAX,
                           iou_threshold=NMS_IOU, cap=None, mean=None, std=None, clip_wh=None):
    """detect_per_class without the decoded tensor: the score filter runs first, only the candidates' NMS
    boxes and the kept rows are decoded - from regression[B,A,12] (3D directional model: NMS on columns 16..19, 20-column
    rows out) or regression[B,A,4] (2D model: mean / std / optional clip as BBoxTransform + ClipBoxes).
    Same result, bit for bit, as decode -> detect_per_class.  Everything up to the assembly is ONE library call
    (g3d_detect_tail_short: score filter, one launch for gather + decode + sort + NMS of every segment, offsets - or the
    general chain g3d_detect_tail when a segment is longer than 1024 candidates); the assembly launch follows at once into
    buffers sized from an estimate, and the single 16-byte device->host read comes last (_tail_run).  (Pipelining image
    groups over two streams - filter of group k+1 beside the segment kernel of group k - was tried twice, from Python and
    from inside the library: the segment kernel is latency-bound per CTA and the filter's CTAs hold the SMs, so nothing
    overlaps and every extra group costs its launches.)
    Returns (scores f32[K], classes i64[K], boxes f32[K,20|4], image_index i64[K])."""
    if (score_threshold is None) == (ladder_start is None):
        raise ValueError("give exactly one of score_threshold / ladder_start")
    cls = ops._prep(classification, torch.float32)
    reg = ops._prep(regression, torch.float32)
    A = cls.shape[1]
    if cap is None:
        cap = min(keep_max, 16384) if ladder_start is not None else 16384
    cap = int(min(cap, max(A, 1)))
    return _tail_run(cls, reg, anchors, score_threshold, ladder_start, keep_max, iou_threshold, cap, mean, std, clip_wh)


def detect_multi_frame(classification, boxes, box_col=16, ladder_start=LADDER_START_MULTI, keep_max=KEEP_MAX,
                       iou_threshold=NMS_IOU):
    """MULTI_FRAME branch of the 3D model (model.py:311-344): max over classes, one ladder over the whole batch,
    batched_nms with the image index as the group.  Returns (scores, classes, boxes[K,20], imIndexes)."""
    cls = ops._prep(classification, torch.float32)
    bx = ops._prep(boxes, torch.float32)
    B, A, C = cls.shape
    smax, amax = ops.rowmax(cls)                                        # [B,A]
    flat = smax.reshape(-1)
    N = flat.numel()
    _, _, thr = ops.threshold_ladder(flat, 1, 1, N, N, ladder_start, keep_max)
    cap = int(min(keep_max, 16384, max(N, 1)))
    idx, count = ops.filter_compact(flat, 1, 1, N, N, thr, cap)
    seg_offsets, cand_scores, _, cand_src = ops.gather_candidates(flat, 1, 1, N, N, idx, count, cap)
    n = int(seg_offsets[1].item())
    src = cand_src[:n].to(torch.int64)                                  # ascending flat index = boolean-mask order
    scores = cand_scores[:n]
    im_indexes = src // A
    anchor_boxes = bx.reshape(B * A, -1)[src]
    classes = amax.reshape(-1)[src]
    keep = batched_nms(anchor_boxes[:, box_col:box_col + 4], scores, im_indexes, iou_threshold)
    return scores[keep], classes[keep], anchor_boxes[keep], im_indexes[keep]


def _flatten_batch(classification, regression, anchors):
    """The reference's default branch squeezes the batch dimension (`torch.squeeze(classification[:, :, i])`, 3D
    model.py:366,381; 2D retinanet/model.py:288,295): for B == 1 that is one image, for B > 1 the boolean mask is [B,A] and
    the masked scores / boxes of ALL images are concatenated (image-major) before the ladder and the per-class NMS - the
    batch behaves like one image with B*A anchors.  Reproduced literally: the same tensors, flattened."""
    B, A, C = classification.shape
    if B == 1:
        return classification, regression, anchors
    anc = anchors.reshape(-1, A, anchors.shape[-1])
    if anc.shape[0] == 1:
        anc = anc.expand(B, A, anc.shape[-1])
    return (classification.reshape(1, B * A, C), regression.reshape(1, B * A, regression.shape[-1]),
            anc.reshape(1, B * A, anc.shape[-1]).contiguous())


class PostProcess3D(nn.Module):
    """Everything ResNet.forward does after the heads in the 3D directional model (model.py:306-397):
    forward(classification[B,A,8], regression[B,A,12], anchors[1,A,4], LOCALIZE=False, MULTI_FRAME=False)."""

    def __init__(self):
        super().__init__()
        self.regressBoxes = BBoxTransform3D()

    def forward(self, classification, regression, anchors, LOCALIZE=False, MULTI_FRAME=False):
        if MULTI_FRAME:
            return detect_multi_frame(classification, self.regressBoxes(anchors, regression))
        if LOCALIZE:
            return self.regressBoxes(anchors, regression), classification
        # default branch: filter first, decode only what survives (no [B,A,20] tensor).  A batch is processed the way the
        # reference processes it - as one flattened image (see _flatten_batch); for per-image detections of a batch use
        # detect_per_class_fused, which also returns the image index.
        cls1, reg1, anc1 = _flatten_batch(classification, regression, anchors)
        scores, classes, boxes, _ = detect_per_class_fused(cls1, reg1, anc1, ladder_start=LADDER_START_SINGLE)
        return [scores, classes, boxes]


class PostProcess2D(nn.Module):
    """Inference tail of the 2D model (retinanet/model.py:270-311): decode, clip, scores > 0.05, nms 0.5 per class."""

    def __init__(self):
        super().__init__()
        self.regressBoxes = BBoxTransform2D()
        self.clipBoxes = ClipBoxes()

    def forward(self, classification, regression, anchors, img_batch, LOCALIZE=False):
        _, _, height, width = img_batch.shape
        if LOCALIZE:
            return self.regressBoxes(anchors, regression, clip_wh=(width, height)), classification  # decode + clip fused
        mean, std = self.regressBoxes._host_params()
        cls1, reg1, anc1 = _flatten_batch(classification, regression, anchors)      # see PostProcess3D
        scores, classes, boxes, _ = detect_per_class_fused(cls1, reg1, anc1, score_threshold=SCORE_THRESHOLD_2D, mean=mean,
                                                           std=std, clip_wh=(width, height))
        return [scores, classes, boxes]
