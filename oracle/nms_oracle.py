"""ORACLE (test infrastructure): greedy NMS, the score ladder and the detection post-processing glue.

torchvision.ops.nms is a third-party dependency of the reference and is NOT vendored in /root/reference; the reference
pins no version (no requirements file; Python-3.7-era bytecode suggests torchvision 0.5-0.10).  This module restates
its published algorithm (torchvision/csrc/ops/cpu/nms_kernel.cpp) and is pinned against the torchvision installed in
this image (0.26.0) through tests/golden/nms_*.npz:
    sort scores descending (stable); walk in that order; a kept box i suppresses every later box j with
    inter / (area_i + area_j - inter) > iou_threshold, all in float32, inter = max(0,dx) * max(0,dy);
    the float IoU is compared with the double threshold.

    nms              call sites: retinanet/model.py:297 ; 3D model.py:383 ; MC3D_crop_tracker.py:507,614,634
    batched_nms      3D model.py:19-57 (offset trick)
    ladder_threshold 3D model.py:368-374 (start 1e-25) and :322-328 (start 1e-7)
    detect_3d / detect_2d / detect_multi_frame   3D model.py:346-397, 2D retinanet/model.py:270-311, 3D model.py:311-344
"""
import numpy as np
import torch


def nms(boxes, scores, iou_threshold):
    """boxes[N,4] float32, scores[N] -> int64 kept indices in descending-score order (numpy, O(N * kept))."""
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu().numpy().astype(np.float32)
    n = b.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    with np.errstate(invalid="ignore", divide="ignore"):
        for pos in range(n):
            i = order[pos]
            if suppressed[i]:
                continue
            keep.append(i)
            rest = order[pos + 1:]
            w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


def batched_nms(boxes, scores, idxs, iou_threshold):
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    offsets = idxs.to(boxes) * (boxes.max() + 1)
    return nms(boxes + offsets[:, None], scores, iou_threshold)


def ladder_threshold(scores, start, keep=10000):
    """the reference's while loop: returns (mask of the last tested threshold, that threshold as float32)"""
    threshold = start
    count = 1000000
    mask = None
    while count > keep:
        mask = scores > threshold
        count = int(mask.sum())
        last = np.float32(threshold)
        threshold *= (10 ** .2)
    return mask, last


def detect_3d(classification, transformed, iou=0.5, start=1e-25):
    """default branch: classification[B,A,C], transformed[B,A,20] -> [scores, classes, boxes].  `torch.squeeze` of the
    reference (3D model.py:366,381) leaves [B,A] tensors for B > 1, and the boolean mask then concatenates the images:
    the batch is one flattened image."""
    S, Cl, Bx = [], [], []
    boxes = transformed.reshape(-1, transformed.shape[-1])
    for c in range(classification.shape[2]):
        sc = classification[:, :, c].reshape(-1)
        mask, _ = ladder_threshold(sc, start)
        if mask.sum() == 0:
            continue
        sc_k, bx_k = sc[mask], boxes[mask]
        keep = nms(bx_k[:, 16:20], sc_k, iou)
        S.append(sc_k[keep]); Cl.append(torch.full((keep.numel(),), c, dtype=torch.int64)); Bx.append(bx_k[keep])
    if not S:
        return [torch.empty(0), torch.empty(0, dtype=torch.int64), torch.empty(0, 20)]
    return [torch.cat(S), torch.cat(Cl), torch.cat(Bx)]


def detect_2d(classification, transformed, thr=0.05, iou=0.5):
    """2D retinanet/model.py:287-309; a batch is flattened exactly as in detect_3d (:288,295)"""
    S, Cl, Bx = [], [], []
    boxes = transformed.reshape(-1, transformed.shape[-1])
    for c in range(classification.shape[2]):
        sc = classification[:, :, c].reshape(-1)
        mask = sc > thr
        if mask.sum() == 0:
            continue
        sc_k, bx_k = sc[mask], boxes[mask]
        keep = nms(bx_k, sc_k, iou)
        S.append(sc_k[keep]); Cl.append(torch.full((keep.numel(),), c, dtype=torch.int64)); Bx.append(bx_k[keep])
    if not S:
        return [torch.empty(0), torch.empty(0, dtype=torch.int64), torch.empty(0, 4)]
    return [torch.cat(S), torch.cat(Cl), torch.cat(Bx)]


def detect_multi_frame(classification, transformed, iou=0.5, start=1e-7):
    B, A, C = classification.shape
    im_idx = torch.arange(B).unsqueeze(1).repeat(1, A).reshape(-1)
    boxes = transformed.reshape(-1, transformed.shape[2])
    scores, classes = classification.reshape(-1, C).max(dim=1)
    mask, _ = ladder_threshold(scores, start)
    scores, classes, boxes, im_idx = scores[mask], classes[mask], boxes[mask], im_idx[mask]
    keep = batched_nms(boxes[:, 16:20], scores, im_idx, iou)
    return scores[keep], classes[keep], boxes[keep], im_idx[keep]
