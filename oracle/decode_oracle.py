"""ORACLE (test infrastructure): box decoding and clipping.

    decode3d   pytorch_retinanet_detector_directional/retinanet/utils.py:102-149 (12 -> 20)
    decode2d   retinanet/utils.py:102-126
    clip       retinanet/utils.py:134-144 (== 3D copy utils.py:157-167)
"""
import torch

from .losses_oracle import SIGN_H, SIGN_L, SIGN_W


def decode3d(boxes, regression):
    """boxes[1,A,4], regression[B,A,12] -> [B,A,20]; float32 ops in the reference's order."""
    w = boxes[:, :, 2] - boxes[:, :, 0]
    h = boxes[:, :, 3] - boxes[:, :, 1]
    cx = boxes[:, :, 0] + 0.5 * w
    cy = boxes[:, :, 1] + 0.5 * h
    r = regression
    out = torch.zeros(r.shape[0], r.shape[1], 20)
    for k in range(8):
        for c in (0, 1):
            v = r[:, :, c] + SIGN_L[k] * r[:, :, 2 + c]
            v = v + SIGN_W[k] * r[:, :, 4 + c]
            v = v + SIGN_H[k] * r[:, :, 6 + c]
            out[:, :, 2 * k + c] = v
    out[:, :, 16:20] = r[:, :, 8:12]
    out[:, :, 0::2] = out[:, :, 0::2] * w.unsqueeze(2) + cx.unsqueeze(2)
    out[:, :, 1::2] = out[:, :, 1::2] * h.unsqueeze(2) + cy.unsqueeze(2)
    return out


def decode2d(boxes, deltas, mean=None, std=None):
    mean = torch.zeros(4) if mean is None else mean
    std = torch.tensor([0.1, 0.1, 0.2, 0.2]) if std is None else std
    w = boxes[:, :, 2] - boxes[:, :, 0]
    h = boxes[:, :, 3] - boxes[:, :, 1]
    cx = boxes[:, :, 0] + 0.5 * w
    cy = boxes[:, :, 1] + 0.5 * h
    dx = deltas[:, :, 0] * std[0] + mean[0]
    dy = deltas[:, :, 1] * std[1] + mean[1]
    dw = deltas[:, :, 2] * std[2] + mean[2]
    dh = deltas[:, :, 3] * std[3] + mean[3]
    pcx = cx + dx * w
    pcy = cy + dy * h
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=2)


def clip(boxes, height, width):
    """returns a clipped copy (the reference clips in place)"""
    out = boxes.clone()
    out[..., 0] = out[..., 0].clamp(min=0)
    out[..., 1] = out[..., 1].clamp(min=0)
    out[..., 2] = out[..., 2].clamp(max=width)
    out[..., 3] = out[..., 3].clamp(max=height)
    return out
