"""ORACLE (test infrastructure): tracker-frame geometry.

    md_iou        MC3D_crop_tracker.py:1030-1049 (float64, no epsilon)
    footprint     MC3D_crop_tracker.py:625-632 (min/max of the 4 bottom space corners)
    im_box        MC3D_crop_tracker.py:602-607
    association_cost   MC3D_crop_tracker.py:663-689 (1 - md_iou on broadcast footprints)
"""
import torch

from . import homography_oracle as ho
from . import nms_oracle


def md_iou(a, b):
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    zero = torch.zeros(area_a.shape, dtype=torch.float64)
    iw = torch.max(zero, torch.min(a[..., 2], b[..., 2]) - torch.max(a[..., 0], b[..., 0]))
    ih = torch.max(zero, torch.min(a[..., 3], b[..., 3]) - torch.max(a[..., 1], b[..., 1]))
    inter = iw * ih
    return torch.div(inter, area_a + area_b - inter)


def footprint(states):
    sp = ho.state_to_space(states)
    out = torch.zeros(states.shape[0], 4)
    out[:, 0] = sp[:, 0:4, 0].min(dim=1).values
    out[:, 1] = sp[:, 0:4, 1].min(dim=1).values
    out[:, 2] = sp[:, 0:4, 0].max(dim=1).values
    out[:, 3] = sp[:, 0:4, 1].max(dim=1).values
    return out


def im_box(corners):
    return torch.stack((corners[:, :, 0].min(1).values, corners[:, :, 1].min(1).values,
                        corners[:, :, 0].max(1).values, corners[:, :, 1].max(1).values), dim=1)


def association_cost(first_states, second_states):
    fa, fb = footprint(first_states), footprint(second_states)
    f, s = fa.shape[0], fb.shape[0]
    return 1.0 - md_iou(fa.unsqueeze(1).repeat(1, s, 1).double(), fb.unsqueeze(0).repeat(f, 1, 1).double())


def space_nms(states, scores, threshold=0.1):
    return nms_oracle.nms(footprint(states), scores, threshold)


def im_nms(corners, scores, threshold=0.8, groups=None):
    boxes = im_box(corners)
    if groups is not None:
        boxes = boxes + 10000
    return nms_oracle.nms(boxes.float(), scores, threshold)
