"""ORACLE (test infrastructure): tracker-frame geometry.

    md_iou        MC3D_crop_tracker.py:1030-1049 (float64, no epsilon)
    footprint     MC3D_crop_tracker.py:625-632 (min/max of the 4 bottom space corners)
    im_box        MC3D_crop_tracker.py:602-607
    association_cost   MC3D_crop_tracker.py:663-689 (1 - md_iou on broadcast footprints)
    cross_camera_pairs / estimate_ts_bias   MC3D_crop_tracker.py:237-315

Parity pin: tests/test_oracle_golden.py::test_tracker / test_estimate_ts_bias against tests/golden/tracker.npz and
tests/golden/ts_bias.npz, produced by the UNMODIFIED MC_Crop_Tracker methods (tests/golden/make_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
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


def cross_camera_pairs(boxes, camera_idxs, phi):
    """the double loop of estimate_ts_bias (:277-289): (i, j), i < j, different cameras, iou[i, j] > phi, loop order"""
    fp = footprint(boxes)
    d = fp.shape[0]
    iou = md_iou(fp.unsqueeze(0).repeat(d, 1, 1).double(), fp.unsqueeze(1).repeat(1, d, 1).double()).reshape(d, d)
    cams = torch.as_tensor(camera_idxs).reshape(-1)
    hit = (iou > phi) & (cams[:, None] != cams[None, :])
    return torch.nonzero(torch.triu(hit, diagonal=1))          # row-major == i outer, j inner


def estimate_ts_bias(boxes, camera_idxs, objs, timestamps, ts_bias, mu_v, phi, alpha):
    """returns the updated copy of ts_bias (list of python floats), :251-315 step by step"""
    ts_bias = list(ts_bias)
    if len(camera_idxs) == 0 or len(objs) == 0:
        return ts_bias
    wb = objs[objs[:, 5] == -1, 6]
    eb = objs[objs[:, 5] == 1, 6]
    wb_vel = torch.mean(wb) * -1
    eb_vel = torch.mean(eb)
    if torch.isnan(wb_vel):
        wb_vel = torch.tensor(-float(mu_v))
    if torch.isnan(eb_vel):
        eb_vel = torch.tensor(float(mu_v))
    cams = [int(c) for c in camera_idxs]
    entries = []
    for i, j in cross_camera_pairs(boxes, camera_idxs, phi).tolist():
        entries.append((cams[i], cams[j], boxes[j, 0] - boxes[i, 0], boxes[i, 5]))
        entries.append((cams[j], cams[i], boxes[i, 0] - boxes[j, 0], boxes[i, 5]))
    if not entries:
        return ts_bias
    dx = torch.tensor([float(e[2]) for e in entries])
    dt_expected = torch.tensor([timestamps[e[1]] - timestamps[e[0]] for e in entries])
    vel = torch.ones(len(entries)) * eb_vel
    for k, e in enumerate(entries):
        if e[3] == -1:
            vel[k] = wb_vel
    time_error = dx / vel - dt_expected
    for k, te in enumerate(time_error):
        c1, c2 = entries[k][0], entries[k][1]
        if c1 != 0:
            ts_bias[c1] = float((1 - alpha) * ts_bias[c1] + alpha * (-te + ts_bias[c2]))
    return ts_bias


def select_best_box(a_priori, preds, confs, classes, n_objs, W):
    """MC3D_crop_tracker.py:974-1028: per object the detection maximising (1 - W) * IoU(footprint(pred), footprint(prior))
    + W * conf (float64 IoU of the float32 footprints, first index on ties).  preds[n*d,6] or [n,d,6]."""
    preds = preds.reshape(n_objs, -1, preds.shape[-1])
    d = preds.shape[1]
    fp_pred = footprint(preds.reshape(-1, preds.shape[-1])).reshape(n_objs, d, 4)
    fp_prior = footprint(a_priori).unsqueeze(1).repeat(1, d, 1)
    scores = (1 - W) * md_iou(fp_pred.double(), fp_prior.double()) + W * confs
    keep = torch.argmax(scores, dim=1)
    idx = torch.arange(n_objs)
    return preds[idx, keep, :], classes[idx, keep], confs[idx, keep]


def pairwise_iou_eps(a, b, eps=1e-6):
    """mot_evaluator.py:87-118 for every (i, j): intersection / (area_a + area_b - intersection + eps), float64"""
    a, b = a.double(), b.double()
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = (torch.min(a[:, None, 2], b[None, :, 2]) - torch.max(a[:, None, 0], b[None, :, 0])).clamp(min=0)
    ih = (torch.min(a[:, None, 3], b[None, :, 3]) - torch.max(a[:, None, 1], b[None, :, 1])).clamp(min=0)
    inter = iw * ih
    return inter / (area_a[:, None] + area_b[None, :] - inter + eps)


def remove_overlaps(boxes, frames_alive, phi_over):
    """MC3D_crop_tracker.py:482-518: NMS on the footprints with the number of frames alive as the confidence"""
    return nms_oracle.nms(footprint(boxes).float(), torch.as_tensor(frames_alive).float(), phi_over)


def parse_detections(scores, labels, boxes, camera_idxs, H1, H2, P1, P2, sigma_d, phi_im, phi_space, heights,
                     perform_nms=True, refine_height=False):
    """MC3D_crop_tracker.py:319-383 step by step on CPU tensors; H1/H2/P1/P2: per-camera matrices [ncam,3,3] / [ncam,3,4] of
    the wrapper's two homographies; heights[d]: the guess_heights values.  Returns (states, labels, scores, camera_idxs)."""
    keep = torch.where(scores > torch.ones(scores.shape) * sigma_d)
    labels, det, scores, cams, heights = labels[keep], boxes[keep], scores[keep], camera_idxs[keep], heights[keep]
    if len(det) == 0:
        return [], [], [], []
    det = det.reshape(-1, 10, 2)[:, :8, :]
    if perform_nms:
        idx = im_nms(det, scores, threshold=phi_im, groups=cams)
        labels, det, scores, cams, heights = labels[idx], det[idx], scores[idx], cams[idx], heights[idx]
    c = cams.long()
    states = ho.wrapper_im_to_state(det, H1[c], H2[c], heights)
    if refine_height:
        repro = ho.wrapper_state_to_im(states, P1[c], P2[c])
        refined = ho.height_from_template(repro, heights, det)
        states = ho.wrapper_im_to_state(det, H1[c], H2[c], refined)
    if perform_nms:
        idx = space_nms(states, scores, threshold=phi_space)
        labels, states, scores, cams = labels[idx], states[idx], scores[idx], cams[idx]
    return states, labels, scores, cams
