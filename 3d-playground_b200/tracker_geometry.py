"""Tracker-frame geometry helpers (BASELINE config 5) over the CUDA kernels.

Same math as the methods of MC_Crop_Tracker (MC3D_crop_tracker.py) and minimal_3D_track.py that sit on the hot path:
md_iou (:1030-1049), im_nms (:592-615), space_nms (:617-635), the state -> footprint idiom (:625-632, :668-682,
:498-502), match_hungarian's cost matrix (:687-689), estimate_ts_bias's d x d IoU (:268-280) and select_best_box's
IoU + argmax (:974-1028).  The Hungarian solve, the Kalman filter and the tracker loop are out of scope.
CPU tensors are accepted (copied to the GPU and back); the arithmetic always runs in the kernels.
"""
import torch

from . import ops
from .homography_impl import _exec_device, _ret, _to_dev


def md_iou(a, b):
    """a, b: [n, m, 4] (any leading shape) float64 pre-broadcast boxes -> IoU [n, m] float64, no epsilon."""
    dev = _exec_device(a, b)
    return _ret(ops.md_iou(_to_dev(a, dev), _to_dev(b, dev)), a)


def pairwise_iou(first, second, eps=0.0):
    """IoU matrix [n, m] float64 of first[n,4] x second[m,4] without materialising the broadcast operands."""
    dev = _exec_device(first, second)
    return _ret(ops.pairwise_iou(_to_dev(first, dev), _to_dev(second, dev), eps=eps), first)


def state_footprint(states):
    """[d,6] states -> [d,4] float32 (xmin, ymin, xmax, ymax) of the bottom face in road-plane coordinates."""
    dev = _exec_device(states)
    return _ret(ops.state_footprint(_to_dev(states, dev)), states)


def association_cost(first_states, second_states):
    """dist = 1 - md_iou(footprint(first), footprint(second)) as float64 [f, s] (match_hungarian, :663-689)."""
    dev = _exec_device(first_states, second_states)
    fa = ops.state_footprint(_to_dev(first_states, dev))
    fb = ops.state_footprint(_to_dev(second_states, dev))
    return _ret(ops.pairwise_iou(fa, fb, one_minus=True), first_states)


def self_iou(states):
    """d x d footprint IoU of one set of states (estimate_ts_bias, :268-280).  Note the reference's operand order:
    iou[i, j] = md_iou(boxes[j], boxes[i])."""
    dev = _exec_device(states)
    fp = ops.state_footprint(_to_dev(states, dev))
    return _ret(ops.pairwise_iou(fp, fp).t().contiguous(), states)


def im_nms(detections, scores, threshold=0.8, groups=None):
    """detections[d,8,2] image corners -> kept indices (MC3D_crop_tracker.py:592-615).  As in the reference, `groups`
    only adds the same scalar 10000 to every box (the per-group offset is computed and discarded, :610-612)."""
    dev = _exec_device(detections, scores)
    boxes = ops.corners_to_box(_to_dev(detections, dev))
    if groups is not None:
        boxes = boxes + 10000
    keep = ops.nms(boxes.to(torch.float32), _to_dev(scores, dev).to(torch.float32), threshold)
    return _ret(keep, detections)


def space_nms(states, scores, threshold=0.1):
    """states[d,6] -> kept indices by NMS on the road-plane footprints (MC3D_crop_tracker.py:617-635)."""
    dev = _exec_device(states, scores)
    fp = ops.state_footprint(_to_dev(states, dev))
    keep = ops.nms(fp, _to_dev(scores, dev).to(torch.float32), threshold)
    return _ret(keep, states)


def select_best_box(a_priori, preds, confs, classes, n_objs, W):
    """select_best_box (MC3D_crop_tracker.py:974-1028): per object, the detection maximising (1-W)*IoU + W*conf."""
    dev = _exec_device(a_priori, preds)
    preds_d = _to_dev(preds, dev).reshape(-1, preds.shape[-1])
    d = preds_d.shape[0] // n_objs
    fp_pred = ops.state_footprint(preds_d).reshape(n_objs, d, 4)
    fp_prior = ops.state_footprint(_to_dev(a_priori, dev)).unsqueeze(1).expand(n_objs, d, 4)
    ious = ops.md_iou(fp_pred.double(), fp_prior.double().contiguous())
    confs_d, classes_d = _to_dev(confs, dev), _to_dev(classes, dev)
    scores = (1 - W) * ious + W * confs_d
    keep = torch.argmax(scores, dim=1)
    idx = torch.arange(n_objs, device=dev)
    best = preds_d.reshape(n_objs, d, -1)[idx, keep, :]
    return _ret(best, preds), _ret(classes_d[idx, keep], preds), _ret(confs_d[idx, keep], preds)
