"""Tracker-frame geometry helpers (BASELINE config 5) over the CUDA kernels.

Same math as the methods of MC_Crop_Tracker (MC3D_crop_tracker.py) and minimal_3D_track.py that sit on the hot path:
md_iou (:1030-1049), im_nms (:592-615), space_nms (:617-635), the state -> footprint idiom (:625-632, :668-682,
:498-502), match_hungarian's cost matrix (:687-689), estimate_ts_bias's d x d IoU (:268-280) and select_best_box's
IoU + argmax (:974-1028), estimate_ts_bias (:237-315, pair mining on the GPU), and `FrameGeometry`: one frame's
cost matrix + space NMS + projection + image NMS replayed as a single CUDA graph.  The Hungarian solve and the tracker
loop are out of scope.  CPU tensors are accepted (copied to the GPU and back); the arithmetic always runs in the kernels.
"""
import numpy as np
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


def cross_camera_pairs(states, camera_idxs, threshold=0.1):
    """int64 [K,2]: every (i, j), i < j, of detections from different cameras whose road-plane footprints overlap with
    IoU > threshold, in the order of the reference's double loop (MC3D_crop_tracker.py:267-289)."""
    dev = _exec_device(states, camera_idxs)
    fp = ops.state_footprint(_to_dev(states, dev))
    cams = _to_dev(torch.as_tensor(camera_idxs), dev).to(torch.int32)
    return _ret(ops.cross_camera_pairs(fp, cams, threshold), states)


def estimate_ts_bias(boxes, camera_idxs, objs, timestamps, ts_bias, mu_v, phi_nms_space=0.1, ts_alpha=0.05):
    """MC_Crop_Tracker.estimate_ts_bias (MC3D_crop_tracker.py:237-315) as a function of the tracker's fields: boxes[d,6]
    detections in state form, camera_idxs[d], objs = filter.view(with_direction=True)[1] ([n,7]: ..., direction, speed),
    timestamps / ts_bias: per-camera lists, mu_v = filter.mu_v.  Updates ts_bias in place (python floats) and returns it.

    The d x d float64 IoU matrix and the O(d^2) Python loop of the reference become one pair-mining kernel; what is left
    on the host is the reference's own sequential relaxation over the K matched pairs (order-dependent by design)."""
    if len(camera_idxs) == 0 or len(objs) == 0:
        return ts_bias
    dev = _exec_device(boxes, objs)
    b = _to_dev(boxes, dev).to(torch.float32)
    # mean speed per direction (:258-265); an empty direction (NaN mean) falls back to the filter's prior.  These two
    # reductions over the few filter rows stay the reference's own CPU `torch.mean` calls: the FP32 summation order of
    # ATen's CPU kernel is part of the bias floats that come out, and no device reduction reproduces it.
    o = torch.as_tensor(objs).detach().cpu()
    speed, direction = o[:, 6], o[:, 5]
    means = torch.stack((speed[direction == -1].mean() * -1, speed[direction == 1].mean()))
    wb_vel = np.float32(-float(mu_v)) if torch.isnan(means[0]) else np.float32(means[0])
    eb_vel = np.float32(float(mu_v)) if torch.isnan(means[1]) else np.float32(means[1])
    cams_d = _to_dev(torch.as_tensor(camera_idxs), dev).to(torch.int32).reshape(-1)
    pairs = ops.cross_camera_pairs(ops.state_footprint(b), cams_d, phi_nms_space)
    if pairs.shape[0] == 0:
        return ts_bias
    i, j = pairs[:, 0], pairs[:, 1]
    rec = torch.stack((cams_d[i].float(), cams_d[j].float(), b[j, 0] - b[i, 0], b[i, 5]), dim=1).cpu().numpy()
    alpha32, keep = np.float32(ts_alpha), 1 - ts_alpha
    for ci, cj, dx, dr in rec:
        ci, cj = int(ci), int(cj)
        vel = wb_vel if dr == -1 else eb_vel
        # the reference appends (cam_i, cam_j, x_j - x_i) and then (cam_j, cam_i, x_i - x_j) for every pair (:288-289)
        for c1, c2, off in ((ci, cj, dx), (cj, ci, -dx)):
            if c1 == 0:                                   # every bias is relative to camera 0 (:314)
                continue
            te = np.float32(off) / vel - np.float32(timestamps[c2] - timestamps[c1])
            ts_bias[c1] = float(np.float32(keep * ts_bias[c1]) + alpha32 * (-te + np.float32(ts_bias[c2])))
    return ts_bias


def remove_overlaps(boxes, frames_alive, phi_over):
    """MC_Crop_Tracker.remove_overlaps (MC3D_crop_tracker.py:482-518) without the filter bookkeeping: boxes[n,>=6] are the
    tracked objects' states (the filter view advanced to the newest time stamp), frames_alive[n] the number of frames each
    has been tracked - used as the NMS confidence, so of two overlapping tracklets the younger one goes.  Returns
    (keepers int64[k] as torchvision.nms orders them, removed bool[n]); phi_over <= 0 keeps everything (:489)."""
    dev = _exec_device(boxes, frames_alive)
    n = boxes.shape[0]
    if n == 0 or phi_over <= 0:
        keep = torch.arange(n, dtype=torch.int64, device=dev)
        return _ret(keep, boxes), _ret(torch.zeros(n, dtype=torch.bool, device=dev), boxes)
    fp = ops.state_footprint(_to_dev(boxes, dev))
    keep = ops.nms(fp, _to_dev(torch.as_tensor(frames_alive), dev).to(torch.float32), phi_over)
    removed = torch.ones(n, dtype=torch.bool, device=dev)
    removed[keep] = False
    return _ret(keep, boxes), _ret(removed, boxes)


def parse_detections(hg, scores, labels, boxes, camera_idxs, cameras, sigma_d, phi_nms_im, phi_nms_space, n_best=200,
                     perform_nms=True, refine_height=False, heights=None):
    """MC_Crop_Tracker.parse_detections (MC3D_crop_tracker.py:319-383) with the detections staying on the GPU from the
    detector's output to the tracker's state space (the reference moves them to the CPU first, :1080-1083):

        score cut `scores > sigma_d` (:337-344) -> drop the 2D box columns (:349-350) -> im_nms with the reference's scalar
        group offset (:354, :592-615) -> guess_heights + im_to_state (:363-364), optionally the two-pass height refinement
        (:366-370) in ONE fused kernel (g3d_im_to_state_refined) -> space_nms (:376-381).

    hg: the drop-in Homography / Homography_Wrapper; cameras: list of camera names indexed by camera_idxs (:361).
    scores[d], labels[d] (integer classes: like the reference, guess_heights then falls back to the "other" height unless
    `heights[d]` is given), boxes[d,20], camera_idxs[d].  n_best is accepted and, as in the reference body, not used.
    Returns (boxes[k,6] float32 states, labels[k], scores[k], camera_idxs[k]); four empty lists when nothing survives the
    cut, as the reference does (:332-333, :346-347).  estimate_ts_bias (:372-374) is a separate call (estimate_ts_bias)."""
    if len(scores) == 0:
        return [], [], [], []
    dev = _exec_device(scores, boxes)
    sc, lb = _to_dev(scores, dev), _to_dev(torch.as_tensor(labels), dev)
    bx, cam = _to_dev(boxes, dev), _to_dev(torch.as_tensor(camera_idxs), dev)
    keepers = torch.nonzero(sc > float(sigma_d)).reshape(-1)          # the one data-dependent size of the function
    if keepers.numel() == 0:
        return [], [], [], []
    sc, lb, cam = sc[keepers], lb[keepers], cam[keepers]
    det = bx[keepers].reshape(-1, 10, 2)[:, :8, :].contiguous()
    if perform_nms:
        idxs = im_nms(det, sc, threshold=phi_nms_im, groups=cam)
        sc, lb, cam, det = sc[idxs], lb[idxs], cam[idxs], det[idxs].contiguous()
    if heights is None:
        h = hg.guess_heights(lb.tolist())                   # integer labels: the "other" height, as in the reference
    else:
        h = _to_dev(torch.as_tensor(heights), dev)[keepers]
        h = h[idxs] if perform_nms else h
    # camera names -> the homography's own camera indices, on the device (a Python list of d names is what dominates the
    # reference's call)
    bank = hg._bank()
    lut = torch.tensor([bank.index[c] for c in cameras], dtype=torch.uint8, device=dev)
    cam_names = lut[cam.long()]
    h = _to_dev(h, dev)
    if refine_height:
        states = hg.im_to_state_refined(det, heights=h, name=cam_names) if hasattr(hg, "im_to_state_refined") else None
        if states is None:
            states = hg.im_to_state(det, heights=h, name=cam_names)
            repro = hg.state_to_im(states, name=cam_names)
            states = hg.im_to_state(det, heights=hg.height_from_template(repro, h, det), name=cam_names)
    else:
        states = hg.im_to_state(det, heights=h, name=cam_names)
    if perform_nms:
        idxs = space_nms(states, sc, threshold=phi_nms_space)
        states, sc, lb, cam = states[idxs], sc[idxs], lb[idxs], cam[idxs]
    return _ret(states, boxes), _ret(lb, boxes), _ret(sc, boxes), _ret(cam, boxes)


class FrameGeometry:
    """One tracker frame's geometry as a single CUDA graph.

    Per frame the reference runs, from Python with a host round trip after every step: match_hungarian's cost matrix
    `1 - md_iou(footprint(pre), footprint(det))` (MC3D_crop_tracker.py:663-689), the road-plane NMS of parse_detections
    (:376-381, space_nms :617-635) and the image-plane NMS of the projected boxes (state_to_im + im_nms, :592-615).  The
    three are independent given the inputs, and each is a handful of launch-latency-bound kernels, so they are captured
    once - forked onto three streams - and replayed with one graph launch per frame: no allocation, no host sync until
    the single 8-byte read of the two keep-list lengths.

    The buffers have a fixed capacity; the number of detections of the current frame lives in a device-side segment
    table that the NMS kernels read (nms.cu, single-segment path), so the same graph serves every frame with
    n_pre <= capacity_pre and n_det <= capacity_det.  Rows beyond the current counts hold stale (finite) data that only
    feed entries of the cost matrix outside the returned view.

    P: float64 [n_cam, 2, 3, 4] projection matrices as `ops.state_to_im` takes them (Homography_Wrapper's two per camera).
    """

    def __init__(self, P, capacity_pre, capacity_det, phi_space=0.1, phi_im=0.3, wrapper=True, device=None, graph=True):
        dev = torch.device(device) if device is not None else (P.device if isinstance(P, torch.Tensor) and P.is_cuda
                                                                else torch.device("cuda", torch.cuda.current_device()))
        if dev.type != "cuda":
            raise ops.Geom3dError("FrameGeometry runs on a CUDA device (no CPU fallback)")
        self.device = dev
        self.P = _to_dev(torch.as_tensor(P), dev).to(torch.float64).contiguous()
        self.cap_pre, self.cap_det = int(capacity_pre), int(capacity_det)
        self.phi_space, self.phi_im, self.wrapper = float(phi_space), float(phi_im), bool(wrapper)
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)      # noqa: E731
        self.pre, self.det = z(self.cap_pre, 6), z(self.cap_det, 6)
        self.pre[:, 2:5] = 1.0
        self.det[:, 2:5] = 1.0                                  # unit cuboids: finite footprints in the unused rows
        self.scores = z(self.cap_det)
        self.cams = z(self.cap_det, dtype=torch.uint8)
        self.seg = z(2, dtype=torch.int32)
        self._seg_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._counts_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._side = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        self._use_graph, self._graph, self._out = bool(graph), None, None

    def _launch(self):
        """the frame's launch sequence on the current stream + two forked side streams"""
        main = torch.cuda.current_stream(self.device)
        s_space, s_im = self._side
        s_space.wait_stream(main)
        s_im.wait_stream(main)
        with torch.cuda.stream(s_space):
            keep_space, cnt_space = ops.nms_segmented(ops.state_footprint(self.det), self.scores, self.seg, self.cap_det,
                                                      self.phi_space)
        with torch.cuda.stream(s_im):
            corners = ops.state_to_im(self.det, self.P, self.cams, wrapper=self.wrapper)
            keep_im, cnt_im = ops.nms_segmented(ops.corners_to_box(corners).to(torch.float32), self.scores, self.seg,
                                                self.cap_det, self.phi_im)
        cost = ops.pairwise_iou(ops.state_footprint(self.pre), ops.state_footprint(self.det), one_minus=True)
        main.wait_stream(s_space)
        main.wait_stream(s_im)
        counts = torch.cat((cnt_space, cnt_im))
        return {"cost": cost, "keep_space": keep_space, "keep_im": keep_im, "corners": corners, "counts": counts}

    def _capture(self):
        cap_stream = torch.cuda.Stream(self.device)
        cap_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cap_stream):
            self._launch()                                      # warm-up: module load, allocator pools
        torch.cuda.current_stream(self.device).wait_stream(cap_stream)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap_stream, capture_error_mode="thread_local"):
            out = self._launch()
        self._graph, self._out = g, out

    def __call__(self, pre_states, det_states, det_scores, det_cams):
        """pre_states[n_pre,>=6], det_states[n_det,>=6], det_scores[n_det], det_cams[n_det] (camera index per detection).
        Returns dict(cost f64[n_pre,n_det] = 1 - IoU, space_keep i64[k1], im_keep i64[k2], corners f64[n_det,8,2]) - views
        of the instance's buffers, valid until the next call."""
        n_pre, n_det = int(pre_states.shape[0]), int(det_states.shape[0])
        if n_pre > self.cap_pre or n_det > self.cap_det:
            raise ValueError(f"frame of {n_pre} x {n_det} exceeds the capacity {self.cap_pre} x {self.cap_det}")
        with torch.cuda.device(self.device):
            self.pre[:n_pre].copy_(torch.as_tensor(pre_states)[:, :6], non_blocking=True)
            self.det[:n_det].copy_(torch.as_tensor(det_states)[:, :6], non_blocking=True)
            self.scores[:n_det].copy_(torch.as_tensor(det_scores), non_blocking=True)
            self.cams[:n_det].copy_(torch.as_tensor(det_cams), non_blocking=True)
            self._seg_host[1] = n_det
            self.seg.copy_(self._seg_host, non_blocking=True)
            if self._use_graph:
                if self._graph is None:
                    self._capture()
                self._graph.replay()
                out = self._out
            else:
                out = self._launch()
            self._counts_host.copy_(out["counts"], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        k1, k2 = int(self._counts_host[0]), int(self._counts_host[1])
        return {"cost": out["cost"][:n_pre, :n_det], "space_keep": out["keep_space"][:k1], "im_keep": out["keep_im"][:k2],
                "corners": out["corners"][:n_det]}
