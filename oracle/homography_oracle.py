"""ORACLE (test infrastructure): state <-> space <-> image transforms of homography.py, on plain matrices.

    state_to_space        homography.py:305-320
    space_to_state        homography.py:274-303
    space_to_im           homography.py:438-476 (P given per object or one for all)
    im_to_space           homography.py:388-435
    state_to_im / im_to_state   homography.py:479-500
    height_from_template  homography.py:519-551
    wrapper_*             Homography_Wrapper, homography.py:840-862 (second correspondence where y > 60)
Matrices are passed explicitly (P[3,4] / H[3,3] float64 numpy or tensors, or per-object stacks [d,3,4] / [d,3,3]).
Pinned by tests/golden/homography_*.npz (reference outputs) and the reference's own CSV rows.
"""
import torch


def _t(m):
    return torch.as_tensor(m, dtype=torch.float64)


def state_to_space(states):
    d = states.shape[0]
    x, y, l, w, h, dr = (states[:, i] for i in range(6))
    out = torch.zeros(d, 8, 3)
    front = x + dr * l
    lo = y - dr * w / 2.0
    hi = y + dr * w / 2.0
    for k in range(8):
        out[:, k, 0] = x if (k & 2) else front
        out[:, k, 1] = hi if (k & 1) else lo
        if k & 4:
            out[:, k, 2] = -h
    return out


def space_to_state(points):
    d = points.shape[0]
    out = torch.zeros(d, 6)
    p = points
    out[:, 0] = (p[:, 2, 0] + p[:, 3, 0]) / 2.0
    out[:, 1] = (p[:, 0, 1] + p[:, 1, 1] + p[:, 2, 1] + p[:, 3, 1]) / 4.0
    signed = ((p[:, 0, 0] + p[:, 1, 0]) - (p[:, 2, 0] + p[:, 3, 0])) / 2.0
    out[:, 2] = torch.abs(signed)
    out[:, 3] = torch.abs(((p[:, 0, 1] + p[:, 2, 1]) - (p[:, 1, 1] + p[:, 3, 1])) / 2.0)
    out[:, 4] = torch.mean(torch.abs(p[:, 0:4, 2] - p[:, 4:8, 2]), dim=1)
    out[:, 5] = torch.sign(signed)
    return out


def space_to_im(points, P):
    """points[d,m,3]; P [3,4] or [d,3,4] -> float64 [d,m,2]"""
    d, m = points.shape[0], points.shape[1]
    P = _t(P)
    pts = torch.cat((points.reshape(-1, 3).double(), torch.ones(d * m, 1, dtype=torch.float64)), dim=1)
    if P.dim() == 3:
        Pe = P.unsqueeze(1).expand(d, m, 3, 4).reshape(-1, 3, 4)
        proj = torch.bmm(Pe, pts.unsqueeze(2)).squeeze(2)
    else:
        proj = torch.matmul(P, pts.t()).t()
    return torch.stack((proj[:, 0] / proj[:, 2], proj[:, 1] / proj[:, 2]), dim=1).reshape(d, m, 2)


def im_to_space(points, H, heights):
    """points[d,8,2]; H [3,3] or [d,3,3]; heights[d] -> float64 [d,8,3]"""
    d = points.shape[0]
    H = _t(H)
    pts = torch.cat((points.reshape(-1, 2).double(), torch.ones(d * 8, 1, dtype=torch.float64)), dim=1)
    if H.dim() == 3:
        He = H.unsqueeze(1).expand(d, 8, 3, 3).reshape(-1, 3, 3)
        q = torch.bmm(He, pts.unsqueeze(2)).squeeze(2)
    else:
        q = torch.matmul(pts, H.t())
    xy = torch.stack((q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]), dim=1).reshape(d, 8, 2)
    out = torch.cat((xy, torch.zeros(d, 8, 1, dtype=torch.float64)), dim=2)
    out[:, 4:8, 2] = heights.double().unsqueeze(1)
    return out


def state_to_im(states, P):
    return space_to_im(state_to_space(states), P)


def im_to_state(points, H, heights):
    return space_to_state(im_to_space(points, H, heights))


def height_from_template(template_boxes, template_space_heights, boxes):
    def im_height(b):
        top, bottom = torch.mean(b[:, 4:8, :], dim=1), torch.mean(b[:, 0:4, :], dim=1)
        return torch.sum(torch.sqrt(torch.pow(top - bottom, 2)), dim=1)
    ratio = im_height(template_boxes) / template_space_heights
    return im_height(boxes) / ratio


def _select(first, second, use_second):
    out = first.clone()
    out[use_second] = second[use_second]
    return out


def wrapper_space_to_im(points, P1, P2):
    return _select(space_to_im(points, P1), space_to_im(points, P2), points[:, 0, 1] > 60)


def wrapper_state_to_im(states, P1, P2):
    return wrapper_space_to_im(state_to_space(states), P1, P2)


def wrapper_im_to_space(points, H1, H2, heights):
    a = im_to_space(points, H1, heights)
    return _select(a, im_to_space(points, H2, heights), a[:, 0, 1] > 60)


def wrapper_im_to_state(points, H1, H2, heights):
    return space_to_state(wrapper_im_to_space(points, H1, H2, heights))


def refined_im_to_state(points, H1, H2, P1, P2, heights, wrapper=True):
    """MC3D_crop_tracker.py:364-370: two-pass height refinement.  Returns (state[d,6] f32, refined heights f64)."""
    if wrapper:
        s0 = wrapper_im_to_state(points, H1, H2, heights)
        repro = wrapper_state_to_im(s0, P1, P2)
    else:
        s0 = im_to_state(points, H1, heights)
        repro = state_to_im(s0, P1)
    h1 = height_from_template(repro, heights, points)
    s1 = wrapper_im_to_state(points, H1, H2, h1) if wrapper else im_to_state(points, H1, h1)
    return s1, h1
