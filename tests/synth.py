"""Seeded synthetic inputs shared by the golden-vector generator, the parity tests and bench.py.
Everything is generated on the CPU with an explicit torch.Generator, so the same call gives the same tensors here,
on the GPU box and inside bench.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from geom3d_b200.anchors_impl import anchors_for_image  # noqa: E402  (host-side, numpy only)

CAMERAS = ["p1c1", "p1c2", "p1c3", "p1c4", "p1c5", "p1c6", "p2c1", "p2c2", "p2c3", "p2c4", "p2c5", "p2c6",
           "p3c1", "p3c2", "p3c3", "p3c4", "p3c5", "p3c6"]
# p1c1 projection recovered from the reference's 3D_tracking_results.csv (SURVEY.md §8d cfg 4)
P_P1C1 = np.array([[-0.42364, 4.67915, 0.34719, 425.41],
                   [0.28708, -0.11913, 3.41593, 282.73],
                   [-1.36398e-3, 4.32862e-4, 3.17938e-4, 1.0]], dtype=np.float64)


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def anchors(height, width):
    return torch.from_numpy(anchors_for_image(height, width)).unsqueeze(0)  # [1,A,4]


def gt_annotations_3d(B, G, height, width, g, n_pad=0, empty_images=(), num_classes=8, box_lo=30.0, box_hi=150.0,
                      box_hi_h=110.0, size_scale=None):
    """[B, G + n_pad, 27]: 16 corner coords (bottom face = box corners, top face shifted by (+10,-15)), enclosing 2D box,
    class, 6 zeros; the last n_pad rows of every image and all rows of `empty_images` are padding (-1)."""
    m = min(80.0, 0.25 * min(height, width))
    cx = m + torch.rand(B, G, generator=g) * (width - 2 * m)
    cy = m + torch.rand(B, G, generator=g) * (height - 2 * m)
    scale = min(1.0, min(height, width) / 540.0) if size_scale is None else size_scale
    w = (box_lo + torch.rand(B, G, generator=g) * (box_hi - box_lo)) * scale
    h = (box_lo + torch.rand(B, G, generator=g) * (box_hi_h - box_lo)) * scale
    x0, x1, y0, y1 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
    bottom = [x0, y1, x1, y1, x0, y0 + 0.5 * h, x1, y0 + 0.5 * h]          # fbl fbr bbl bbr (x,y)
    top = [v + (10.0 if i % 2 == 0 else -15.0) for i, v in enumerate(bottom)]
    corners = torch.stack(bottom + top, dim=2)                             # [B,G,16]
    xs, ys = corners[..., 0::2], corners[..., 1::2]
    box = torch.stack((xs.min(-1).values, ys.min(-1).values, xs.max(-1).values, ys.max(-1).values), dim=2)
    cls = torch.randint(0, num_classes, (B, G, 1), generator=g).float()
    ann = torch.cat((corners, box, cls, torch.zeros(B, G, 6)), dim=2)
    if n_pad:
        ann = torch.cat((ann, -torch.ones(B, n_pad, 27)), dim=1)
    for j in empty_images:
        ann[j] = -1.0
    return ann.contiguous()


def gt_annotations_2d(B, G, height, width, g, n_pad=0, empty_images=(), num_classes=8, **kw):
    a3 = gt_annotations_3d(B, G, height, width, g, n_pad, empty_images, num_classes, **kw)
    return torch.cat((a3[..., 16:20], a3[..., 20:21]), dim=2).contiguous()


def head_outputs(B, A, C, R, g, positives_hint=None):
    """classification probabilities U(0,1)*0.1 (near the 0.01 prior) and regressions N(0,0.1)."""
    cls = torch.rand(B, A, C, generator=g) * 0.1
    reg = torch.randn(B, A, R, generator=g) * 0.1
    return cls.contiguous(), reg.contiguous()


def detection_scores(B, A, C, g, objects=200, per_object=25, lo=0.05, hi=1.0, background=0.04):
    """cfg 3: per image `objects` clusters of `per_object` consecutive anchors scoring U(lo,hi) in one class; the rest
    U(0, background).  Returns scores [B,A,C]."""
    s = torch.rand(B, A, C, generator=g) * background
    for b in range(B):
        starts = torch.randint(0, max(A - per_object, 1), (objects,), generator=g)
        classes = torch.randint(0, C, (objects,), generator=g)
        for st, c in zip(starts.tolist(), classes.tolist()):
            s[b, st:st + per_object, c] = lo + torch.rand(per_object, generator=g) * (hi - lo)
    return s.contiguous()


def clustered_boxes(N, g, objects=None, extent=1000.0, jitter=6.0):
    """NMS test boxes: `objects` cluster centres, jittered copies, scores U(0,1)."""
    objects = max(N // 25, 1) if objects is None else objects
    ctr = torch.rand(objects, 2, generator=g) * extent
    wh = 20 + torch.rand(objects, 2, generator=g) * 80
    which = torch.randint(0, objects, (N,), generator=g)
    c = ctr[which] + torch.randn(N, 2, generator=g) * jitter
    s = wh[which] * (1 + 0.1 * torch.randn(N, 2, generator=g))
    boxes = torch.cat((c - s / 2, c + s / 2), dim=1).contiguous()
    scores = torch.rand(N, generator=g)
    return boxes, scores


def camera_matrices(n_cams=18):
    """Synthetic per-camera P (3x4) and H (3x3 image->road plane): the p1c1 projection translated along the roadway;
    the second ("WB") correspondence of each camera is a slightly perturbed copy.  Returns P[n,2,3,4], H[n,2,3,3]."""
    P = np.zeros((n_cams, 2, 3, 4))
    H = np.zeros((n_cams, 2, 3, 3))
    rng = np.random.RandomState(7)
    for i in range(n_cams):
        shift = np.eye(4)
        shift[0, 3] = -(i * 110.0)                      # camera i looks at x in [110 i, 110 i + ...]
        for j in range(2):
            Pi = P_P1C1 @ shift
            if j == 1:
                Pi = Pi * (1.0 + 1e-3 * rng.randn(3, 4))
            Pi = Pi / Pi[2, 3]
            P[i, j] = Pi
            H[i, j] = np.linalg.inv(Pi[:, [0, 1, 3]])
    return P, H


def vehicle_states(d, g, n_cams=18):
    """[d,6] float32 states (x,y,l,w,h,dir) over the 18-camera range + a camera index per state (uint8)."""
    cam = torch.randint(0, n_cams, (d,), generator=g)
    x = cam.float() * 110.0 + 100.0 + torch.rand(d, generator=g) * 400.0
    y = torch.rand(d, generator=g) * 120.0
    l = 10 + torch.rand(d, generator=g) * 50
    w = 5 + torch.rand(d, generator=g) * 4
    h = 4 + torch.rand(d, generator=g) * 9
    dr = torch.where(torch.rand(d, generator=g) < 0.5, -torch.ones(d), torch.ones(d))
    return torch.stack((x, y, l, w, h, dr), dim=1).contiguous(), cam.to(torch.uint8)


TINY = dict(box_lo=26.0, box_hi=46.0, box_hi_h=44.0, size_scale=1.0)   # GT sizes that match level-3 anchors in 64..128 px images


def dense_detection_inputs(seed, height, width, C=8):
    """A frame with > 10000 candidates per class, so the adaptive ladder has to climb (3D model.py:368-374)."""
    g = gen(seed)
    A = anchors(height, width).shape[1]
    cls = torch.rand(1, A, C, generator=g) * 0.3
    reg = torch.randn(1, A, 12, generator=g) * 0.1
    reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5]) + torch.randn(1, A, 4, generator=g) * 0.05
    return cls.contiguous(), reg.contiguous()


def digest(t):
    """sha256 of a tensor's raw little-endian bytes (bit-exact comparison of large outputs without storing them)."""
    import hashlib
    a = t.detach().cpu().contiguous().numpy()
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8).copy()
