"""Profiling driver: runs each kernel family of the hot path a few times at the bench.py shapes (no timing here).
Used under ncu via gpurun (see profiles/README.md); `python profiles/run_kernels.py [loss|decode|nms|homography|all] [B]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from geom3d_b200 import losses_impl, ops, postprocess  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
g = synth.gen(0)
anc = synth.anchors(1080, 1920).to(dev)
A = anc.shape[1]
iters = 3

if what in ("loss", "all"):
    ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
    cls = (torch.rand(B, A, 8, device=dev) * 0.1).requires_grad_(True)
    reg = (torch.randn(B, A, 12, device=dev) * 0.1).requires_grad_(True)
    for _ in range(iters):
        cls.grad = reg.grad = None
        l = losses_impl.focal_loss(cls, reg, anc, ann)[0]
        l.backward(torch.ones(3, device=dev))
    torch.cuda.synchronize()
    del cls, reg
if what in ("decode", "all"):
    reg = torch.randn(B, A, 12, device=dev) * 0.1
    dl = torch.randn(B, A, 4, device=dev) * 0.5
    for _ in range(iters):
        ops.decode3d(anc, reg)
        ops.decode2d(anc, dl, [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2], (1920, 1080))
    torch.cuda.synchronize()
    del reg, dl
if what in ("nms", "all"):
    Bn = min(B, 16)
    cls = synth.detection_scores(Bn, A, 8, g).to(dev)
    reg = torch.randn(Bn, A, 12, device=dev) * 0.1
    reg[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5], device=dev) + torch.randn(Bn, A, 4, device=dev) * 0.05
    boxes = ops.decode3d(anc, reg)
    for _ in range(iters):
        postprocess.detect_per_class(cls, boxes, box_col=16, score_threshold=0.05)
        postprocess.detect_per_class(cls[:1], boxes[:1], box_col=16, ladder_start=1e-25)
    torch.cuda.synchronize()
    del cls, reg, boxes
if what in ("homography", "all"):
    P, H = synth.camera_matrices(18)
    Pd, Hd = torch.from_numpy(P).to(dev), torch.from_numpy(H).to(dev)
    st, cam = synth.vehicle_states(10_000_000, g)
    st, cam = st.to(dev), cam.to(dev)
    for _ in range(iters):
        im = ops.state_to_im(st, Pd, cam, wrapper=True)
        ops.im_to_state(im, st[:, 4].contiguous(), Hd, cam, wrapper=True)
        ops.state_to_im(st[:1_000_000], Pd, None, wrapper=True, all_cams=True)
    torch.cuda.synchronize()
print("done", what)
