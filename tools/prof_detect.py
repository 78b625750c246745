"""cfg3 inference tail once (for ncu launch lists): python tools/prof_detect.py [B] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops, postprocess
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
g = synth.gen(7)
anc = synth.anchors(1080, 1920).to(dev); A = anc.shape[1]
cls = torch.rand(B, A, 8, device=dev) * 0.04
small = synth.detection_scores(1, A, 8, g)
hot = torch.nonzero(small[0] > 0.04)
for b in range(B):
    shift = (hot[:, 0] + 1237 * b) % A
    cls[b, shift.to(dev), hot[:, 1].to(dev)] = small[0][hot[:, 0], hot[:, 1]].to(dev)
reg3 = torch.randn(B, A, 12, device=dev) * 0.1
reg3[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5], device=dev) + torch.randn(B, A, 4, device=dev) * 0.05
torch.cuda.synchronize()
for it in range(iters):
    t0 = time.perf_counter()
    out = postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05)
    torch.cuda.synchronize()
    print("iter", it, "wall ms", (time.perf_counter() - t0) * 1e3, "detections", out[0].numel(), flush=True)
