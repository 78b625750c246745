"""Cycles per phase of tail_short_kernel (detect_tail.cu) at bench.py's config-3 workload.  Needs a library built with
G3D_NVCC_FLAGS=-DG3D_TAIL_CLOCKS python 3d-playground_b200/build.py --force (thread 0 of every CTA then leaves clock64()
stamps at the phase boundaries in the unused end of its segment's keep list).  python tools/tail_phases.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
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
thr = torch.full((B * 8,), 0.05, dtype=torch.float32, device=dev)
for _ in range(3):
    t = ops.detect_tail(cls, B, 8, A, A * 8, thr, 16384, anc, reg3, 0.5, short=True)
torch.cuda.synchronize()
clk = t["keep"].view(B * 8, 16384)[:, -16:-7].cpu().double()
d = (clk[:, 1:] - clk[:, :-1])
names = ["1 keys+sort", "2 tables/decode", "2b binning", "3a windows", "3b cells", "3c wide", "4 fixed point", "5 keep"]
print("B", B, "mean cycles per phase (x1000):", {n: round(float(v) / 1e3, 1) for n, v in zip(names, d.mean(0))})
print("max:", {n: round(float(v) / 1e3, 1) for n, v in zip(names, d.max(0).values)})
print("total mean", float((clk[:, -1] - clk[:, 0]).mean()) / 1e3, "max", float((clk[:, -1] - clk[:, 0]).max()) / 1e3)
