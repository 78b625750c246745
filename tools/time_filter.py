"""compact8_kernel (score filter) alone at bench.py's config-3 data: python tools/time_filter.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops
B = 64
dev = torch.device("cuda", 0)
g = synth.gen(7)
A = synth.anchors(1080, 1920).shape[1]
cls = torch.rand(B, A, 8, device=dev) * 0.04
small = synth.detection_scores(1, A, 8, g)
hot = torch.nonzero(small[0] > 0.04)
for b in range(B):
    shift = (hot[:, 0] + 1237 * b) % A
    cls[b, shift.to(dev), hot[:, 1].to(dev)] = small[0][hot[:, 0], hot[:, 1]].to(dev)
thr = torch.full((B * 8,), 0.05, dtype=torch.float32, device=dev)
for _ in range(3):
    ops.filter_compact(cls, B, 8, A, A * 8, thr, 16384)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    ops.filter_compact(cls, B, 8, A, A * 8, thr, 16384)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 30
print(f"filter_compact (memset + compact8_kernel): {t * 1e3:.1f} us = {B * A * 32 / t / 1e6:.0f} GB/s = {B * A * 32 / t / 1e6 / 6547.8:.3f} of the measured HBM peak")
