"""A few loss steps at BASELINE configs[1] for ncu (python tools/prof_loss.py [B] [steps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from geom3d_b200 import ops  # noqa: E402
from geom3d_b200.anchors_impl import Anchors  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
anc = Anchors()(torch.zeros(1, 3, 1080, 1920, device=dev))
A = anc.shape[1]
ann = synth.gt_annotations_3d(B, 200, 1080, 1920, synth.gen(100)).to(dev)
torch.manual_seed(100)
cls = torch.rand(B, A, 8, device=dev) * 0.1
reg = torch.randn(B, A, 12, device=dev) * 0.1
ones = torch.ones(3, device=dev)
for _ in range(steps):
    f = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0)
    ops.focal_loss_backward(f, ones)
torch.cuda.synchronize()
print("losses", f["losses"].tolist())
