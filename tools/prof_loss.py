"""One full-size training-loss step (BASELINE configs[1]) for ncu: `python tools/prof_loss.py [B] [steps]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import synth  # noqa: E402
from geom3d_b200 import losses_impl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
g = synth.gen(100)
from geom3d_b200.anchors_impl import Anchors  # noqa: E402
anc = Anchors()(torch.zeros(1, 3, 1080, 1920, device=dev))       # tagged pyramid table -> GT-centric assignment
A = anc.shape[1]
ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
torch.manual_seed(100)
cls = (torch.rand(B, A, 8, device=dev) * 0.1).requires_grad_(True)
reg = (torch.randn(B, A, 12, device=dev) * 0.1).requires_grad_(True)
ones = torch.ones(3, device=dev)
for _ in range(steps):
    cls.grad = None
    reg.grad = None
    losses = losses_impl.focal_loss(cls, reg, anc, ann)[0]
    losses.backward(ones)
torch.cuda.synchronize()
print("done loss", [float(x) for x in losses])
