import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, synth
from geom3d_b200 import ops
B = 32
dev = torch.device("cuda", 0)
g = synth.gen(100)
anc = synth.anchors(1080, 1920).to(dev); A = anc.shape[1]
ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
torch.manual_seed(100)
cls = torch.rand(B, A, 8, device=dev) * 0.1
reg = torch.randn(B, A, 12, device=dev) * 0.1
for it in range(2):
    out = ops.focal_loss_forward(cls, reg, anc, ann, grad_cls_expected=1.0)
torch.cuda.synchronize(); print("ok")
