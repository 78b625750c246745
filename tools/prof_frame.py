"""cfg5 tracker frame for ncu launch lists: python tools/prof_frame.py [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops, tracker_geometry
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
g = synth.gen(7)
P, Hm = synth.camera_matrices(18)
Pd = torch.from_numpy(P).to(dev)
s5, c5 = synth.vehicle_states(2000, g)
j5 = s5.clone(); j5[:, :2] += torch.randn(2000, 2, generator=g) * torch.tensor([3.0, 0.5])
s5, j5, c5 = s5.to(dev), j5.to(dev), c5.to(dev)
sc5 = torch.rand(2000, device=dev)
torch.cuda.synchronize()
for it in range(iters):
    t0 = time.perf_counter()
    cost = tracker_geometry.association_cost(s5, j5)
    k1 = tracker_geometry.space_nms(s5, sc5, 0.1)
    corners = ops.state_to_im(s5, Pd, c5, wrapper=True)
    k2 = tracker_geometry.im_nms(corners, sc5, 0.3)
    torch.cuda.synchronize()
    print("iter", it, "wall us", (time.perf_counter() - t0) * 1e6, k1.numel(), k2.numel(), flush=True)
