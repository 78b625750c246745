"""im_to_state / state_to_im timing on 10M states: python tools/time_hg.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops
dev = torch.device("cuda", 0)
g = synth.gen(7)
P, Hm = synth.camera_matrices(18)
Pd, Hd = torch.from_numpy(P).to(dev), torch.from_numpy(Hm).to(dev)
d = 10_000_000
st, cam = synth.vehicle_states(d, g); st, cam = st.to(dev), cam.to(dev)
im = ops.state_to_im(st, Pd, cam, wrapper=True); hts = st[:, 4].contiguous()
def timed(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
t1 = timed(lambda: ops.state_to_im(st, Pd, cam, wrapper=True))
t2 = timed(lambda: ops.im_to_state(im, hts, Hd, cam, wrapper=True))
print(f"state_to_im {t1*1e3:.1f} us ({d*153/t1/1e6:.0f} GB/s alg)   im_to_state {t2*1e3:.1f} us ({d*97/t2/1e6:.0f} GB/s alg)")
