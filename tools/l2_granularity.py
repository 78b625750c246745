"""Experiment: does cudaLimitMaxL2FetchGranularity change what im_to_state (f64 input: 64 needed bytes of every 128-byte
object) pulls from HBM?  python tools/l2_granularity.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from geom3d_b200 import ops  # noqa: E402

rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x05      # cudaLimitMaxL2FetchGranularity
dev = torch.device("cuda:0")
P, Hm = synth.camera_matrices(18)
Pd, Hd = torch.from_numpy(P).to(dev), torch.from_numpy(Hm).to(dev)
d = 10_000_000
st, cam = synth.vehicle_states(d, synth.gen(7))
st, cam = st.to(dev), cam.to(dev)
im = ops.state_to_im(st, Pd, cam, wrapper=True)
hts = st[:, 4].contiguous()


def t():
    for _ in range(3):
        ops.im_to_state(im, hts, Hd, cam, wrapper=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.im_to_state(im, hts, Hd, cam, wrapper=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


v = ctypes.c_size_t(0)
print("get", rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT), "default granularity", v.value, "im_to_state ms", t())
for g in (32, 64, 128):
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(g))
    rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    print("set", g, "rc", rc, "now", v.value, "im_to_state ms", t())
