"""Kalman predict / update at bench.py's size (1 M objects, 6 states / 5 measurements): python tools/time_kf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geom3d_b200 import ops
dev = torch.device("cuda", 0)
nk, Sk, Mk = 1_000_000, 6, 5
Fk = torch.eye(Sk)
Hk = torch.zeros(Mk, Sk); Hk[:Mk, :Mk] = torch.eye(Mk)
Qk, Rk = torch.eye(Sk) * 0.5, torch.eye(Mk) * 0.8
Xk = torch.randn(nk, Sk, device=dev) * 20
Ck = torch.randn(nk, Sk, Sk, device=dev)
Pk = (Ck @ Ck.transpose(1, 2) + torch.eye(Sk, device=dev) * 3.0).contiguous()
Dk = torch.ones(nk, device=dev)
Tk = torch.zeros(nk, dtype=torch.float64, device=dev)
dtk = torch.full((nk,), 1 / 30.0, dtype=torch.float64, device=dev)
rowsk = torch.arange(nk, device=dev)
zk = torch.randn(nk, Mk, dtype=torch.float64, device=dev) * 20


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tp = timed(lambda: ops.kf_predict_(Xk, Pk, Dk, dtk, Fk, Qk, 1 / 30.0, Tk))
tu = timed(lambda: ops.kf_update_(Xk, Pk, rowsk, zk, Hk, Rk, None))
pb = nk * (2 * (Sk + Sk * Sk) * 4 + 4 + 8 + 16)
ub = nk * (2 * (Sk + Sk * Sk) * 4 + 8 + Mk * 8)
print(f"predict {tp * 1e3:.1f} us ({pb / tp / 1e6:.0f} GB/s), update {tu * 1e3:.1f} us ({ub / tu / 1e6:.0f} GB/s = {ub / tu / 1e6 / 6547.8:.3f} of peak)")
