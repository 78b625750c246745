"""cfg3 detection tail (bench.py's workload): event-timed short path vs general chain.  python tools/prof_tail.py [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops, postprocess
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
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


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


key = (dev.index, B * 8)
res = {}
for name, general in (("short", 0), ("general", 1 << 30)):
    postprocess._TAIL_GENERAL[key] = general
    out = postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05)
    res[name] = out
    ms = timed(lambda: postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05), iters)
    print(f"{name}: {ms:.4f} ms / {B} images = {B / ms * 1e3:.0f} img/s, detections {out[0].numel()}", flush=True)
    # the device part alone (no host read, no assembly): the library call back to back
    thr = torch.full((B * 8,), 0.05, dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.detect_tail(cls, B, 8, A, A * 8, thr, 16384, anc, reg3, 0.5, short=(general == 0)), iters)
    print(f"{name}: library call alone {ms:.4f} ms", flush=True)
assert all(torch.equal(a, b) for a, b in zip(res["short"], res["general"]))
t = ops.detect_tail(cls, B, 8, A, A * 8, thr, 16384, anc, reg3, 0.5, short=True)
print("max count", int(t["count"].max()), "mean", float(t["count"].float().mean()), "left", int(t["summary"][2]))

# ---- host time of one call (the GPU idles while the host prepares the next one)
import time
postprocess._TAIL_GENERAL[key] = 0
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05)
print(f"host wall per call {(time.perf_counter() - t0) / iters * 1e6:.1f} us")
