"""cfg5 tracker frame: separate drop-in calls vs the one-graph FrameGeometry (graph and eager), CUDA-event timed.
python tools/time_frame.py [n_objects]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops, tracker_geometry
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda", 0)
g = synth.gen(7)
P, Hm = synth.camera_matrices(18)
Pd = torch.from_numpy(P).to(dev)
s5, c5 = synth.vehicle_states(n, g)
j5 = s5.clone(); j5[:, :2] += torch.randn(n, 2, generator=g) * torch.tensor([3.0, 0.5])
s5, j5, c5 = s5.to(dev), j5.to(dev), c5.to(dev)
sc5 = torch.rand(n, device=dev)


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def separate():
    cost = tracker_geometry.association_cost(s5, j5)
    k1 = tracker_geometry.space_nms(j5, sc5, 0.1)
    corners = ops.state_to_im(j5, Pd, c5, wrapper=True)
    k2 = tracker_geometry.im_nms(corners, sc5, 0.3)
    return cost, k1, k2


fg = tracker_geometry.FrameGeometry(Pd, n, n)
fe = tracker_geometry.FrameGeometry(Pd, n, n, graph=False)
a, b = separate(), fg(s5, j5, sc5, c5)
assert torch.equal(a[0], b["cost"]) and torch.equal(a[1], b["space_keep"]) and torch.equal(a[2], b["im_keep"])
print(f"n={n}  separate calls {timed(separate):.1f} us   FrameGeometry eager {timed(lambda: fe(s5, j5, sc5, c5)):.1f} us   "
      f"FrameGeometry graph {timed(lambda: fg(s5, j5, sc5, c5)):.1f} us")
fg._graph.replay(); torch.cuda.synchronize()
print(f"graph replay alone (no input copies, no length read) {timed(lambda: fg._graph.replay()):.1f} us")
# parts, each alone on one stream
fp = ops.state_footprint(s5)
print(f"parts: cost {timed(lambda: tracker_geometry.association_cost(s5, j5)):.1f}  "
      f"space nms (device-side) {timed(lambda: ops.nms_segmented(ops.state_footprint(s5), sc5, fg.seg, n, 0.1)):.1f}  "
      f"state_to_im {timed(lambda: ops.state_to_im(s5, Pd, c5, wrapper=True)):.1f}  "
      f"im nms {timed(lambda: ops.nms_segmented(ops.corners_to_box(b['corners']).float(), sc5, fg.seg, n, 0.3)):.1f} us")
