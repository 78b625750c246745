"""Per-kernel timing of the loss step at BASELINE configs[1] (B = 32, 1080p, 200 GT / image) for several placements of
the dreg zero fill (g3d_set_tuning).  Run on the GPU box:  python tools/sweep_loss.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from geom3d_b200 import ops  # noqa: E402
from geom3d_b200.anchors_impl import Anchors  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    anc = Anchors()(torch.zeros(1, 3, 1080, 1920, device=dev))
    A = anc.shape[1]
    g = synth.gen(100)
    ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
    torch.manual_seed(100)
    cls = torch.rand(B, A, 8, device=dev) * 0.1
    reg = torch.randn(B, A, 12, device=dev) * 0.1
    ones = torch.ones(3, device=dev)
    names = ["prologue", "assign", "resolve", "stream", "positives", "bwd"]

    def run(label, table, iters=10):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(iters)]
        for e in sum(ev, []):
            e.record()
        for _ in range(3):
            f = ops.focal_loss_forward(cls, reg, table, ann, want_assign=False, grad_expected=1.0)
            ops.focal_loss_backward(f, ones)
        torch.cuda.synchronize()
        for i in range(iters):
            f = ops.focal_loss_forward(cls, reg, table, ann, want_assign=False, grad_expected=1.0, trace_events=ev[i][:6])
            ops.focal_loss_backward(f, ones)
            ev[i][6].record()
        torch.cuda.synchronize()
        ms = [sum(e[k].elapsed_time(e[k + 1]) for e in ev) / iters for k in range(6)]
        total = sum(e[0].elapsed_time(e[6]) for e in ev) / iters
        print(f"{label:34s} total {total*1e3:7.1f} us | " + " ".join(f"{n} {m*1e3:6.1f}" for n, m in zip(names, ms)), flush=True)
        return f

    ref = run("anchor-centric (untagged table)", anc.clone())
    f = run("gt-centric", anc)
    assert torch.equal(f["dreg"], ref["dreg"]) and torch.equal(f["dcls"], ref["dcls"]), "gradients differ"
    assert torch.equal(f["losses"], ref["losses"])
    # untraced eager steps (K4 as a programmatic dependent launch behind K3) with and without PDL
    for pdl in (1, 0):
        ops.set_tuning("pdl", pdl)
        for _ in range(3):
            f = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0)
            ops.focal_loss_backward(f, ones)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0)
            ops.focal_loss_backward(f, ones)
        e1.record()
        torch.cuda.synchronize()
        print(f"{'eager, untraced, pdl=' + str(pdl):34s} total {e0.elapsed_time(e1)*50:7.1f} us", flush=True)
        assert torch.equal(f["dreg"], ref["dreg"]) and torch.equal(f["losses"], ref["losses"])
    ops.set_tuning("pdl", 1)
    # the same step as one CUDA graph
    f = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        f = ops.focal_loss_forward(cls, reg, anc, ann, want_assign=False, grad_expected=1.0)
        ops.focal_loss_backward(f, ones)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{'graph replay of the step':34s} total {e0.elapsed_time(e1)*50:7.1f} us", flush=True)
    assert torch.equal(f["dreg"], ref["dreg"]) and torch.equal(f["losses"], ref["losses"])
    # forward only
    for label, table in (("fwd-only gt-centric", anc), ("fwd-only anchor-centric", anc.clone())):
        for _ in range(3):
            ops.focal_loss_forward(cls, reg, table, ann, want_assign=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.focal_loss_forward(cls, reg, table, ann, want_assign=False)
        e1.record()
        torch.cuda.synchronize()
        print(f"{label:34s} total {e0.elapsed_time(e1)*100:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
