"""Randomised comparison of the short-segment detection tail with the general chain (no sanitizer on the GPU pool: this is
the race / bounds check of tail_short_kernel).  python tools/stress_tail.py [rounds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from geom3d_b200 import ops

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = synth.gen(4242)
mean, std = np.zeros(4, dtype=np.float32), np.ones(4, dtype=np.float32)
taken = left_total = 0
for it in range(rounds):
    C = int(torch.randint(1, 9, (1,), generator=g))
    B = int(torch.randint(1, 4, (1,), generator=g))
    per = int(torch.randint(1, 1025, (1,), generator=g))                  # candidates per segment, roughly
    N = min(per * C, 6000)
    kind = it % 4
    if kind == 0:
        boxes = synth.clustered_boxes(N, g, objects=max(N // int(torch.randint(2, 60, (1,), generator=g)), 1))[0]
    elif kind == 1:
        c = torch.rand(N, 2, generator=g) * 10 ** float(torch.rand(1, generator=g) * 4)
        wh = 10 ** (torch.rand(N, 2, generator=g) * 4 - 2)
        boxes = torch.cat((c - wh / 2, c + wh / 2), dim=1)
    elif kind == 2:
        boxes = synth.clustered_boxes(N, g, jitter=float(torch.rand(1, generator=g) * 30))[0]
        boxes[::13, 2] = boxes[::13, 0]
        boxes[5::17, 1] = float("nan")
    else:
        boxes = synth.clustered_boxes(N, g, extent=200.0, jitter=20.0)[0]   # dense: many suppressing pairs
    iou = float(torch.rand(1, generator=g) * 0.9)
    cls = torch.rand(B, N, C, generator=g)
    hot = torch.randint(0, C, (B, N), generator=g)
    cls.scatter_add_(2, hot.unsqueeze(2), torch.ones(B, N, 1))
    cls = cls.cuda().contiguous()
    anc = boxes.float().reshape(1, N, 4).contiguous().cuda()
    reg = torch.zeros(B, N, 4, device="cuda")
    thr = torch.full((B * C,), 1.0, dtype=torch.float32, device="cuda")
    cap = min(16384, N)
    outs = []
    for short in (True, False):
        t = ops.detect_tail(cls, B, C, N, N * C, thr, cap, anc, reg, iou, mean, std, None, short=short)
        s = t["summary"].cpu()
        out = ops.assemble_detections(t["keep"], t["keep_count"], t["seg_offsets"], t["cand_scores"], t["cand_src"], B, C, N,
                                      anc, reg, mean, std, None, out_offsets=t["out_offsets"], K=int(s[0]))
        outs.append((out, int(s[2])))
    left = outs[0][1]
    left_total += left
    if left == 0:
        taken += 1
        for a_, b_ in zip(outs[0][0], outs[1][0]):
            if a_.is_floating_point():
                a_, b_ = a_.view(torch.int32), b_.view(torch.int32)
            assert a_.shape == b_.shape and torch.equal(a_, b_), (it, kind, N, C, B, iou)
print(f"{rounds} random tails: {taken} entirely on the short path and identical to the general chain, {left_total} segments left to it in the others")
