"""Per-launch times of the loss forward for a few settings: python tools/time_fused.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from geom3d_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
g = synth.gen(100)
anc = synth.anchors(1080, 1920).to(dev); A = anc.shape[1]
ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
torch.manual_seed(100)
cls = torch.rand(B, A, 8, device=dev) * 0.1
reg = torch.randn(B, A, 12, device=dev) * 0.1
for fused, mix in (("0", "0"), ("1", "0"), ("1", "2"), ("1", "3"), ("1", "5")):
    os.environ["G3D_LOSS_FUSED"] = fused; os.environ["G3D_FUSED_MIX"] = mix
    ts = []
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in ev: e.record()
        out = ops.focal_loss_forward(cls, reg, anc, ann, grad_cls_expected=1.0, trace_events=ev)
        torch.cuda.synchronize()
        ts.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)])
    t = ts[-1]
    print(f"B={B} fused={fused} mix={mix}: launches (us) {t[0]:.1f} {t[1]:.1f} {t[2]:.1f}  sum {sum(t):.1f}  losses {[round(float(x),6) for x in out['losses']]}", flush=True)
