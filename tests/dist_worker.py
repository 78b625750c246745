"""Worker of tests/test_gpu_dist.py (multi-GPU parity of dist.sharded_focal_loss over real NCCL):
    python -m torch.distributed.run --nproc-per-node N tests/dist_worker.py
Every rank holds the FULL batch (same seed), computes the loss of its shard through sharded_focal_loss and compares the
global losses and its local gradients with the single-GPU loss of the full batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import synth
from geom3d_b200 import dist as gdist, losses_impl

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B = 2 * world + 3                       # uneven shards on purpose (every rank gets at least 2 images)
g = synth.gen(5)
from geom3d_b200.anchors_impl import Anchors
anc = Anchors()(torch.zeros(1, 3, 200, 168, device=dev)); A = anc.shape[1]     # the tagged table: GT-centric assignment
ann = synth.gt_annotations_3d(B, 9, 200, 168, g, n_pad=1, empty_images=(1,), **synth.TINY).to(dev)
cls, reg = synth.head_outputs(B, A, 8, 12, g)
cf, rf = cls.to(dev).requires_grad_(True), reg.to(dev).requires_grad_(True)
full = losses_impl.focal_loss(cf, rf, anc, ann)[0]
w = torch.tensor([1.0, 0.7, 1.3], device=dev)
(full * w).sum().backward()
lo, hi = gdist.shard_range(B, rank, world)
cl, rl = cls[lo:hi].to(dev).requires_grad_(True), reg[lo:hi].to(dev).requires_grad_(True)
got = gdist.sharded_focal_loss(cl, rl, anc, ann[lo:hi].contiguous())
(got * w).sum().backward()
def rel(a, b):
    """max |a-b| / max(|b|, mean |b| over the non-zero entries): the metric of tests/conftest.assert_close_rel"""
    a, b = a.detach().double(), b.detach().double()
    nz = b[b != 0]
    scale = float(nz.abs().mean()) if nz.numel() else 1.0
    return float(((a - b).abs() / torch.maximum(b.abs(), torch.full_like(b, scale))).max()) if a.numel() else 0.0
errs = [rel(got, full.detach()), rel(cl.grad, cf.grad[lo:hi]), rel(rl.grad, rf.grad[lo:hi])]
ok = all(e < 1e-5 for e in errs)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
print(f"check_dist rank {rank}/{world} shard [{lo},{hi}) errs (losses, dcls, dreg) = {errs}", flush=True)
path = "peer-memory exchange" if gdist._PeerExchange.get(None, dev) is not None else "NCCL all-gather"
# a second and third step through the same exchange buffers (epochs / parities advance), gradients unchanged
for _ in range(2):
    cl.grad = None; rl.grad = None
    again = gdist.sharded_focal_loss(cl, rl, anc, ann[lo:hi].contiguous())
    (again * w).sum().backward()
    ok2 = torch.equal(again, got) and rel(cl.grad, cf.grad[lo:hi]) < 1e-5
    flag2 = torch.tensor([1 if ok2 else 0], device=dev); dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
    flag = torch.minimum(flag, flag2)
if rank == 0: print("check_dist", "OK" if int(flag) else "FAILED", "world", world, "via", path, [float(x) for x in got.detach()])
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
