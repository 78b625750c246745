"""2-rank diagnostic: where does the time of one sharded loss step go?  torchrun --nproc-per-node 2 tools/diag_dist.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import synth
from geom3d_b200 import dist as gdist, losses_impl, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B = 32
g = synth.gen(100 + rank)
anc = synth.anchors(1080, 1920).to(dev); A = anc.shape[1]
ann = synth.gt_annotations_3d(B, 200, 1080, 1920, g).to(dev)
torch.manual_seed(100 + rank)
cls = (torch.rand(B, A, 8, device=dev) * 0.1).requires_grad_(True)
reg = (torch.randn(B, A, 12, device=dev) * 0.1).requires_grad_(True)
ones = torch.ones(3, device=dev)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

x = torch.zeros(5, dtype=torch.float64, device=dev)
def ag():
    out = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(out, x)
def ag2():
    out = torch.empty(world * 5, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, x)
def ar():
    dist.all_reduce(x)
def step():
    cls.grad = None; reg.grad = None
    l = gdist.sharded_focal_loss(cls, reg, anc, ann); l.backward(ones)
def step_local():
    cls.grad = None; reg.grad = None
    l = losses_impl.focal_loss(cls, reg, anc, ann)[0]; l.backward(ones)
def fwd_only():
    with torch.no_grad():
        gdist.sharded_focal_loss(cls, reg, anc, ann)
res = dict(all_gather=timeit(ag), all_gather_into_tensor=timeit(ag2), all_reduce=timeit(ar), step_local=timeit(step_local), step_sharded=timeit(step), fwd_sharded_nograd=timeit(fwd_only))
if rank == 0: print(res)
dist.barrier(); dist.destroy_process_group()
