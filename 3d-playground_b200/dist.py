"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests) for the only exchange the path has - the scalar loss / normaliser reduction.

The reference runs the loss under nn.DataParallel (train_detector_3D_angle.py:316-318): images are split across
replicas, every replica returns its own batch means and the trainer averages them (:374-378).  Here every rank owns a
contiguous range of images, computes the per-image terms locally with the fused kernels, and 5 scalars per rank are
all-gathered and summed in rank order (bit-reproducible, unlike a tree all-reduce whose order depends on topology):

    [sum_j cls_j, sum_j reg_j, sum_{j non-empty} vp_j, #images, #non-empty images]

The result is the loss of the GLOBAL batch exactly as the reference forms it on one device (mean over all images; the
vp mean over the images that have ground truth), which equals DataParallel's mean-of-means when shards are equal.
Decode / NMS / homography work shards by image / state range with no collective at all.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """contiguous [lo, hi) slice of n_items owned by `rank`; sizes differ by at most one"""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_shard_stats(stats, group=None):
    """stats: this rank's float64[5] shard statistics (as g3d_focal_loss_fwd_bwd writes them; any device the backend
    supports).  Returns float64[world,5] in rank order - the input of ops.combine_shard_stats / combine_on_host."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        gathered = torch.empty((world * 5,), dtype=torch.float64, device=stats.device)   # flat: gloo and NCCL both take it
        dist.all_gather_into_tensor(gathered, stats.contiguous().reshape(5), group=group)
        return gathered.reshape(world, 5)
    return stats.reshape(1, 5)


def combine_on_host(gathered, rank):
    """The arithmetic of g3d_combine_shard_stats (focal_loss.cu: combine_shard_stats_kernel) restated with torch ops, for
    checking the kernel: rank-order sums, global means, this rank's gradient scales.  Not used by the GPU path."""
    total = torch.zeros(5, dtype=torch.float64)
    for r in range(gathered.shape[0]):
        total = total + gathered[r].cpu()
    losses = torch.stack((total[0] / total[3], total[1] / total[3], total[2] / total[4])).to(torch.float32)
    bl, nl = gathered[rank, 3].cpu(), gathered[rank, 4].cpu()
    scale = torch.stack((bl / total[3], bl / total[3], nl / total[4] if float(nl) > 0 else torch.zeros((), dtype=torch.float64)))
    return losses, scale.to(torch.float32)


class _PeerExchange:
    """Peer-mapped exchange buffers of the ranks of one NVLink / NVSwitch box (torch.distributed._symmetric_memory): the
    shard statistics travel as plain stores into the peers' memory from one tiny kernel (ops.exchange_shard_stats) - no
    NCCL call inside the step.  Falls back to the NCCL all-gather when symmetric memory cannot be set up."""
    _cache = {}

    def __init__(self, group, dev):
        self.group, self.dev = group, dev
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buf = self.handle = None
        self.peer_ptrs_dev = 0

    def _allocate(self):
        """local step (no communication): the symmetric buffer"""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        n = int(_lib.lib().g3d_exchange_buffer_doubles(self.world))
        self.buf = symm_mem.empty(n, dtype=torch.float64, device=self.dev)
        self.buf.zero_()

    def _rendezvous(self):
        """collective step: every rank of the group must call it"""
        import torch.distributed._symmetric_memory as symm_mem
        pg = self.group if self.group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.buf, pg)
        self.peer_ptrs_dev = int(self.handle.buffer_ptrs_dev)

    @classmethod
    def get(cls, group, dev):
        """the exchange of (group, device), or None if it is disabled (G3D_PEER_EXCHANGE=0) or not available.  All ranks
        take the same branch: each step that can fail locally is followed by an all-reduce of the outcome, and the
        collective rendezvous is only entered once every rank has its buffer."""
        import os
        key = (id(group), dev.index)
        if key in cls._cache:
            return cls._cache[key]

        def everyone(ok):
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return bool(int(flag.item()))

        ex, ok = None, os.environ.get("G3D_PEER_EXCHANGE", "1") != "0" and dev.type == "cuda"
        if ok:
            try:
                ex = cls(group, dev)
                ex._allocate()
            except Exception:   # noqa: BLE001 - no symmetric memory on this system: use NCCL
                ok = False
        if everyone(ok):
            try:
                ex._rendezvous()
            except Exception:   # noqa: BLE001 - no peer access between these devices
                ok = False
            ok = everyone(ok)
        else:
            ok = False
        if ok:
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)                  # every buffer is zero before anybody stores into it
        cls._cache[key] = ex if ok else None
        return cls._cache[key]


class _ShardedFocalLossFn(torch.autograd.Function):
    """forward: local fused loss (+ gradients for the expected upstream value B_local / B_global = 1 / world) -> exchange of
    the 5 shard statistics the kernel wrote -> global means and this rank's gradient scales.  On one NVLink box the exchange
    is ONE tiny kernel that stores into the peers' memory (_PeerExchange); otherwise an NCCL all-gather + a combine kernel.
    The backward adds nothing (it only verifies the scale on the device)."""

    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations, group, trace_events, hyper):
        from . import ops
        on = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(group) if on else 1
        rank = dist.get_rank(group) if on else 0
        needs_grad = classifications.requires_grad or regressions.requires_grad
        fwd = ops.focal_loss_forward(classifications, regressions, anchors, annotations, want_assign=False,
                                     grad_expected=(1.0 / world) if needs_grad else None,
                                     trace_events=trace_events, want_shard_stats=True, hyper=hyper)
        exchange = _PeerExchange.get(group, classifications.device) if world > 1 else None
        if exchange is not None:       # stores over NVLink into the peers' buffers: one launch, no collective call
            losses, scale = ops.exchange_shard_stats(fwd["shard_stats"], exchange.peer_ptrs_dev, world, rank)
        else:
            gathered = gather_shard_stats(fwd["shard_stats"], group)
            losses, scale = ops.combine_shard_stats(gathered, rank)
        ctx.fwd, ctx.scale = fwd, scale
        ctx.in_dtypes = (classifications.dtype, regressions.dtype)
        ctx.save_for_backward(classifications, regressions)
        return losses

    @staticmethod
    def backward(ctx, g):
        from . import ops
        _ = ctx.saved_tensors
        dcls, dreg = ops.focal_loss_backward(ctx.fwd, g.to(torch.float32), grad_scale=ctx.scale, take=True)
        return dcls.to(ctx.in_dtypes[0]), dreg.to(ctx.in_dtypes[1]), None, None, None, None, None


def sharded_focal_loss(classifications, regressions, anchors, annotations, group=None, trace_events=None, hyper=None):
    """Loss of the global batch from this rank's image shard.  Returns float32[3] (cls, reg, vp), identical on every
    rank and differentiable w.r.t. the local classifications / regressions (gradients need no collective: they are
    per-image local, scaled by 1/B_global)."""
    return _ShardedFocalLossFn.apply(classifications, regressions, anchors, annotations, group, trace_events, hyper)
