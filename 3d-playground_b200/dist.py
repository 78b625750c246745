"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests) for the only exchange the path has - the scalar loss / normaliser reduction.

The reference runs the loss under nn.DataParallel (train_detector_3D_angle.py:316-318): images are split across
replicas, every replica returns its own batch means and the trainer averages them (:374-378).  Here every rank owns a
contiguous range of images, computes the per-image terms locally with the fused kernel, and 5 scalars per rank are
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


def local_stats(per_image, gt_count):
    """per_image[B_l,4] = (cls_j, reg_j, vp_j, num_pos_j), gt_count[B_l] -> float64[5] shard statistics"""
    nonempty = gt_count > 0
    pi = per_image.double()
    return torch.stack((pi[:, 0].sum(), pi[:, 1].sum(), (pi[:, 2] * nonempty).sum(),
                        torch.tensor(float(per_image.shape[0]), dtype=torch.float64, device=per_image.device),
                        nonempty.sum().double()))


def combine_stats(stats, group=None):
    """all-gather the [5] statistics of every rank and reduce them in rank order.
    Returns (losses float32[3] = global (cls, reg, vp) means, totals float64[5])."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        gathered = [torch.empty_like(stats) for _ in range(dist.get_world_size(group))]
        dist.all_gather(gathered, stats.contiguous(), group=group)
        total = torch.stack(gathered).sum(dim=0)   # fixed (rank) order
    else:
        total = stats
    losses = torch.stack((total[0] / total[3], total[1] / total[3], total[2] / total[4])).to(torch.float32)
    return losses, total


class _ShardedFocalLossFn(torch.autograd.Function):
    """forward: local fused loss (+ gradients for the expected upstream value) -> all-gather of the 5 shard statistics
    the kernel wrote -> one tiny kernel forms the global means and this rank's gradient scales.  Two extra launches and
    one collective per step; the backward adds none (the scale is applied inside the gradient kernels)."""

    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations, group, trace_events=None):
        from . import ops
        on = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(group) if on else 1
        rank = dist.get_rank(group) if on else 0
        # expected upstream gradient of the LOCAL classification mean: B_local / B_global = 1 / world for equal shards
        # (a hint: the backward kernel checks it against the real value on the device and recomputes if it is off)
        needs_grad = classifications.requires_grad or regressions.requires_grad
        fwd = ops.focal_loss_forward(classifications, regressions, anchors, annotations,
                                     grad_cls_expected=(1.0 / world) if needs_grad else None,
                                     trace_events=trace_events if needs_grad else None, want_shard_stats=True)
        stats = fwd["shard_stats"]
        if world > 1:
            gathered = torch.empty((world, 5), dtype=torch.float64, device=stats.device)
            dist.all_gather_into_tensor(gathered, stats, group=group)
        else:
            gathered = stats.reshape(1, 5)
        losses, scale = ops.combine_shard_stats(gathered, rank)
        ctx.fwd, ctx.scale = fwd, scale
        ctx.in_dtypes = (classifications.dtype, regressions.dtype)
        return losses

    @staticmethod
    def backward(ctx, g):
        from . import ops
        dcls, dreg = ops.focal_loss_backward(ctx.fwd, g.to(torch.float32), grad_scale=ctx.scale)
        ctx.fwd = None
        return dcls.to(ctx.in_dtypes[0]), dreg.to(ctx.in_dtypes[1]), None, None, None, None


def sharded_focal_loss(classifications, regressions, anchors, annotations, group=None, trace_events=None):
    """Loss of the global batch from this rank's image shard.  Returns float32[3] (cls, reg, vp), identical on every
    rank and differentiable w.r.t. the local classifications / regressions (gradients need no collective: they are
    per-image local, scaled by 1/B_global)."""
    return _ShardedFocalLossFn.apply(classifications, regressions, anchors, annotations, group, trace_events)
