"""Host side of the FocalLoss drop-ins (both retinanet copies) over the fused CUDA kernels.

Reference: pytorch_retinanet_detector_directional/retinanet/losses.py:24-362 (3D: 3 outputs, 12-d regression,
>= 21 annotation columns) and retinanet/losses.py:24-177 (2D: 2 outputs, 4-d regression, 5 annotation columns).
The variant is chosen from the regression width, so one implementation serves both module trees.
"""
import torch
import torch.nn as nn

from . import ops


def calc_iou(a, b):
    """calc_iou(a[A,4], b[G,4]) -> [A,G]   (losses.py:5-22, identical in both copies)."""
    return ops.calc_iou(a, b)


class _FocalLossFn(torch.autograd.Function):
    """losses[3] = (cls, reg, vp).  When an input requires grad, the forward launches already write both gradients for the
    expected upstream gradients (1 for `(cls + reg + vp).backward()`, train_detector_3D_angle.py:374-382); the backward
    kernels verify that expectation on the device and recompute only what differs."""

    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations, expected_grad, trace_events, hyper, persistent=None):
        needs_grad = classifications.requires_grad or regressions.requires_grad
        fwd = ops.focal_loss_forward(classifications, regressions, anchors, annotations, want_assign=False,
                                     grad_expected=expected_grad if needs_grad else None,
                                     trace_events=trace_events, hyper=hyper, persistent=persistent)
        ctx.fwd = fwd
        ctx.in_dtypes = (classifications.dtype, regressions.dtype)
        # saved for autograd's bookkeeping: an in-place edit of an input between forward and backward is detected (version
        # counters), and a second backward without retain_graph raises autograd's own error
        ctx.save_for_backward(classifications, regressions)
        n_nonempty = fwd["losses"][3:4].clone()
        ctx.mark_non_differentiable(n_nonempty, fwd["per_image"], fwd["gt_count"])
        return fwd["losses"][:3].clone(), n_nonempty, fwd["per_image"], fwd["gt_count"]

    @staticmethod
    def backward(ctx, g_losses, _g_ne, _g_pi, _g_gc):
        _ = ctx.saved_tensors
        # first backward: the buffers the forward wrote (verified / completed on the device) leave ctx.fwd, so autograd can
        # adopt them as .grad without a copy; later ones (retain_graph) compute into fresh buffers
        dcls, dreg = ops.focal_loss_backward(ctx.fwd, g_losses.to(torch.float32).contiguous(), take=True)
        return dcls.to(ctx.in_dtypes[0]), dreg.to(ctx.in_dtypes[1]), None, None, None, None, None, None


def focal_loss(classifications, regressions, anchors, annotations, expected_grad=1.0, trace_events=None, hyper=None,
               persistent=None):
    """Functional form.  Returns (losses[3], n_nonempty[1], per_image[B,4], gt_count[B]); losses is differentiable
    w.r.t. classifications and regressions.  expected_grad: the upstream gradient(s) the caller expects for the three
    losses, a float or (cls, reg, vp) (a performance hint only - any upstream gradient gives the right result).
    trace_events: see ops.focal_loss_forward (per-kernel timing for bench.py).  hyper: dict of hyper-parameter overrides.
    persistent: an ops.PersistentGrads (see there) - the gradients are then views of its buffers."""
    return _FocalLossFn.apply(classifications, regressions, anchors, annotations, expected_grad, trace_events, hyper,
                              persistent)


class FocalLoss(nn.Module):
    """Drop-in for both FocalLoss classes: no constructor arguments needed, no parameters, no buffers.

    forward(classifications[B,A,C], regressions[B,A,12|4], anchors[1,A,4], annotations[B,G,>=21|5])
      -> 3D: (cls[1], reg[1], vp[1])      (losses.py:362)
         2D: (cls[1], reg[1])             (retinanet/losses.py:177)

    The reference hard-codes its hyper-parameters inside forward; here they are keyword arguments with the same defaults:
    alpha, gamma, top_weighting (losses.py:28-30), pos_iou / neg_iou (:124 / :121), beta (smooth-L1, :346-348), clamp_min /
    clamp_max (:56).

    check_empty - the reference's error for a 3D batch in which no image has a ground truth row (torch.stack([]) raises
    RuntimeError, losses.py:362):
      "lazy" (default): no synchronisation; the count of non-empty images is copied to pinned host memory asynchronously
                        and examined at the NEXT forward call (or by .check()), which raises then; the step itself
                        returns NaN for the vp loss.
      True:             raise immediately (one 4-byte device->host read per forward, a host synchronisation).
      False:            never raise (NaN vp loss).

    persistent_grad (default False) - keep the two gradient buffers and the workspace from step to step
    (ops.PersistentGrads): the regression gradient's zeros (48 B per anchor, 43 % of the bytes a step moves) are then
    written once, and each step only clears the ~1 % of rows the previous step wrote: 0.31 -> ~0.21 ms per step at
    BASELINE configs[1].  Opt-in because it changes who owns the gradient: what autograd receives are the module's
    buffers, valid until the next forward - fine when `regressions` / `classifications` are network outputs (their
    producer's backward reads them at once), wrong for code that keeps `.grad` of a leaf across steps.
    """

    def __init__(self, check_empty="lazy", expected_grad=1.0, alpha=0.25, gamma=2.0, top_weighting=0.5, pos_iou=0.5,
                 neg_iou=0.4, beta=1.0 / 9.0, clamp_min=1e-4, clamp_max=1.0 - 1e-4, persistent_grad=False):
        super().__init__()
        self._persistent = ops.PersistentGrads() if persistent_grad else None
        self.check_empty = check_empty
        self.expected_grad = expected_grad   # e.g. 1/n_replicas under nn.DataParallel + .mean() (a hint, see focal_loss)
        given = dict(alpha=alpha, gamma=gamma, top_weighting=top_weighting, pos_iou=pos_iou, neg_iou=neg_iou, beta=beta,
                     clamp_min=clamp_min, clamp_max=clamp_max)
        self.hyper = {k: v for k, v in given.items() if float(v) != float(dict(ops.HYPER_DEFAULTS)[k])} or None
        self._pending = []                   # (pinned host count, event) of earlier forwards, for check_empty="lazy"

    def check(self, wait=False):
        """examine the empty-batch flags of earlier forwards (check_empty="lazy"); wait=True synchronises on them"""
        still = []
        for host, ev in self._pending:
            if wait:
                ev.synchronize()
            if ev.query():
                if float(host[0]) == 0.0:
                    self._pending = []
                    raise RuntimeError("stack expects a non-empty TensorList (an earlier FocalLoss.forward saw a batch "
                                       "without any ground-truth row; reported late because check_empty='lazy')")
            else:
                still.append((host, ev))
        self._pending = still

    def forward(self, classifications, regressions, anchors, annotations):
        losses, n_nonempty, _, _ = focal_loss(classifications, regressions, anchors, annotations, self.expected_grad,
                                              hyper=self.hyper, persistent=self._persistent)
        if regressions.shape[-1] == 12:
            if self.check_empty is True:
                if float(n_nonempty.item()) == 0.0:
                    raise RuntimeError("stack expects a non-empty TensorList")  # the reference's torch.stack(vp_losses)
            elif self.check_empty == "lazy" and not torch.cuda.is_current_stream_capturing():
                self.check()
                host = torch.empty(1, dtype=torch.float32, pin_memory=True)
                host.copy_(n_nonempty, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(n_nonempty.device))
                self._pending.append((host, ev))
                if len(self._pending) > 64:
                    self.check(wait=True)
            return losses[0:1], losses[1:2], losses[2:3]
        return losses[0:1], losses[1:2]
