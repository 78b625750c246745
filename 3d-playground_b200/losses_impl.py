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
    """losses[3] = (cls, reg, vp).  When an input requires grad, the forward pass already writes the classification
    gradient for the expected upstream gradient (1 for `(cls + reg + vp).backward()`,
    train_detector_3D_angle.py:374-382) in the same sweep over the classification tensor; the backward kernel verifies
    that expectation on the device and recomputes only if it does not hold."""

    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations, expected_grad, trace_events):
        needs_grad = classifications.requires_grad or regressions.requires_grad
        fwd = ops.focal_loss_forward(classifications, regressions, anchors, annotations,
                                     grad_cls_expected=float(expected_grad) if needs_grad else None,
                                     trace_events=trace_events if needs_grad else None)
        ctx.fwd = fwd
        ctx.in_dtypes = (classifications.dtype, regressions.dtype)
        ctx.mark_non_differentiable(fwd["per_image"], fwd["gt_count"])
        return fwd["losses"][:3].clone(), fwd["losses"][3:4].clone(), fwd["per_image"], fwd["gt_count"]

    @staticmethod
    def backward(ctx, g_losses, _g_ne, _g_pi, _g_gc):
        dcls, dreg = ops.focal_loss_backward(ctx.fwd, g_losses.to(torch.float32).contiguous())
        ctx.fwd = None
        return dcls.to(ctx.in_dtypes[0]), dreg.to(ctx.in_dtypes[1]), None, None, None, None


def focal_loss(classifications, regressions, anchors, annotations, expected_grad=1.0, trace_events=None):
    """Functional form.  Returns (losses[3], n_nonempty[1], per_image[B,4], gt_count[B]); losses is differentiable
    w.r.t. classifications and regressions.  expected_grad: the upstream gradient the caller expects for the
    classification loss (a performance hint only - any upstream gradient gives the right result).  trace_events: see
    ops.focal_loss_forward (per-kernel timing for bench.py)."""
    return _FocalLossFn.apply(classifications, regressions, anchors, annotations, expected_grad, trace_events)


class FocalLoss(nn.Module):
    """Drop-in for both FocalLoss classes: no constructor arguments needed, no parameters, no buffers.

    forward(classifications[B,A,C], regressions[B,A,12|4], anchors[1,A,4], annotations[B,G,>=21|5])
      -> 3D: (cls[1], reg[1], vp[1])      (losses.py:362)
         2D: (cls[1], reg[1])             (retinanet/losses.py:177)

    check_empty=True keeps the reference's error behaviour for the 3D copy - a batch in which no image has a ground
    truth row makes torch.stack([]) raise RuntimeError (losses.py:362) - at the cost of one 4-byte device->host read;
    with check_empty=False the vp loss is NaN in that case and no synchronisation happens.
    """

    def __init__(self, check_empty=True, expected_grad=1.0):
        super().__init__()
        self.check_empty = check_empty
        self.expected_grad = expected_grad   # e.g. 1/n_replicas under nn.DataParallel + .mean() (a hint, see focal_loss)

    def forward(self, classifications, regressions, anchors, annotations):
        losses, n_nonempty, _, _ = focal_loss(classifications, regressions, anchors, annotations, self.expected_grad)
        if regressions.shape[-1] == 12:
            if self.check_empty and float(n_nonempty.item()) == 0.0:
                raise RuntimeError("stack expects a non-empty TensorList")  # the reference's torch.stack(vp_losses)
            return losses[0:1], losses[1:2], losses[2:3]
        return losses[0:1], losses[1:2]
