"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's IoU assignment and losses.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It restates, with plain PyTorch CPU ops in the reference's operation order (so the float32 roundings coincide):

    calc_iou                    retinanet/losses.py:5-22 == pytorch_retinanet_detector_directional/retinanet/losses.py:5-22
    assignment                  3D losses.py:93-131 ; 2D retinanet/losses.py:81-104
    focal classification loss   3D losses.py:56-87,133-152 ; 2D retinanet/losses.py:49-79,106-125
    3D corner / direction loss  3D losses.py:156-358
    2D box regression loss      2D retinanet/losses.py:129-173
    batch reduction             3D losses.py:359-362 ; 2D retinanet/losses.py:175-177

Parity pin: tests/test_oracle_golden.py checks every function here against vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py imports it from /root/reference), bit-exactly for IoU / argmax / assignment.
Everything is differentiable, so torch.autograd on these functions is the oracle for the backward kernels too.
"""
import torch

ALPHA = 0.25
GAMMA = 2.0
TOP_WEIGHTING = 0.5
# corner k = centre + sl*L + sw*W + sh*H  (3D losses.py:311-327; same table in utils.py:114-130)
SIGN_L = (-1, -1, +1, +1, -1, -1, +1, +1)
SIGN_W = (-1, +1, -1, +1, -1, +1, -1, +1)
SIGN_H = (+1, +1, +1, +1, -1, -1, -1, -1)


def calc_iou(a, b):
    """[A,4] x [G,4] -> [A,G] float32 IoU; union clamped at 1e-8."""
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    iw = torch.min(a[:, 2:3], b[:, 2]) - torch.max(a[:, 0:1], b[:, 0])
    ih = torch.min(a[:, 3:4], b[:, 3]) - torch.max(a[:, 1:2], b[:, 1])
    iw = iw.clamp(min=0)
    ih = ih.clamp(min=0)
    union = (area_a[:, None] + area_b - iw * ih).clamp(min=1e-8)
    return (iw * ih) / union


def valid_rows(annotation, variant_3d):
    """rows whose class column is not -1 (3D: col 20 of the first 21 columns; 2D: col 4)."""
    col = 20 if variant_3d else 4
    return annotation[annotation[:, col] != -1]


def assignment_boxes(rows, variant_3d):
    """the 2D boxes anchors are matched against: 3D = min/max over the 8 projected corners, 2D = cols 0:4."""
    if not variant_3d:
        return rows[:, :4]
    xs, ys = rows[:, 0:16:2], rows[:, 1:16:2]
    return torch.stack((xs.min(1).values, ys.min(1).values, xs.max(1).values, ys.max(1).values), dim=1)


def assign(anchors, annotation, variant_3d, pos_iou=0.5, neg_iou=0.4):
    """One image.  Returns (iou_max[A], iou_argmax[A] into the filtered rows, positive[A] bool, negative[A] bool, rows).
    pos_iou / neg_iou: the reference's constants (3D losses.py:124 / :121), exposed for the drop-in's keyword arguments."""
    rows = valid_rows(annotation, variant_3d)
    A = anchors.shape[0]
    if rows.shape[0] == 0:
        z = torch.zeros(A)
        return z, torch.zeros(A, dtype=torch.int64), torch.zeros(A, dtype=torch.bool), torch.ones(A, dtype=torch.bool), rows
    iou = calc_iou(anchors, assignment_boxes(rows, variant_3d))
    iou_max, iou_argmax = iou.max(dim=1)
    return iou_max, iou_argmax, iou_max >= pos_iou, iou_max < neg_iou, rows


def focal_classification_sum(classification, positive, negative, pos_class, alpha=ALPHA, gamma=GAMMA, clamp_min=1e-4,
                             clamp_max=1.0 - 1e-4):
    """sum over anchors x classes of alpha_t (1-p_t)^gamma * bce, ignoring anchors that are neither positive nor negative."""
    p = classification.clamp(clamp_min, clamp_max)
    targets = torch.full_like(p, -1.0)
    targets[negative] = 0
    targets[positive] = 0
    targets[positive, pos_class[positive]] = 1
    is_pos = targets == 1.0
    alpha_t = torch.where(is_pos, torch.full_like(p, alpha), 1.0 - torch.full_like(p, alpha))
    weight = alpha_t * torch.pow(torch.where(is_pos, 1.0 - p, p), gamma)
    bce = -(targets * torch.log(p) + (1.0 - targets) * torch.log(1.0 - p))
    loss = torch.where(targets != -1.0, weight * bce, torch.zeros_like(p))
    return loss.sum()


def smooth_l1(diff, beta=None):
    if beta is None:        # the reference's literal expression (3D losses.py:345-349)
        return torch.where(diff <= 1.0 / 9.0, 0.5 * 9.0 * torch.pow(diff, 2), diff - 0.5 / 9.0)
    return torch.where(diff <= beta, (0.5 / beta) * torch.pow(diff, 2), diff - 0.5 * beta)


def corner_predictions(reg):
    """[P,12] -> [P,20]: 8 corners (x,y) from centre + direction vectors, then the 2D box regression passthrough."""
    cols = []
    for k in range(8):
        for c in (0, 1):
            v = reg[:, c] + SIGN_L[k] * reg[:, 2 + c]
            v = v + SIGN_W[k] * reg[:, 4 + c]
            v = v + SIGN_H[k] * reg[:, 6 + c]
            cols.append(v)
    cols += [reg[:, 8], reg[:, 9], reg[:, 10], reg[:, 11]]
    return torch.stack(cols, dim=1)


def direction_targets(t):
    """GT direction vectors from the raw 16 corner coordinates: (front->back, left->right, top->bottom), each (x,y)."""
    def pick(ix):
        return t[:, ix[0]] + t[:, ix[1]] + t[:, ix[2]] + t[:, ix[3]]
    out = []
    for plus, minus in (((4, 6, 12, 14), (0, 2, 8, 10)), ((2, 6, 10, 14), (0, 4, 8, 12)), ((0, 2, 4, 6), (8, 10, 12, 14))):
        for c in (0, 1):
            out.append((pick([i + c for i in plus]) - pick([i + c for i in minus])) / 4.0)
    return out  # [lx, ly, wx, wy, hx, hy]


def cosine_loss(rx, ry, tx, ty):
    rn = torch.sqrt(torch.pow(rx, 2) + torch.pow(ry, 2))
    tn = torch.sqrt(torch.pow(tx, 2) + torch.pow(ty, 2))
    return 1 - (rx * tx + ry * ty) / (rn * tn)


def regression_terms_3d(reg_pos, gt_pos, anchors_pos, top_weighting=TOP_WEIGHTING, beta=None):
    """positives only.  Returns (regression_loss, vp_loss) scalars of one image (means over P x 20 and over P)."""
    t = gt_pos[:, :20]
    d = direction_targets(t)
    vp = (cosine_loss(reg_pos[:, 2], reg_pos[:, 3], d[0], d[1]) + cosine_loss(reg_pos[:, 4], reg_pos[:, 5], d[2], d[3])
          + cosine_loss(reg_pos[:, 6], reg_pos[:, 7], d[4], d[5])) / 3.0
    aw = anchors_pos[:, 2] - anchors_pos[:, 0]
    ah = anchors_pos[:, 3] - anchors_pos[:, 1]
    acx = anchors_pos[:, 0] + 0.5 * aw
    acy = anchors_pos[:, 1] + 0.5 * ah
    tn = torch.empty_like(t)
    tn[:, 0::2] = (t[:, 0::2] - acx[:, None]) / aw[:, None]
    tn[:, 1::2] = (t[:, 1::2] - acy[:, None]) / ah[:, None]
    diff = torch.abs(tn - corner_predictions(reg_pos))
    w = torch.ones(20)
    w[8:16] = top_weighting
    return smooth_l1(diff * w, beta).mean(), vp.mean()


def regression_terms_2d(reg_pos, gt_pos, anchors_pos, beta=None):
    aw = anchors_pos[:, 2] - anchors_pos[:, 0]
    ah = anchors_pos[:, 3] - anchors_pos[:, 1]
    acx = anchors_pos[:, 0] + 0.5 * aw
    acy = anchors_pos[:, 1] + 0.5 * ah
    gw = gt_pos[:, 2] - gt_pos[:, 0]
    gh = gt_pos[:, 3] - gt_pos[:, 1]
    gcx = gt_pos[:, 0] + 0.5 * gw
    gcy = gt_pos[:, 1] + 0.5 * gh
    gw, gh = gw.clamp(min=1), gh.clamp(min=1)
    t = torch.stack(((gcx - acx) / aw, (gcy - acy) / ah, torch.log(gw / aw), torch.log(gh / ah)), dim=1)
    t = t / torch.tensor([[0.1, 0.1, 0.2, 0.2]])
    return smooth_l1(torch.abs(t - reg_pos), beta).mean()


def focal_loss(classifications, regressions, anchors, annotations, hyper=None):
    """Whole-batch loss.  3D (regression width 12) -> (cls[1], reg[1], vp[1]); 2D (width 4) -> (cls[1], reg[1]).
    Also returns per-image details as a 4th/3rd element: [(iou_max, iou_argmax, positive, negative, n_rows)] per image.
    hyper: overrides of the constants the reference hard-codes (alpha, gamma, top_weighting, pos_iou, neg_iou, beta,
    clamp_min, clamp_max) - the drop-in exposes them as keyword arguments with the reference's values as defaults."""
    hp = dict(hyper or {})
    cls_kw = {k: hp[k] for k in ("alpha", "gamma", "clamp_min", "clamp_max") if k in hp}
    thr_kw = {k: hp[k] for k in ("pos_iou", "neg_iou") if k in hp}
    reg_kw = {k: hp[k] for k in ("beta",) if k in hp}
    reg3_kw = dict(reg_kw, **({"top_weighting": hp["top_weighting"]} if "top_weighting" in hp else {}))
    variant_3d = regressions.shape[-1] == 12
    anchor = anchors.reshape(-1, 4)
    B = classifications.shape[0]
    cls_l, reg_l, vp_l, info = [], [], [], []
    for j in range(B):
        ann = annotations[j, :, :21] if variant_3d else annotations[j]
        iou_max, iou_argmax, positive, negative, rows = assign(anchor, ann, variant_3d, **thr_kw)
        n_pos = positive.sum()
        if rows.shape[0] == 0:
            pos_class = torch.zeros(anchor.shape[0], dtype=torch.int64)
            cls_l.append(focal_classification_sum(classifications[j], positive, negative, pos_class, **cls_kw))
            reg_l.append(torch.tensor(0.0))
            info.append((iou_max, iou_argmax, positive, negative, 0))
            continue
        assigned = rows[iou_argmax]
        pos_class = assigned[:, 20 if variant_3d else 4].long()
        s = focal_classification_sum(classifications[j], positive, negative, pos_class, **cls_kw)
        cls_l.append(s / n_pos.float().clamp(min=1.0))
        if n_pos > 0:
            if variant_3d:
                r, v = regression_terms_3d(regressions[j][positive], assigned[positive], anchor[positive], **reg3_kw)
                vp_l.append(v)
            else:
                r = regression_terms_2d(regressions[j][positive], assigned[positive], anchor[positive], **reg_kw)
            reg_l.append(r)
        else:
            reg_l.append(torch.tensor(0.0))
            if variant_3d:
                vp_l.append(torch.tensor(0.0))
        info.append((iou_max, iou_argmax, positive, negative, rows.shape[0]))
    out = [torch.stack(cls_l).mean(dim=0, keepdim=True), torch.stack(reg_l).mean(dim=0, keepdim=True)]
    if variant_3d:
        out.append(torch.stack(vp_l).mean(dim=0, keepdim=True))   # raises on an all-empty batch, as the reference does
    return tuple(out) + (info,)
