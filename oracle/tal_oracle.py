"""CPU oracle for the Stack-B training loss -- TEST INFRASTRUCTURE ONLY.

A plain-torch restatement of `ComputeLoss.__call__` of utils/loss_tal.py:105-215 (DFL on): anchor points
(utils/tal/anchor_generator.py:8-21), DFL box decode (:154-160), `TaskAlignedAssigner` (utils/tal/assigner.py:51-179,
topk 10, alpha 0.5, beta 6; CIoU overlaps from utils/metrics2.py:254-289), BCE class term (optionally inside FocalLoss,
:32-60), the box term -- requested as SIoU, evaluated as GIoU by utils/metrics2.py:279-311, see box_term_iou -- and
the distribution focal loss (:63-103).  It is written per image over the image's OWN label list (the reference pads
every image to the longest list; padded rows never become positive) because that is the formulation the CUDA kernels
use.
Where the reference leaves a choice open -- `torch.topk` among equal (zero) metrics, `argmax` among equal overlaps -- the
lowest index wins here.  Pinned by tests/test_oracle_post.py against fixtures made by the UNMODIFIED reference
(oracle/gen_golden_tal.py -> tests/golden/post_tal.pt).  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn.functional as F


TOPK, ALPHA, BETA, EPS = 10, 0.5, 6.0, 1e-9          # utils/loss_tal.py:134-137, utils/tal/assigner.py:52
REG_MAX = 16                                         # models/yolo_snn.py:95
GAINS = (7.5, 0.5, 1.5)                              # box, cls, dfl (utils/loss_tal.py:210-212)


def anchors_of(grids, strides):
    """-> anchor points [A, 2] in grid units (cell centres), stride per anchor [A, 1]."""
    pts, st = [], []
    for (ny, nx), s in zip(grids, strides):
        sy, sx = torch.meshgrid(torch.arange(ny).float() + 0.5, torch.arange(nx).float() + 0.5, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((ny * nx, 1), float(s)))
    return torch.cat(pts), torch.cat(st)


def _parts(b1, b2, eps=1e-7):
    """Shared head of utils/metrics2.py:264-281 for xyxy boxes (broadcasting)."""
    x11, y11, x12, y12 = b1.unbind(-1)
    x21, y21, x22, y22 = b2.unbind(-1)
    w1, h1 = x12 - x11, y12 - y11 + eps
    w2, h2 = x22 - x21, y22 - y21 + eps
    inter = (torch.min(x12, x22) - torch.max(x11, x21)).clamp(0) * (torch.min(y12, y22) - torch.max(y11, y21)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    cw = torch.max(x12, x22) - torch.min(x11, x21)
    ch = torch.max(y12, y22) - torch.min(y11, y21)
    return (x11, y11, x12, y12, x21, y21, x22, y22, w1, h1, w2, h2, inter / union, cw, ch, union)


def ciou_xyxy(b1, b2, eps=1e-7):
    """utils/metrics2.py:282-289"""
    x11, y11, x12, y12, x21, y21, x22, y22, w1, h1, w2, h2, iou, cw, ch, _ = _parts(b1, b2, eps)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((x21 + x22 - x11 - x12) ** 2 + (y21 + y22 - y11 - y12) ** 2) / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def box_term_iou(b1, b2, eps=1e-7):
    """What `bbox_iou(pred, target, xywh=False, SIoU=True)` of utils/metrics2.py evaluates (utils/loss_tal.py:74): the
    SIoU branch (:290-308) sits INSIDE `if CIoU or DIoU:` (:282), so with only SIoU=True control falls through to the
    GIoU return (:310-311)."""
    *_, iou, cw, ch, union = _parts(b1, b2, eps)
    c_area = cw * ch + eps
    return iou - (c_area - union) / c_area


def assign_image(scores, boxes_px, pts_px, labels, gts, topk=TOPK, alpha=ALPHA, beta=BETA):
    """TaskAlignedAssigner for ONE image: scores [A, nc] (sigmoid), boxes_px [A, 4], pts_px [A, 2], labels [M] (long),
    gts [M, 4] xyxy pixels (all valid).  -> (gt index per anchor [A] long, fg [A] bool, normalised score per anchor [A])."""
    A, M = scores.shape[0], gts.shape[0]
    if M == 0:
        return torch.zeros(A, dtype=torch.long), torch.zeros(A, dtype=torch.bool), torch.zeros(A)
    ov = ciou_xyxy(gts[:, None, :], boxes_px[None, :, :]).clamp(0)                       # [M, A] (:120)
    align = scores[:, labels].T.pow(alpha) * ov.pow(beta)                               # (:121)
    deltas = torch.cat((pts_px[None] - gts[:, None, :2], gts[:, None, 2:] - pts_px[None]), 2)
    in_gt = deltas.amin(2) > EPS                                                        # (:8-22)
    metric = align * in_gt
    order = torch.sort(metric, dim=1, descending=True, stable=True).indices[:, :min(topk, A)]   # ties: lowest index
    in_top = torch.zeros(M, A, dtype=torch.bool)
    in_top[torch.arange(M)[:, None], order] = True
    pos = in_top & in_gt                                                                # (:105)
    cnt = pos.sum(0)
    multi = cnt > 1                                                                     # (:38-45)
    best = ov.argmax(0)                                                                 # first maximum
    pos = torch.where(multi[None, :], F.one_hot(best, M).T.bool(), pos)
    fg = pos.sum(0) > 0
    gt_idx = pos.float().argmax(0)
    align = align * pos                                                                 # (:92-96)
    pos_align = align.amax(1, keepdim=True)
    pos_ov = (ov * pos).amax(1, keepdim=True)
    norm = (align * pos_ov / (pos_align + EPS)).amax(0)
    return gt_idx, fg, norm


def compute_loss(feats: Sequence[torch.Tensor], targets: torch.Tensor, strides, cls_pw: float = 1.0, fl_gamma: float = 0.0,
                 assigner=(TOPK, ALPHA, BETA)):
    """-> (loss [1]-shaped scalar tensor, loss_items [3] = (box, cls, dfl) detached, number of foreground anchors)."""
    N, no = feats[0].shape[:2]
    nc = no - 4 * REG_MAX
    grids = [tuple(f.shape[2:]) for f in feats]
    flat = torch.cat([f.reshape(N, no, -1) for f in feats], 2)                          # (:164-165)
    dist_logits = flat[:, :4 * REG_MAX].permute(0, 2, 1)                                # [N, A, 64]
    cls_logits = flat[:, 4 * REG_MAX:].permute(0, 2, 1)                                 # [N, A, nc]
    pts, st = anchors_of(grids, strides)
    A = pts.shape[0]
    H, W = grids[0][0] * float(strides[0]), grids[0][1] * float(strides[0])             # (:172)
    ltrb = dist_logits.reshape(N, A, 4, REG_MAX).softmax(3).matmul(torch.arange(REG_MAX).float())
    boxes = torch.cat((pts - ltrb[..., :2], pts + ltrb[..., 2:]), -1)                   # xyxy, grid units (:160)
    t_score = torch.zeros(N, A, nc)
    t_box = torch.zeros(N, A, 4)
    fg_all = torch.zeros(N, A, dtype=torch.bool)
    with torch.no_grad():
        for b in range(N):
            mine = targets[targets[:, 0] == b]
            xy, wh = mine[:, 2:4] * torch.tensor([W, H]), mine[:, 4:6] * torch.tensor([W, H])
            gts = torch.cat((xy - wh / 2, xy + wh / 2), 1)
            gt_idx, fg, norm = assign_image(cls_logits[b].sigmoid(), boxes[b] * st, pts * st, mine[:, 1].long(), gts, *assigner)
            if mine.shape[0]:
                t_box[b] = gts[gt_idx] / st                                             # (:191)
                t_score[b, torch.arange(A), mine[gt_idx, 1].long()] = norm * fg
            fg_all[b] = fg
    tss = max(t_score.sum(), 1)                                                         # (:192)
    bce = F.binary_cross_entropy_with_logits(cls_logits, t_score, pos_weight=torch.tensor([cls_pw]), reduction="none")
    if fl_gamma > 0:                                                                    # FocalLoss, utils/loss_tal.py:32-60
        prob = cls_logits.sigmoid()
        p_t = t_score * prob + (1 - t_score) * (1 - prob)
        bce = bce * (t_score * 0.25 + (1 - t_score) * 0.75) * (1.0 - p_t) ** fl_gamma
    lcls = bce.sum() / tss
    lbox, ldfl = torch.zeros(()), torch.zeros(())
    if fg_all.any():
        w = t_score.sum(-1)[fg_all]
        lbox = ((1.0 - box_term_iou(boxes[fg_all], t_box[fg_all])) * w).sum() / tss          # (:74-81)
        pts_all = pts[None].expand(N, A, 2)[fg_all]
        tb = t_box[fg_all]
        tgt = torch.cat((pts_all - tb[:, :2], tb[:, 2:] - pts_all), 1).clamp(0, REG_MAX - 1 - 0.01)
        tl = tgt.long()
        wl = (tl + 1).float() - tgt
        lg = dist_logits[fg_all].reshape(-1, REG_MAX)
        ce_l = F.cross_entropy(lg, tl.reshape(-1), reduction="none").view(tl.shape)
        ce_r = F.cross_entropy(lg, tl.reshape(-1) + 1, reduction="none").view(tl.shape)
        ldfl = ((ce_l * wl + ce_r * (1 - wl)).mean(-1) * w).sum() / tss                 # (:84-103)
    items = torch.stack((lbox * GAINS[0], lcls * GAINS[1], ldfl * GAINS[2]))
    return items.sum() * N, items.detach(), int(fg_all.sum())
