"""CPU oracle for the Stack-A training loss -- TEST INFRASTRUCTURE ONLY.

A plain-torch restatement of `ComputeLoss.__call__` / `build_targets` (utils/loss.py:162-290) on the path the shipped
hyper-parameter files select (data/hyps/hyp.scratch.yaml: fl_gamma 0, slide_ratio 0): SIoU box term
(utils/metrics.py:227-307 with SIoU=True), BCE-with-logits objectness / class terms with `pos_weight`
(utils/loss.py:138-139), label smoothing (:142), per-level balance (:156).  It is written around an explicit
candidate table -- (offset k, anchor a, target j), ordered k-major like `t.repeat((5, 1, 1))[j]` (:258-266) -- because
that is the formulation the CUDA kernels use; duplicates of an (image, anchor, cell) index resolve to the LAST
candidate in that order, which is what the reference's serial CPU `index_put_` (:205) does.
Gradients come from autograd through this restatement.  Pinned by tests/test_oracle_post.py against fixtures made by
the UNMODIFIED reference (oracle/gen_golden_loss.py -> tests/golden/post_loss.pt).  Only tests/, smoke() and
bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

# (dx, dy) of the five candidate cells: the target's own cell and the nearer horizontal / vertical neighbours
# (utils/loss.py:241-246, scaled by the bias g = 0.5)
_OFFSETS = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (-0.5, 0.0), (0.0, -0.5))
_G = 0.5
BALANCE = {3: [4.0, 1.0, 0.4]}                                   # utils/loss.py:156
BALANCE_DEFAULT = [4.0, 1.0, 0.25, 0.06, 0.02]


def smooth_bce(eps: float) -> Tuple[float, float]:
    """utils/loss.py:13-15"""
    return 1.0 - 0.5 * eps, 0.5 * eps


def candidates(targets: torch.Tensor, anchors_l: torch.Tensor, ny: int, nx: int, anchor_t: float) -> Dict[str, torch.Tensor]:
    """build_targets for one level (utils/loss.py:248-288) as a table over ALL 5 * na * nt candidates with a validity
    mask; rows are in the reference's output order (offset-major, then anchor, then target)."""
    na, nt = anchors_l.shape[0], targets.shape[0]
    gx = targets[:, 2] * nx                         # targets * gain, gain[2:6] = (nx, ny, nx, ny) (:250-253)
    gy = targets[:, 3] * ny
    gw = targets[:, 4] * nx
    gh = targets[:, 5] * ny
    rw = gw[None, :] / anchors_l[:, 0:1]            # [na, nt] (:256)
    rh = gh[None, :] / anchors_l[:, 1:2]
    fits = torch.maximum(torch.maximum(rw, 1 / rw), torch.maximum(rh, 1 / rh)) < anchor_t       # (:257)
    ix, iy = nx - gx, ny - gy                       # inverse coordinates (:263)
    near = torch.stack([torch.ones(nt, dtype=torch.bool),
                        (gx % 1 < _G) & (gx > 1), (gy % 1 < _G) & (gy > 1),
                        (ix % 1 < _G) & (ix > 1), (iy % 1 < _G) & (iy > 1)])                     # [5, nt] (:264-266)
    valid = (near[:, None, :] & fits[None, :, :]).reshape(-1)                                   # [5 * na * nt]
    off = torch.tensor(_OFFSETS)
    k = torch.arange(5).repeat_interleave(na * nt)
    a = torch.arange(na).repeat_interleave(nt).repeat(5)
    j = torch.arange(nt).repeat(5 * na)
    ci = (gx[j] - off[k, 0]).long().clamp(0, nx - 1)              # truncation, then the in-place clamp (:278-283)
    cj = (gy[j] - off[k, 1]).long().clamp(0, ny - 1)
    return dict(valid=valid, a=a, j=j, b=targets[j, 0].long(), c=targets[j, 1].long(), gi=ci, gj=cj,
                tbox=torch.stack([gx[j] - ci, gy[j] - cj, gw[j], gh[j]], 1))                     # (:284)


def siou(pb: torch.Tensor, tb: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """bbox_iou(pbox.T, tbox, x1y1x2y2=False, SIoU=True) (utils/metrics.py:236-254, 256-257, 286-307), rows = boxes."""
    px1, px2 = pb[:, 0] - pb[:, 2] / 2, pb[:, 0] + pb[:, 2] / 2
    py1, py2 = pb[:, 1] - pb[:, 3] / 2, pb[:, 1] + pb[:, 3] / 2
    tx1, tx2 = tb[:, 0] - tb[:, 2] / 2, tb[:, 0] + tb[:, 2] / 2
    ty1, ty2 = tb[:, 1] - tb[:, 3] / 2, tb[:, 1] + tb[:, 3] / 2
    inter = (torch.min(px2, tx2) - torch.max(px1, tx1)).clamp(0) * (torch.min(py2, ty2) - torch.max(py1, ty1)).clamp(0)
    w1, h1 = px2 - px1, py2 - py1 + eps
    w2, h2 = tx2 - tx1, ty2 - ty1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(px2, tx2) - torch.min(px1, tx1)
    ch = torch.max(py2, ty2) - torch.min(py1, ty1)
    sx = (tx1 + tx2 - px1 - px2) * 0.5 + eps
    sy = (ty1 + ty2 - py1 - py2) * 0.5 + eps
    sigma = torch.pow(sx ** 2 + sy ** 2, 0.5)
    s1, s2 = sx.abs() / sigma, sy.abs() / sigma
    s = torch.where(s1 > pow(2, 0.5) / 2, s2, s1)
    angle = torch.cos(torch.arcsin(s) * 2 - math.pi / 2)
    gamma = angle - 2
    dist = 2 - torch.exp(gamma * (sx / cw) ** 2) - torch.exp(gamma * (sy / ch) ** 2)
    ow = (w1 - w2).abs() / torch.max(w1, w2)
    oh = (h1 - h2).abs() / torch.max(h1, h2)
    shape = torch.pow(1 - torch.exp(-1 * ow), 4) + torch.pow(1 - torch.exp(-1 * oh), 4)
    return iou - torch.pow(0.5 * (dist + shape) + eps, 1)


def _slide_weight(true: torch.Tensor, ema: float) -> torch.Tensor:
    """SlideLoss.forward's modulating weight (utils/loss.py:61-69): 1 below ema - 0.1, exp(1 - ema) up to ema,
    exp(1 - true) from ema on."""
    w = torch.ones_like(true)
    w = torch.where((true > ema - 0.1) & (true < ema), torch.full_like(true, math.exp(1.0 - ema)), w)
    return torch.where(true >= ema, torch.exp(-(true - 1.0)), w)


def _elem_loss(x, t, pw, kind, *, focal_gamma=0.0, slide=None, which=None, auto_iou=0.5):
    """The wrapped criterion of utils/loss.py:138-152 with reduction 'mean': BCEWithLogits, optionally inside SlideLoss
    (:38-76, stateful: `slide[which]` is its `ema`, updated on every call with the new value weighted by alpha = 0.999
    as the reference does) or FocalLoss (:80-106, alpha 0.25)."""
    loss = F.binary_cross_entropy_with_logits(x, t, pos_weight=pw, reduction="none")
    if slide is not None:
        a = max(float(auto_iou), 0.2)
        ema = slide.get(which)
        ema = a if ema is None else float(torch.tensor(0.999, dtype=torch.float32) * torch.tensor(a, dtype=torch.float32)
                                            + torch.tensor(1 - 0.999) * torch.tensor(ema, dtype=torch.float32))
        slide[which] = ema
        loss = loss * _slide_weight(t, ema)
    elif focal_gamma > 0:
        prob = torch.sigmoid(x)
        p_t = t * prob + (1 - t) * (1 - prob)
        loss = loss * (t * 0.25 + (1 - t) * 0.75) * (1.0 - p_t) ** focal_gamma
    return loss.mean()


def compute_loss(p: Sequence[torch.Tensor], targets: torch.Tensor, anchors: torch.Tensor, hyp: dict, gr: float = 1.0,
                 slide_state: dict = None):
    """-> (loss [1], loss_items [3] = (lbox, lobj, lcls) detached, per-level match counts, per-level objectness BCE).
    p[i]: [N, na, ny, nx, 5 + nc] raw outputs, anchors [nl, na, 2] in grid units, targets [nt, 6].
    slide_state: the two SlideLoss EMAs ({'cls': .., 'obj': ..}, empty before the first call) when hyp['slide_ratio'] > 0."""
    nl, nc = len(p), p[0].shape[-1] - 5
    cp, cn = smooth_bce(hyp.get("label_smoothing", 0.0))
    use_slide = hyp.get("slide_ratio", 0.0) > 0
    gamma = hyp.get("fl_gamma", 0.0)
    if use_slide and gamma > 0:
        raise TypeError("the reference cannot combine SlideLoss and FocalLoss (FocalLoss.forward takes no auto_iou)")
    if use_slide and slide_state is None:
        raise ValueError("slide_ratio > 0 needs the caller's persistent slide_state dict")
    crit = dict(focal_gamma=gamma, slide=slide_state if use_slide else None)
    balance = BALANCE.get(nl, BALANCE_DEFAULT)
    cls_pw, obj_pw = torch.tensor([hyp["cls_pw"]]), torch.tensor([hyp["obj_pw"]])
    lbox, lobj, lcls = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    counts, objs = [], []
    for i, pi in enumerate(p):
        N, na, ny, nx, _ = pi.shape
        tobj = torch.zeros(N, na, ny, nx)
        cand = candidates(targets.float(), anchors[i].float(), ny, nx, hyp["anchor_t"])
        sel = cand["valid"].nonzero().flatten()
        n = int(sel.numel())
        counts.append(n)
        if n:
            b, a, gj, gi = cand["b"][sel], cand["a"][sel], cand["gj"][sel], cand["gi"][sel]
            ps = pi[b, a, gj, gi]
            pxy = ps[:, :2].sigmoid() * 2 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anchors[i][a]
            iou = siou(torch.cat((pxy, pwh), 1), cand["tbox"][sel])
            lbox = lbox + (1.0 - iou).mean()
            auto_iou = float(iou.detach().mean())
            score = iou.detach().clamp(0)
            flat = ((b * na + a) * ny + gj) * nx + gi
            last = torch.full((tobj.numel(),), -1, dtype=torch.long).scatter_reduce(
                0, flat, torch.arange(n), reduce="amax", include_self=True)
            won = last[flat] == torch.arange(n)                    # the last writer of each cell (:205)
            tobj.view(-1)[flat[won]] = (1.0 - gr) + gr * score[won]
            if nc > 1:
                t = torch.full_like(ps[:, 5:], cn)
                t[torch.arange(n), cand["c"][sel]] = cp
                lcls = lcls + _elem_loss(ps[:, 5:], t, cls_pw, "cls", which="cls", auto_iou=auto_iou, **crit)
        obji = _elem_loss(pi[..., 4], tobj, obj_pw, "obj", which="obj", auto_iou=auto_iou if n else 0.5, **crit)
        objs.append(float(obji.detach()))
        lobj = lobj + obji * balance[i]
    lbox = lbox * hyp["box"]
    lobj = lobj * hyp["obj"]
    lcls = lcls * hyp["cls"]
    bs = p[0].shape[0]
    return (lbox + lobj + lcls) * bs, torch.cat((lbox, lobj, lcls)).detach(), counts, objs
