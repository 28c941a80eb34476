"""Stack-B training loss: `ComputeLoss` with the reference's constructor and call signature
(utils/loss_tal.py:105-215: TaskAlignedAssigner + box term + DFL + BCE) on the device, forward and gradient in one
C-ABI call (`ecsy_tal_loss`): no per-image Python loop, no padded [batch, max labels, anchors] tensors, 9 launches.

    compute_loss = ComputeLoss(model)                       # reads model.hyp and the DDetect head
    loss, loss_items = compute_loss(pred, targets)          # pred: list of [N, 64 + nc, ny, nx]; targets [nt, 6] (GPU)
    loss.backward()

fl_gamma > 0 wraps the class BCE in FocalLoss like the reference.  The assigner's hyper-parameters (topk, alpha, beta) are read
from the YOLOM / YOLOA / YOLOB environment variables at construction like the reference (:134-137).  use_dfl=False raises
NotImplementedError: the reference's own DDetect head (reg_max = 16, 64 box channels) cannot be trained that way either --
without the DFL decode `bbox_decode` hands the 64-channel distribution to dist2bbox, which expects 4 (:103-113 of the
reference's ComputeLoss.bbox_decode).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import torch

from . import _cabi
from .functional import _chk_cuda, _st, _timed

REG_MAX = 16
GAINS = (7.5, 0.5, 1.5)      # box, cls, dfl (utils/loss_tal.py:210-212)
ASSIGNER = (10, 0.5, 6.0)    # TaskAlignedAssigner topk, alpha, beta (utils/loss_tal.py:134-137 defaults)


def tal_loss(feats: Sequence[torch.Tensor], targets: torch.Tensor, strides: Sequence[float], cls_pw: float = 1.0,
             gains: Sequence[float] = GAINS, need_grad: bool = True, fl_gamma: float = 0.0,
             assigner: Sequence[float] = ASSIGNER) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """-> (out [6] = (loss, box, cls, dfl, foreground anchors, target_scores.sum()), gradients per level or [])."""
    feats = [x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous() for x in feats]
    _chk_cuda(*feats, targets)
    nl = len(feats)
    if nl < 1 or any(x.dim() != 4 for x in feats):
        raise ValueError("expected a list of [N, 64 + nc, ny, nx] tensors")
    N, no = feats[0].shape[:2]
    if no <= 4 * REG_MAX or any(x.shape[0] != N or x.shape[1] != no for x in feats):
        raise ValueError("levels disagree on batch / outputs, or fewer than 64 + 1 channels")
    if len(strides) != nl:
        raise ValueError("one stride per level")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise ValueError(f"targets must be [nt, 6] = (image, class, x, y, w, h), got {tuple(targets.shape)}")
    dev = feats[0].device
    tg = targets.detach().to(device=dev, dtype=torch.float32).contiguous()
    nt = tg.shape[0]
    grads = [torch.empty_like(x) for x in feats] if need_grad else []
    out = torch.empty(6, device=dev, dtype=torch.float32)
    ny = (C.c_int * nl)(*[x.shape[2] for x in feats])
    nx = (C.c_int * nl)(*[x.shape[3] for x in feats])
    fp = (C.c_void_p * nl)(*[x.data_ptr() for x in feats])
    gp = (C.c_void_p * nl)(*[g.data_ptr() for g in grads]) if need_grad else None
    st = (C.c_float * nl)(*[float(s) for s in strides])
    L = _cabi.lib()
    ws = torch.empty(max(L.ecsy_tal_loss_ws_bytes(nl, N, nt, ny, nx), 256), device=dev, dtype=torch.uint8)
    with _timed("loss", 9 if nt else 7):
        _cabi.check(L.ecsy_tal_loss(fp, gp, tg.data_ptr() if nt else None, nt, nl, N, no - 4 * REG_MAX, ny, nx, st,
                                    float(cls_pw), float(gains[0]), float(gains[1]), float(gains[2]), float(fl_gamma),
                                    int(assigner[0]), float(assigner[1]), float(assigner[2]), out.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _st()), "tal_loss")
    return out, grads


class _TalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, targets, *feats):
        out, grads = tal_loss(feats, targets, need_grad=any(ctx.needs_input_grad[2:]), **cfg)
        ctx.save_for_backward(*grads)
        ctx.mark_non_differentiable(out)
        ctx.dtypes = [x.dtype for x in feats]
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        grads = ctx.saved_tensors
        if not grads:
            return (None, None) + (None,) * len(ctx.dtypes)
        return (None, None) + tuple((g * g_loss).to(dt) for g, dt in zip(grads, ctx.dtypes))


class ComputeLoss:
    """Drop-in for utils/loss_tal.py:105 `ComputeLoss`: `(loss, loss_items)` with loss_items = (box, cls, dfl)."""

    def __init__(self, model, use_dfl=True):
        h = model.hyp
        self.fl_gamma = float(h.get("fl_gamma", 0.0))        # > 0: FocalLoss around the class BCE (:116-119)
        if not use_dfl:
            raise NotImplementedError("use_dfl=False needs a 4-channel box head; the DDetect head of this path has reg_max = 16")
        self.assigner = (int(os.getenv('YOLOM', 10)), float(os.getenv('YOLOA', 0.5)), float(os.getenv('YOLOB', 6.0)))
        if self.assigner[0] < 1 or self.assigner[1] < 0 or self.assigner[2] < 0:
            raise ValueError(f"TaskAlignedAssigner hyper-parameters (YOLOM, YOLOA, YOLOB) = {self.assigner}")
        m = model.module if hasattr(model, 'module') and hasattr(model.module, 'model') else model
        m = m.model[-1]
        if getattr(m, "reg_max", REG_MAX) != REG_MAX:
            raise NotImplementedError("reg_max != 16")
        self.hyp = h
        self.stride = m.stride
        self.nc, self.nl, self.no, self.reg_max = m.nc, m.nl, m.no, m.reg_max
        self._strides = [float(s) for s in m.stride]          # host copy once: no synchronisation per step
        self.use_dfl = use_dfl

    def __call__(self, p, targets, img=None, epoch=0):
        feats = p[1] if isinstance(p, tuple) else p
        cfg = dict(strides=self._strides, cls_pw=self.hyp["cls_pw"], fl_gamma=max(self.fl_gamma, 0.0), assigner=self.assigner)
        loss, out = _TalLossFn.apply(cfg, targets, *feats)
        return loss, out[1:4]
