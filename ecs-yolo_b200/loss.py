"""Stack-A training loss: `ComputeLoss` with the reference's constructor and call signature
(utils/loss.py:130-234) on the device, forward and gradient in one C-ABI call (`ecsy_yolo_loss`): no boolean-mask
round trips to the host, no per-level index lists, ~10 launches instead of ~150.

    compute_loss = ComputeLoss(model)                 # reads model.hyp and the Detect head, like utils/loss.py:131-160
    loss, loss_items = compute_loss(pred, targets)    # pred: list of [N, na, ny, nx, 5 + nc]; targets [nt, 6] on the GPU
    loss.backward()                                   # the gradient was computed with the forward

Covers the path the shipped hyper-parameter files select (data/hyps/hyp.scratch*.yaml: fl_gamma = 0, slide_ratio = 0):
SIoU box term, BCE-with-logits objectness / class terms with pos_weight, label smoothing, per-level balance,
autobalance -- and the two wrappers the reference can put around the BCE terms: FocalLoss (fl_gamma > 0) and the
stateful SlideLoss (slide_ratio > 0; its two EMAs live in a device tensor owned by this object).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from .functional import _chk_cuda, _st, _timed

BALANCE = {3: [4.0, 1.0, 0.4]}                     # utils/loss.py:156
BALANCE_DEFAULT = [4.0, 1.0, 0.25, 0.06, 0.02]


def smooth_BCE(eps=0.1):
    """utils/loss.py:13-15"""
    return 1.0 - 0.5 * eps, 0.5 * eps


def yolo_loss(p: Sequence[torch.Tensor], targets: torch.Tensor, anchors: torch.Tensor, balance: Sequence[float],
              box: float, obj: float, cls: float, cls_pw: float = 1.0, obj_pw: float = 1.0, cp: float = 1.0,
              cn: float = 0.0, anchor_t: float = 4.0, gr: float = 1.0, need_grad: bool = True, fl_gamma: float = 0.0,
              slide_state: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """-> (out [4 + nl] = (loss, lbox, lobj, lcls, objectness BCE mean per level), grads per level or []).
    Tensor-level entry over `ecsy_yolo_loss`; the gradients are for an upstream gradient of 1.
    fl_gamma > 0: FocalLoss around the BCE terms; slide_state: a zero-initialised fp32 CUDA tensor [4] that persists
    across calls (the two SlideLoss EMAs + has-value flags) -- passing it selects SlideLoss and updates it in place."""
    p = [x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous() for x in p]
    _chk_cuda(*p, targets, anchors)
    nl = len(p)
    if nl < 1 or any(x.dim() != 5 for x in p):
        raise ValueError("expected a list of [N, na, ny, nx, 5 + nc] tensors")
    N, na, _, _, no = p[0].shape
    if any(x.shape[0] != N or x.shape[1] != na or x.shape[4] != no for x in p):
        raise ValueError("levels disagree on batch / anchors / outputs")
    if tuple(anchors.shape) != (nl, na, 2):
        raise ValueError(f"anchors {tuple(anchors.shape)} != {(nl, na, 2)}")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise ValueError(f"targets must be [nt, 6] = (image, class, x, y, w, h), got {tuple(targets.shape)}")
    dev = p[0].device
    if slide_state is not None:
        _chk_cuda(slide_state)
        if slide_state.dtype != torch.float32 or slide_state.numel() != 4 or not slide_state.is_contiguous():
            raise ValueError("slide_state must be a contiguous fp32 tensor of 4 elements")
        if fl_gamma > 0:
            raise TypeError("SlideLoss and FocalLoss cannot be combined (FocalLoss.forward takes no auto_iou in the reference)")
    tg = targets.detach().to(device=dev, dtype=torch.float32).contiguous()
    an = anchors.detach().to(device=dev, dtype=torch.float32).contiguous()
    nt = tg.shape[0]
    grads = [torch.empty_like(x) for x in p] if need_grad else []
    out = torch.empty(4 + nl, device=dev, dtype=torch.float32)
    ny = (C.c_int * nl)(*[x.shape[2] for x in p])
    nx = (C.c_int * nl)(*[x.shape[3] for x in p])
    pp = (C.c_void_p * nl)(*[x.data_ptr() for x in p])
    gp = (C.c_void_p * nl)(*[g.data_ptr() for g in grads]) if need_grad else None
    bal = (C.c_float * nl)(*[float(b) for b in balance[:nl]])
    L = _cabi.lib()
    ws = torch.empty(max(L.ecsy_yolo_loss_ws_bytes(nl, N, na, nt, ny, nx), 256), device=dev, dtype=torch.uint8)
    with _timed("loss", (4 if nt else 1) + nl + (1 if slide_state is not None else 0)):
        _cabi.check(L.ecsy_yolo_loss(pp, gp, tg.data_ptr() if nt else None, nt, an.data_ptr(), nl, N, na, no - 5, ny, nx,
                                     bal, float(box), float(obj), float(cls), float(cls_pw), float(obj_pw), float(cp),
                                     float(cn), float(anchor_t), float(gr), float(fl_gamma),
                                     slide_state.data_ptr() if slide_state is not None else None, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _st()), "yolo_loss")
    return out, grads


class _YoloLossFn(torch.autograd.Function):
    """loss = f(p_0 .. p_{nl-1}); the gradient is produced by the forward call and scaled by the upstream gradient."""

    @staticmethod
    def forward(ctx, cfg, targets, anchors, *p):
        need = any(ctx.needs_input_grad[3:])
        out, grads = yolo_loss(p, targets, anchors, need_grad=need, **cfg)
        ctx.save_for_backward(*grads)
        ctx.mark_non_differentiable(out)
        ctx.dtypes = [x.dtype for x in p]
        return out[0:1].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        grads = ctx.saved_tensors
        if not grads:
            return (None, None, None) + (None,) * len(ctx.dtypes)
        return (None, None, None) + tuple((g * g_loss).to(dt) for g, dt in zip(grads, ctx.dtypes))


class ComputeLoss:
    """Drop-in for utils/loss.py:130 `ComputeLoss` (same constructor, attributes and `(loss, loss_items)` result)."""

    def __init__(self, model, autobalance=False):
        self.sort_obj_iou = False
        h = model.hyp
        self.slide_ratio = h.get('slide_ratio', 0.0)                              # utils/loss.py:144-147
        self.fl_gamma = h.get('fl_gamma', 0.0)                                    # utils/loss.py:149-152
        if self.slide_ratio > 0 and self.fl_gamma > 0:
            raise TypeError("slide_ratio > 0 together with fl_gamma > 0: the reference's FocalLoss.forward takes no auto_iou "
                            "(utils/loss.py:212 would raise); use one of the two wrappers")
        self._slide_state = None          # the two SlideLoss EMAs live on the device, created at the first call
        self.cp, self.cn = smooth_BCE(eps=h.get('label_smoothing', 0.0))          # utils/loss.py:142
        m = model.module if hasattr(model, 'module') and hasattr(model.module, 'model') else model   # de-parallel (:155)
        det = m.model[-1]
        self.balance = list(BALANCE.get(det.nl, BALANCE_DEFAULT))
        self.ssi = list(det.stride).index(16) if autobalance else 0
        self.gr, self.hyp, self.autobalance = 1.0, h, autobalance
        for k in 'na', 'nc', 'nl', 'anchors':
            setattr(self, k, getattr(det, k))

    def __call__(self, p, targets):
        h = self.hyp
        cfg = dict(balance=list(self.balance), box=h['box'], obj=h['obj'], cls=h['cls'], cls_pw=h['cls_pw'],
                   obj_pw=h['obj_pw'], cp=self.cp, cn=self.cn, anchor_t=h['anchor_t'], gr=self.gr,
                   fl_gamma=self.fl_gamma if self.fl_gamma > 0 else 0.0)
        if self.slide_ratio > 0:
            if self._slide_state is None or self._slide_state.device != p[0].device:
                self._slide_state = torch.zeros(4, device=p[0].device, dtype=torch.float32)
            cfg["slide_state"] = self._slide_state
        loss, out = _YoloLossFn.apply(cfg, targets, self.anchors, *p)
        if self.autobalance:                                                      # utils/loss.py:224-228 (host read)
            obji = out[4:4 + self.nl].tolist()
            self.balance = [b * 0.9999 + 0.0001 / o for b, o in zip(self.balance, obji)]
            self.balance = [x / self.balance[self.ssi] for x in self.balance]
        return loss, out[1:4]
