"""Training-step tail of the hot path: the reference's SGD-Nesterov optimizer with its three
parameter groups (train.py:259-287) and `ModelEMA.update` (utils/torch_utils.py:285-316) as ONE multi-tensor kernel
launch per step (`ecsy_sgd_ema_step`) instead of ~4 launches per tensor plus a Python loop over the state_dict.

    opt = SGDNesterovEMA(model, lr=hyp['lr0'], momentum=hyp['momentum'], weight_decay=hyp['weight_decay'])
    for x['lr'] in opt.param_groups: ...        # the reference's warm-up / scheduler code works unchanged
    loss.backward(); opt.step(); opt.zero_grad()  # step() also advances opt.ema (a ModelEMA-shaped object)
"""
from __future__ import annotations

import math
from copy import deepcopy
from typing import List, Optional

import torch
import torch.nn as nn

from . import _cabi
from . import functional as F_
from .functional import _p, _st, _timed


def param_groups_of(model: nn.Module):
    """train.py:259-270: g0 = BatchNorm3d weights (no decay), g1 = other weights (decay), g2 = biases."""
    g0, g1, g2 = [], [], []
    for v in model.modules():
        if hasattr(v, 'bias') and isinstance(v.bias, nn.Parameter):
            g2.append(v.bias)
        if isinstance(v, (nn.BatchNorm3d, nn.SyncBatchNorm)):   # --sync-bn converts after the reference built its groups (train.py:283, :359)
            g0.append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, nn.Parameter):
            g1.append(v.weight)
    return g0, g1, g2


class _EMA:
    """Attribute-compatible with utils.torch_utils.ModelEMA (`.ema`, `.updates`, `.decay`)."""

    def __init__(self, model: nn.Module, decay: float = 0.9999, updates: int = 0):
        m = model.module if hasattr(model, "module") and isinstance(model.module, nn.Module) else model
        self.ema = deepcopy(m).eval()
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)


class SGDNesterovEMA:
    def __init__(self, model: nn.Module, lr: float, momentum: float = 0.937, weight_decay: float = 0.0,
                 nesterov: bool = True, ema: bool = True, ema_decay: float = 0.9999, updates: int = 0):
        m = model.module if hasattr(model, "module") and isinstance(model.module, nn.Module) else model
        self.model = m
        g0, g1, g2 = param_groups_of(m)
        self.param_groups = [dict(params=g0, lr=lr, weight_decay=0.0, momentum=momentum, nesterov=nesterov),
                             dict(params=g1, lr=lr, weight_decay=weight_decay, momentum=momentum, nesterov=nesterov),
                             dict(params=g2, lr=lr, weight_decay=0.0, momentum=momentum, nesterov=nesterov)]
        self.momentum, self.nesterov = momentum, nesterov
        self.ema: Optional[_EMA] = _EMA(m, ema_decay, updates) if ema else None
        self._tables = None

    # ---- tables (built once; gradient addresses are refreshed every step) -------------------------------------
    def _build(self):
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("SGDNesterovEMA needs the model on a CUDA device: there is no CPU fallback")
        gid = {id(p): j for j, g in enumerate(self.param_groups) for p in g["params"]}
        sd = self.model.state_dict(keep_vars=True)
        esd = self.ema.ema.state_dict(keep_vars=True) if self.ema is not None else {}
        vals: List[torch.Tensor] = []
        groups, emas, is_param = [], [], []
        seen = set()
        for k, v in sd.items():
            if not v.dtype.is_floating_point or id(v) in seen:
                continue
            seen.add(id(v))
            if v.dtype != torch.float32 or not v.is_contiguous():
                raise RuntimeError(f"SGDNesterovEMA: '{k}' must be a contiguous fp32 tensor (fp32 master weights)")
            in_group = id(v) in gid
            if not in_group and (self.ema is None or k not in esd):
                continue
            vals.append(v)
            groups.append(gid.get(id(v), 0))
            is_param.append(in_group)
            emas.append(esd[k].data_ptr() if k in esd else 0)
        self._vals, self._is_param = vals, is_param
        self._moms = [torch.zeros_like(v, memory_format=torch.contiguous_format) if ip else None
                      for v, ip in zip(vals, is_param)]   # zero buffers: momentum*0 + d == the clone torch makes first
        chunk = _cabi.lib().ecsy_optim_chunk()
        ct, co = [], []
        for i, v in enumerate(vals):
            for off in range(0, v.numel(), chunk):
                ct.append(i)
                co.append(off)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        self._tables = dict(
            val=i64([v.data_ptr() for v in vals]), mom=i64([m.data_ptr() if m is not None else 0 for m in self._moms]),
            ema=i64(emas), numel=i64([v.numel() for v in vals]),
            group=torch.tensor(groups, dtype=torch.int32, device=dev),
            chunk_tensor=torch.tensor(ct, dtype=torch.int32, device=dev), chunk_off=i64(co), n_chunks=len(ct))
        # two pinned staging buffers + upload events: the host may run a step ahead of the GPU, so the table the
        # in-flight H2D copy of step n reads must not be rewritten by step n+1 (the buffer of step n-1 is reused only
        # after ITS copy has executed)
        self._grad_host = [torch.zeros(len(vals), dtype=torch.int64).pin_memory() for _ in range(2)]
        self._grad_dev = [torch.zeros(len(vals), dtype=torch.int64, device=dev) for _ in range(2)]
        self._grad_evt = [None, None]
        self._flip = 0
        self._ema_vals = [esd[k] for k in sd if k in esd and esd[k].dtype.is_floating_point]

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, update_ema: bool = True):
        """p <- SGD-Nesterov(p, p.grad) for every parameter with a gradient, then (update_ema) the EMA of every
        floating-point state_dict entry, in one launch.  Reference order: optimizer.step() ... ema.update(model)
        (train.py:576-582): the EMA sees the updated parameters."""
        if self._tables is None:
            self._build()
        t = self._tables
        ptrs = []
        for v, ip in zip(self._vals, self._is_param):
            g = v.grad if ip else None
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.device != v.device):
                raise RuntimeError("SGDNesterovEMA: gradients must be contiguous fp32 tensors on the parameter's device")
            ptrs.append(g.data_ptr() if g is not None else 0)
        b = self._flip
        self._flip ^= 1
        if self._grad_evt[b] is not None:
            self._grad_evt[b].synchronize()
        self._grad_host[b].copy_(torch.tensor(ptrs, dtype=torch.int64))
        self._grad_dev[b].copy_(self._grad_host[b], non_blocking=True)
        self._grad_evt[b] = torch.cuda.Event()
        self._grad_evt[b].record()
        grad_dev = self._grad_dev[b]
        d = 0.0
        do_ema = bool(update_ema and self.ema is not None)
        if do_ema:
            self.ema.updates += 1
            d = self.ema.decay(self.ema.updates)
        n_g = len(self.param_groups)
        import ctypes as C
        lrs = (C.c_float * n_g)(*[float(g["lr"]) for g in self.param_groups])
        wds = (C.c_float * n_g)(*[float(g["weight_decay"]) for g in self.param_groups])
        with _timed("sgd_ema", 1):
            _cabi.check(_cabi.lib().ecsy_sgd_ema_step(
                _p(t["val"]), _p(grad_dev), _p(t["mom"]), _p(t["ema"]), _p(t["numel"]), _p(t["group"]),
                len(self._vals), _p(t["chunk_tensor"]), _p(t["chunk_off"]), t["n_chunks"], lrs, wds, n_g,
                float(self.momentum), 1 if self.nesterov else 0, 0, 1 if do_ema else 0, float(d), float(1 - d), _st()),
                "sgd_ema_step")
        # The kernel writes parameters / EMA tensors through raw pointers: tell torch (version counters) and the
        # derived-weight caches (packed bf16 conv / spread weights, folded tdBN affines keyed on data_ptr + _version)
        # that every one of them changed.
        torch._C._increment_version(self._vals)
        if do_ema and self._ema_vals:
            torch._C._increment_version(self._ema_vals)
        F_.bump_weights_epoch()
