"""PyTorch custom operators (``torch.library``) over the C ABI: ``torch.ops.ecsy.*``.

The drop-in modules (``common.py``) call ``functional`` directly and chain whole blocks in hand-written
``autograd.Function``s; these operators expose the same kernels as first-class PyTorch ops -- schema, fake
(meta) implementations for shape inference / export, and an autograd formula for the differentiable unit of the
path -- so that code outside this package (a converted reference model, ``torch.export``, a custom head) can call the
B200 kernels without touching ctypes.  Tensors use the internal layout: real activations fp32 NHWC
``[Tp, N, H, W, C]`` (Tp == T, or 1 for a T-broadcast tensor), spikes bit-packed int32 ``[T, N, H, W, C/32]``.
CUDA only: there is no CPU kernel behind any of them (the fake implementations only compute shapes).

    bits = torch.ops.ecsy.lif_ecs(x, T, dw_w, dw_b, pw_w, pw_b, scale, shift, 5.0, 0.75, 0.25)
    y    = torch.ops.ecsy.spike_conv(bits, Cin, weight, scale, shift, residual, stride, pad)
    y    = torch.ops.ecsy.lif_spike_conv(x, T, dw_w, dw_b, pw_w, pw_b, weight, stride, pad, 5.0, 0.75, 0.25)  # autograd
    mean, var = torch.ops.ecsy.tdbn_stats(y)
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F_
from .functional import Act, Spikes

_wcache: dict = {}


def _cached(kind: str, tensors, builder):
    """Derived-weight cache of the raw-parameter operators.  The key holds data_ptr / _version / shape; because the caching
    allocator hands the address of a freed parameter to the next model, an entry is only valid while the tensors it was
    built from are still alive (weak references) -- a stale hit would silently run another model's weights."""
    import weakref
    key = (kind, F_.get_splits(), F_._state["conv_ts"], F_.weights_epoch()) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
    ent = _wcache.get(key)
    if ent is not None and all(r() is not None and r().data_ptr() == t.data_ptr() for r, t in zip(ent[0], tensors)):
        return ent[1]
    if len(_wcache) > 512:
        _wcache.clear()
    with torch.no_grad():
        v = builder()
    # the operator may see a fresh alias of the caller's tensor: keep the storage's owner alive-ness via the base object
    _wcache[key] = ([weakref.ref(t._base if t._base is not None else t) for t in tensors], v)
    return v


def _lif_w(dw_w, dw_b, pw_w, pw_b):
    return _cached("lif", (dw_w, dw_b, pw_w, pw_b), lambda: F_.make_lif_w(dw_w, dw_b, pw_w, pw_b))


def _conv_w(weight, stride, pad):
    return _cached(f"conv{stride}.{pad}", (weight,), lambda: F_.make_conv_w(weight, None, stride, pad, 1, True, False))


def _conv_out(h: int, w: int, k: int, stride: int, pad: int) -> Tuple[int, int]:
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


# ---------------------------------------------------------------------------------------------- lif_ecs
@torch.library.custom_op("ecsy::lif_ecs", mutates_args=())
def lif_ecs(x: Tensor, T: int, dw_w: Tensor, dw_b: Tensor, pw_w: Tensor, pw_b: Tensor, scale: Optional[Tensor],
            shift: Optional[Tensor], ecs_tau: float, alpha: float, beta: float) -> Tensor:
    """mem_update.forward (models/common.py:252-283) -> bit-packed spikes."""
    aff = (scale, shift) if scale is not None else None
    return F_.lif_ecs(Act(x.contiguous(), T), _lif_w(dw_w, dw_b, pw_w, pw_b), aff, ecs_tau, alpha, beta).bits


@lif_ecs.register_fake
def _(x, T, dw_w, dw_b, pw_w, pw_b, scale, shift, ecs_tau, alpha, beta):
    _, N, H, W, C = x.shape
    torch._check(C % 64 == 0, lambda: "ecsy::lif_ecs: C must be a multiple of 64")
    return x.new_empty((T, N, H, W, C // 32), dtype=torch.int32)


# ---------------------------------------------------------------------------------------------- module-facing operators
# The drop-in modules keep their derived weights in a per-module cache (common._cached) and hand them over explicitly: no
# hidden state in the operator.
@torch.library.custom_op("ecsy::lif_ecs_w", mutates_args=())
def lif_ecs_w(x: Tensor, T: int, dw_w9: Tensor, dw_b: Tensor, pw_packed: Tensor, pw_b: Tensor, w_eff: Optional[Tensor],
              bconst: Optional[Tensor], w_wave: Optional[Tensor], splits: int, scale: Optional[Tensor], shift: Optional[Tensor],
              ecs_tau: float, alpha: float, beta: float) -> Tensor:
    """mem_update.forward (models/common.py:252-283) on prepacked spread weights (functional.make_lif_w) -> spike bits."""
    w = F_.LifW(dw_w9, dw_b, pw_packed, pw_b, splits, w_eff, bconst, w_wave)
    aff = (scale, shift) if scale is not None else None
    return F_.lif_ecs(Act(x, T), w, aff, ecs_tau, alpha, beta).bits


@lif_ecs_w.register_fake
def _(x, T, dw_w9, dw_b, pw_packed, pw_b, w_eff, bconst, w_wave, splits, scale, shift, ecs_tau, alpha, beta):
    _, N, H, W, C = x.shape
    torch._check(C % 64 == 0, lambda: "ecsy::lif_ecs_w: C must be a multiple of 64")
    return x.new_empty((T, N, H, W, C // 32), dtype=torch.int32)


@torch.library.custom_op("ecsy::spike_conv_w", mutates_args=())
def spike_conv_w(bits: Tensor, cin: int, packed: Tensor, packed_ts: Optional[Tensor], splits: int, cout: int, k: int, stride: int,
                 pad: int, scale: Optional[Tensor], shift: Optional[Tensor], residual: Optional[Tensor]) -> Tensor:
    """Snn_Conv2d on spikes (models/common.py:609-624) on prepacked weights (functional.make_conv_w), folded tdBN, shortcut."""
    T = bits.shape[0]
    w = F_.ConvW(packed, None, None, cout, cin, k, stride, pad, 1, splits, False, packed_ts)
    res = Act(residual, T) if residual is not None else None
    return F_.spike_conv(Spikes(bits, cin), w, scale, shift, res).data


@spike_conv_w.register_fake
def _(bits, cin, packed, packed_ts, splits, cout, k, stride, pad, scale, shift, residual):
    T, N, H, W, _ = bits.shape
    Ho, Wo = _conv_out(H, W, k, stride, pad)
    return bits.new_empty((T, N, Ho, Wo, cout), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------- spike_conv
@torch.library.custom_op("ecsy::spike_conv", mutates_args=())
def spike_conv(bits: Tensor, cin: int, weight: Tensor, scale: Optional[Tensor], shift: Optional[Tensor],
               residual: Optional[Tensor], stride: int, pad: int) -> Tensor:
    """Snn_Conv2d on spikes (models/common.py:609-624) with folded tdBN affine and shortcut add."""
    T = bits.shape[0]
    res = Act(residual.contiguous(), T) if residual is not None else None
    return F_.spike_conv(Spikes(bits.contiguous(), cin), _conv_w(weight, stride, pad), scale, shift, res).data


@spike_conv.register_fake
def _(bits, cin, weight, scale, shift, residual, stride, pad):
    T, N, H, W, _ = bits.shape
    Ho, Wo = _conv_out(H, W, weight.shape[2], stride, pad)
    return weight.new_empty((T, N, Ho, Wo, weight.shape[0]), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------- tdbn_stats
@torch.library.custom_op("ecsy::tdbn_stats", mutates_args=())
def tdbn_stats(y: Tensor) -> Tuple[Tensor, Tensor]:
    """Per-channel mean / biased variance over (T, N, H, W) (models/common.py:668-700)."""
    return F_.bn_stats(Act(y.contiguous(), y.shape[0]))


@tdbn_stats.register_fake
def _(y):
    C = y.shape[-1]
    return y.new_empty((C,)), y.new_empty((C,))


# ---------------------------------------------------------------------------------------------- lif -> conv (autograd)
@torch.library.custom_op("ecsy::lif_spike_conv", mutates_args=())
def lif_spike_conv(x: Tensor, T: int, dw_w: Tensor, dw_b: Tensor, pw_w: Tensor, pw_b: Tensor, weight: Tensor, stride: int,
                   pad: int, ecs_tau: float, alpha: float, beta: float) -> Tuple[Tensor, Tensor]:
    """The differentiable unit of the path: ECS-LIF -> Snn_Conv2d (raw output, no tdBN).  Returns (y, bits); bits
    (1 bit per element) is what the backward keeps instead of the membranes, which it recomputes."""
    sp = F_.lif_ecs(Act(x.contiguous(), T), _lif_w(dw_w, dw_b, pw_w, pw_b), None, ecs_tau, alpha, beta)
    y = F_.spike_conv(sp, _conv_w(weight, stride, pad))
    return y.data, sp.bits


@lif_spike_conv.register_fake
def _(x, T, dw_w, dw_b, pw_w, pw_b, weight, stride, pad, ecs_tau, alpha, beta):
    _, N, H, W, C = x.shape
    Ho, Wo = _conv_out(H, W, weight.shape[2], stride, pad)
    return (x.new_empty((T, N, Ho, Wo, weight.shape[0])), x.new_empty((T, N, H, W, C // 32), dtype=torch.int32))


def _lsc_setup(ctx, inputs, output):
    x, T, dw_w, dw_b, pw_w, pw_b, weight, stride, pad, ecs_tau, alpha, beta = inputs
    ctx.save_for_backward(x, dw_w, dw_b, pw_w, pw_b, weight, output[1])
    ctx.cfg = (T, stride, pad, ecs_tau, alpha, beta)


def _lsc_backward(ctx, g_y, _g_bits):
    """Surrogate-gradient BPTT through the neuron (forward recomputed), dgrad / wgrad of the conv on tcgen05."""
    x, dw_w, dw_b, pw_w, pw_b, weight, bits = ctx.saved_tensors
    T, stride, pad, ecs_tau, alpha, beta = ctx.cfg
    with torch.no_grad():
        C, k = x.shape[-1], weight.shape[2]
        sp = Spikes(bits, C)
        g_y = g_y.contiguous()
        g_w = F_.spike_conv_wgrad(g_y, sp, k, stride, pad)
        splits = F_.get_splits()
        wT = _cached("dgrad", (weight,), lambda: F_.pack_dgrad_weight(weight, splits))
        g_s = F_.conv_dgrad(g_y, wT, splits, sp.H, sp.W, C, k, stride, pad)
        g_x, g_dw, g_db, g_pw, g_pb = F_.lif_ecs_bwd(g_s, Act(x.contiguous(), T), _lif_w(dw_w, dw_b, pw_w, pw_b), pw_w,
                                                     None, ecs_tau, alpha, beta)
        if x.shape[0] != T:   # T-broadcast input: the gradient of the single stored frame is the sum over T
            g_x = g_x.sum(0, keepdim=True)
    return g_x, None, g_dw, g_db, g_pw, g_pb, g_w, None, None, None, None, None


lif_spike_conv.register_autograd(_lsc_backward, setup_context=_lsc_setup)


# ---------------------------------------------------------------------------------------------- training losses
# Forward AND gradient in one call: the gradient w.r.t. every level is an explicit output (for an
# upstream gradient of 1) -- `loss.ComputeLoss` / `loss_tal.ComputeLoss` wrap these in an autograd.Function that scales it.
@torch.library.custom_op("ecsy::yolo_loss", mutates_args=())
def yolo_loss(p: List[Tensor], targets: Tensor, anchors: Tensor, balance: List[float], box: float, obj: float, cls: float,
              cls_pw: float, obj_pw: float, cp: float, cn: float, anchor_t: float, gr: float) -> Tuple[Tensor, List[Tensor]]:
    """utils/loss.py:162-290 (`ComputeLoss.__call__` + `build_targets`, SIoU + BCE) -> (out [4 + nl], d loss / d p)."""
    from .loss import yolo_loss as _impl
    out, grads = _impl(p, targets, anchors, balance, box, obj, cls, cls_pw, obj_pw, cp, cn, anchor_t, gr, True)
    return out, grads


@yolo_loss.register_fake
def _(p, targets, anchors, balance, box, obj, cls, cls_pw, obj_pw, cp, cn, anchor_t, gr):
    torch._check(all(x.dim() == 5 for x in p), lambda: "ecsy::yolo_loss: levels are [N, na, ny, nx, 5 + nc]")
    torch._check(targets.dim() == 2 and targets.shape[1] == 6, lambda: "ecsy::yolo_loss: targets are [nt, 6]")
    return p[0].new_empty((4 + len(p),), dtype=torch.float32), [torch.empty_like(x, dtype=torch.float32) for x in p]


@torch.library.custom_op("ecsy::tal_loss", mutates_args=())
def tal_loss(feats: List[Tensor], targets: Tensor, strides: List[float], cls_pw: float, gain_box: float, gain_cls: float,
             gain_dfl: float) -> Tuple[Tensor, List[Tensor]]:
    """utils/loss_tal.py:162-215 (TaskAlignedAssigner + box + DFL + BCE) -> (out [6], d loss / d feats)."""
    from .loss_tal import tal_loss as _impl
    out, grads = _impl(feats, targets, strides, cls_pw, (gain_box, gain_cls, gain_dfl), True)
    return out, grads


@tal_loss.register_fake
def _(feats, targets, strides, cls_pw, gain_box, gain_cls, gain_dfl):
    torch._check(all(x.dim() == 4 and x.shape[1] > 64 for x in feats),
                 lambda: "ecsy::tal_loss: levels are [N, 64 + nc, ny, nx]")
    torch._check(len(strides) == len(feats), lambda: "ecsy::tal_loss: one stride per level")
    return feats[0].new_empty((6,), dtype=torch.float32), [torch.empty_like(x, dtype=torch.float32) for x in feats]
