"""Drop-in nn.Module classes for the spiking hot path of ECS-YOLO.

Same class names, constructor signatures and ``state_dict`` keys as the reference's
``models/common.py`` (mem_update :236, Snn_Conv2d :593, batch_norm_2d :668, batch_norm_2d1 :682,
BatchNorm3d1 :694, BatchNorm3d2 :753, Conv :362, Conv_B :393, Conv_1 :409, Conv_2 :428, Conv_7 :549,
Sample :844, BasicBlock_1 :1049, BasicBlock_2 :1182, Concat_res2 :1454, Concat :1758, DFL :312), so the
reference's YAML files and state dicts load unchanged.  The forwards call the sm_100a kernels of
libecsy.so through ``functional``; tensors crossing a module boundary are reference-shaped
``[T, N, C, H, W]`` fp32 views over NHWC memory, so chains of these modules never convert layouts.

Known, deliberate differences from the reference: T is read from the tensor, not from a
module-level global; ``mem_update.spread`` is created when the owning block is constructed (the
channel count is known there) instead of on the first forward; the backward is a manual chain of C-ABI
calls per block (``autograd.py``) instead of autograd through every time-loop op.
"""
# NOTE: `Conv.conv` sees a REAL input (models/common.py:372), so it takes the im2col + tcgen05 path.
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import functional as F_
from .functional import Act, Spikes

time_window = 4  # models/common.py:40; used by Model.forward to replicate the image

_caches: "weakref.WeakKeyDictionary[nn.Module, dict]" = weakref.WeakKeyDictionary()


def _cached(mod: nn.Module, key: str, tensors, builder):
    ver = tuple((t.data_ptr(), t._version) for t in tensors if t is not None) + (F_.get_splits(), F_.weights_epoch())
    c = _caches.setdefault(mod, {})
    ent = c.get(key)
    if ent is not None and ent[0] == ver:
        return ent[1]
    with torch.no_grad():
        val = builder()
    c[key] = (ver, val)
    return val


def autopad(k, p=None):
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def _tag_spikes(t: torch.Tensor, sp: Spikes) -> torch.Tensor:
    t._ecsy_spikes = sp  # consumed by Snn_Conv2d.forward so module-by-module use stays bit-packed
    return t


class mem_update(nn.Module):
    """ECS-LIF neuron (models/common.py:236-309)."""

    def __init__(self, act=False, ecs_tau: float = 5., alpha: float = 0.75, beta: float = 0.25, ECS=False):
        super().__init__()
        self.ECS = ECS
        self.actFun = nn.SiLU()
        self.act = act
        self.alpha = alpha
        self.beta = beta
        self.ecs_tau = ecs_tau
        self.spread = None

    def InitEcsSpread(self, x: torch.Tensor):
        if x.ndim != 4:
            raise NotImplementedError("only [N,C,H,W] timesteps are supported")
        self._init_spread(x.shape[1], x.device)

    def _init_spread(self, C: int, device=None):
        if self.spread is None:
            self.spread = nn.Sequential(nn.Conv2d(C, C, kernel_size=3, padding=1, groups=C, device=device),
                                        nn.Conv2d(C, C, kernel_size=1, device=device))
        return self

    def _weights(self) -> F_.LifW:
        dw, pw = self.spread[0], self.spread[1]
        Cp = F_.pad64(dw.weight.shape[0])   # narrow layers (res*-ee.yaml front: 3 / 32 channels) run zero-padded
        return _cached(self, "lif", (dw.weight, dw.bias, pw.weight, pw.bias),
                       lambda: F_.make_lif_w(dw.weight, dw.bias, pw.weight, pw.bias, Cp))

    def analog(self, x: Act, affine=None) -> Act:
        """act=True: real-valued silu(mem) outputs (class Conv)."""
        if self.spread is None:
            self._init_spread(x.C, x.data.device)
        return F_.lif_silu(x, self._weights(), affine, self.ecs_tau, self.alpha, self.beta,
                           inplace=bool(getattr(self.actFun, "inplace", False)))

    def spikes(self, x: Act, affine=None) -> Spikes:
        if self.act:
            raise RuntimeError("mem_update(act=True) produces real values: use analog()")
        if self.spread is None:
            self._init_spread(x.C, x.data.device)
        if F_._state["dispatch"] == "ops" and x.C % 64 == 0 and not self.training and not torch.is_grad_enabled():
            # the inference path goes through the PyTorch custom operator (ops.py) over the C ABI
            from . import ops  # noqa: F401  (registers torch.ops.ecsy.*)
            w = self._weights()
            sc, sh = affine if affine is not None else (None, None)
            bits = torch.ops.ecsy.lif_ecs_w(x.data, x.T, w.dw_w, w.dw_b, w.pw, w.pw_b, w.w_eff, w.bconst, w.w_wave, w.splits,
                                            sc, sh, float(self.ecs_tau), float(self.alpha), float(self.beta))
            return Spikes(bits, x.C)
        return F_.lif_ecs(x, self._weights(), affine, self.ecs_tau, self.alpha, self.beta)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.act:
            return self.analog(Act.from_ref(x)).to_ref()
        sp = self.spikes(Act.from_ref(x))
        return _tag_spikes(sp.to_act().to_ref(), sp)


class Snn_Conv2d(nn.Conv2d):
    """Per-timestep convolution (models/common.py:593-624)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode='zeros', marker='b'):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.marker = marker

    def _geom(self):
        k, s, p, d = self.kernel_size, self.stride, self.padding, self.dilation
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or d != (1, 1) or self.padding_mode != 'zeros':
            raise NotImplementedError("Snn_Conv2d: only square kernels / strides / zero padding, dilation 1")
        return k[0], s[0], p[0]

    def umma_ok(self) -> bool:
        return self.out_channels % 64 == 0 and self.bias is None

    def _w(self) -> F_.ConvW:
        k, s, p = self._geom()
        umma = self.out_channels % 64 == 0
        return _cached(self, "conv", (self.weight, self.bias),
                       lambda: F_.make_conv_w(self.weight, self.bias, s, p, self.groups, umma, True, densify=True))

    def _w_padded(self, cip: int, cop: int) -> F_.ConvW:
        """Weight zero-padded to [cop, cip, k, k]: narrow layers on the 64-channel tensor-core granule."""
        k, s, p = self._geom()

        def build():
            w = torch.nn.functional.pad(self.weight.detach().float(),
                                        (0, 0, 0, 0, 0, cip - self.in_channels, 0, cop - self.out_channels))
            return F_.make_conv_w(w, None, s, p, 1, True, False)
        return _cached(self, "conv_pad", (self.weight,), build)

    def conv_spikes(self, sp: Spikes, scale=None, shift=None, residual: Optional[Act] = None) -> Act:
        if sp.Cr != self.in_channels:
            raise RuntimeError(f"Snn_Conv2d: {sp.Cr} spike channels for a {self.in_channels}-channel conv")
        if self.umma_ok() and sp.C == self.in_channels:
            if F_._state["dispatch"] == "ops" and self.groups == 1 and self.bias is None and not self.training and not torch.is_grad_enabled():
                from . import ops  # noqa: F401
                w = self._w()
                out = torch.ops.ecsy.spike_conv_w(sp.bits, sp.C, w.packed, w.packed_ts, w.splits, w.co, w.k, w.stride, w.pad,
                                                  scale, shift, residual.data if residual is not None else None)
                return Act(out, sp.T)
            return F_.spike_conv(sp, self._w(), scale, shift, residual)
        if self.bias is None and self.groups == 1:
            # narrow input and / or output (res*-ee.yaml: 3 -> 32 -> 32 -> 64): same tcgen05 kernel on padded weights
            co, cop = self.out_channels, F_.pad64(self.out_channels)
            w = self._w_padded(sp.C, cop)
            if cop == co:
                return F_.spike_conv(sp, w, scale, shift, residual)
            if scale is not None:
                scale, shift = F_.pad_channels(scale, cop), F_.pad_channels(shift, cop)
            y = F_.spike_conv(sp, w, scale, shift, None)
            y = Act(y.data[..., :co].contiguous(), y.T)
            return F_.affine_add(y, None, None, residual) if residual is not None else y
        y = F_.real_conv(sp.to_act(), self._w(), scale, shift)
        return F_.affine_add(y, None, None, residual) if residual is not None else y

    def conv_real(self, x: Act, scale=None, shift=None, bias_mul: float = 1.0) -> Act:
        return F_.real_conv(x, self._w(), scale, shift, bias_mul)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        sp = getattr(input, "_ecsy_spikes", None)
        if sp is not None:
            return self.conv_spikes(sp).to_ref()
        return self.conv_real(Act.from_ref(input)).to_ref()


class BatchNorm3d1(nn.BatchNorm3d):
    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            nn.init.constant_(self.weight, F_.thresh)
            nn.init.zeros_(self.bias)


class BatchNorm3d2(nn.BatchNorm3d):
    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            nn.init.constant_(self.weight, 0.2 * F_.thresh)
            nn.init.zeros_(self.bias)


class _tdbn(nn.Module):
    """Threshold-dependent BN over (N, T, H, W) (models/common.py:668-691)."""

    stat_updates = 1  # DDetect evaluates its branches twice per forward (models/yolo_snn.py:115-116)

    def scale_shift(self, y: Act):
        """Per-channel (scale, shift) with y_norm = y*scale + shift; updates running stats in training."""
        bn = self.bn
        if bn.training or not bn.track_running_stats:
            mean, var = F_.bn_stats(y)
            world = 1
            if isinstance(bn, nn.SyncBatchNorm) and bn.training:   # --sync-bn (train.py:359-360)
                from . import dist as D
                mean, var, world = D.sync_bn_stats(mean, var)
            with torch.no_grad():
                if bn.track_running_stats:
                    n = float(y.T * y.N * y.H * y.W) * world
                    for _ in range(self.stat_updates):
                        bn.num_batches_tracked += 1
                        m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                        bn.running_mean.mul_(1.0 - m).add_(mean, alpha=m)
                        bn.running_var.mul_(1.0 - m).add_(var, alpha=m * n / max(n - 1.0, 1.0))
                scale = bn.weight * torch.rsqrt(var + bn.eps)
                shift = bn.bias - mean * scale
            return scale.contiguous(), shift.contiguous()
        return _cached(self, "affine", (bn.weight, bn.bias, bn.running_mean, bn.running_var), self._eval_affine)

    def _eval_affine(self):
        bn = self.bn
        scale = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
        return scale, (bn.bias - bn.running_mean * scale).float().contiguous()

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        y = Act.from_ref(input)
        sc, sh = self.scale_shift(y)
        return F_.affine_add(y, sc, sh).to_ref()


class batch_norm_2d(_tdbn):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = BatchNorm3d1(num_features)


class batch_norm_2d1(_tdbn):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = BatchNorm3d2(num_features)


# ------------------------------------------------------------------------------------------------
# fused chains
# ------------------------------------------------------------------------------------------------
def _lif_conv_bn(lif: mem_update, conv: Snn_Conv2d, bn: _tdbn, x: Act, in_affine=None, residual: Optional[Act] = None):
    """LIF -> conv -> tdBN.  Eval: BN folded into the conv epilogue (+ residual) -> (Act, None).
    Train: raw conv output plus the pending (scale, shift) of its batch statistics."""
    sp = lif.spikes(x, in_affine)
    if bn.bn.training:
        y = conv.conv_spikes(sp)
        return y, bn.scale_shift(y)
    sc, sh = bn.scale_shift(None)
    return conv.conv_spikes(sp, sc, sh, residual), None


def _residual_path(seq: nn.Sequential, x: Act, shortcut: Act, sc_aff=None) -> Act:
    """[LIF, conv, BN, LIF, conv, BN] + shortcut (models/common.py:1191-1202, 1216)."""
    lif1, conv1, bn1, lif2, conv2, bn2 = seq
    y1, a1 = _lif_conv_bn(lif1, conv1, bn1, x)
    if a1 is None:
        out, _ = _lif_conv_bn(lif2, conv2, bn2, y1, None, shortcut)
        return out
    y2, a2 = _lif_conv_bn(lif2, conv2, bn2, y1, a1)
    sa, sb = sc_aff if sc_aff is not None else (None, None)
    return F_.affine_add(y2, a2[0], a2[1], shortcut, sa, sb)


def _make_residual(cin, hid, cout, k, stride, pad):
    seq = nn.Sequential(
        mem_update(act=False),
        Snn_Conv2d(cin, hid, kernel_size=k, stride=stride, padding=pad, bias=False),
        batch_norm_2d(hid),
        mem_update(act=False),
        Snn_Conv2d(hid, cout, kernel_size=k, padding=pad, bias=False),
        batch_norm_2d1(cout),
    )
    seq[0]._init_spread(cin)
    seq[3]._init_spread(hid)
    return seq


class _BasicBlock(nn.Module):
    def _build(self, cin, hid, cout, k, stride):
        pad = 1 if k == 3 else 0
        self.residual_function = _make_residual(cin, hid, cout, k, stride, pad)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(
                nn.MaxPool3d((1, stride, stride), stride=(1, stride, stride)),
                mem_update(act=False),
                Snn_Conv2d(cin, cout, kernel_size=1, stride=1, bias=False),
                batch_norm_2d(cout),
            )
            self.shortcut[1]._init_spread(cin)
        self._stride = stride

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            return AG.BasicBlockFn.apply(self, x, *self.parameters())
        a = Act.from_ref(x)
        if len(self.shortcut) == 0:
            return _residual_path(self.residual_function, a, a).to_ref()
        pool, lif, conv, bn = self.shortcut
        z, aff = _lif_conv_bn(lif, conv, bn, F_.maxpool(a, pool.stride[1]))
        return _residual_path(self.residual_function, a, z, aff).to_ref()


class BasicBlock_2(_BasicBlock):
    """models/common.py:1182-1219"""

    def __init__(self, in_channels, out_channels, k_size=3, stride=1):
        super().__init__()
        self._build(in_channels, out_channels, out_channels, k_size, stride)


class BasicBlock_1(_BasicBlock):
    """models/common.py:1049-1079 (hidden width fixed at 1024)"""

    def __init__(self, in_channels, out_channels, stride=1, ECS=False):
        super().__init__()
        self._build(in_channels, 1024, out_channels, 3, stride)


class Concat_res2(nn.Module):
    """models/common.py:1454-1488: shortcut = max-pool(cat(LIF->1x1->BN, x))."""

    def __init__(self, in_channels, out_channels, k_size=3, stride=1, ECS=False):
        super().__init__()
        self._build(in_channels, out_channels, out_channels, k_size, stride)

    def _build(self, in_channels, hid, out_channels, k_size, stride):
        pad = 1 if k_size == 3 else 0
        self.residual_function = _make_residual(in_channels, hid, out_channels, k_size, stride, pad)
        self.shortcut = nn.Sequential()
        if in_channels < out_channels:
            self.shortcut = nn.Sequential(
                mem_update(act=False),
                Snn_Conv2d(in_channels, out_channels - in_channels, kernel_size=1, stride=1, bias=False),
                batch_norm_2d(out_channels - in_channels),
            )
            self.shortcut[0]._init_spread(in_channels)
        self.pools = nn.MaxPool3d((1, stride, stride), stride=(1, stride, stride))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            return AG.ConcatRes2Fn.apply(self, x, *self.parameters())
        a = Act.from_ref(x)
        if len(self.shortcut) == 0:
            temp = a
        else:
            lif, conv, bn = self.shortcut
            temp, aff = _lif_conv_bn(lif, conv, bn, a)
            if aff is not None:
                temp = F_.affine_add(temp, aff[0], aff[1])
        sc = F_.concat_channels([temp, a], pool=self.pools.stride[1])
        return _residual_path(self.residual_function, a, sc).to_ref()


class BasicBlock_ms(nn.Module):
    """models/common.py:1658-1687 (res*-ee.yaml, the original EMS-YOLO block): hidden width e*out; the shortcut is
    max-pool -> 1x1 conv on the REAL input -> tdBN, with no neuron in front of the conv."""

    def __init__(self, in_channels, out_channels, k_size=3, stride=1, e=0.5):
        super().__init__()
        pad = 1 if k_size == 3 else 0
        self.residual_function = _make_residual(in_channels, int(out_channels * e), out_channels, k_size, stride, pad)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.shortcut = nn.Sequential(
                nn.MaxPool3d((1, stride, stride), stride=(1, stride, stride)),
                Snn_Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=False),
                batch_norm_2d(out_channels),
            )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            return AG.BasicBlockMsFn.apply(self, x, *self.parameters())
        a = Act.from_ref(x)
        if len(self.shortcut) == 0:
            return _residual_path(self.residual_function, a, a).to_ref()
        pool, conv, bn = self.shortcut
        z = F_.maxpool(a, pool.stride[1])
        if bn.bn.training:
            z = conv.conv_real(z)
            return _residual_path(self.residual_function, a, z, bn.scale_shift(z)).to_ref()
        sc, sh = bn.scale_shift(None)
        return _residual_path(self.residual_function, a, conv.conv_real(z, sc, sh)).to_ref()


class ConcatBlock_ms(Concat_res2):
    """models/common.py:1690-1725: Concat_res2's topology with hidden width e*out (res*-ee.yaml)."""

    def __init__(self, in_channels, out_channels, k_size=3, stride=1, e=0.5):
        nn.Module.__init__(self)
        self._build(in_channels, int(out_channels * e), out_channels, k_size, stride)



class Conv_1(nn.Module):
    """Stem: conv on the real image + tdBN, no neuron (models/common.py:409-425)."""

    def __init__(self, c1, c2, k, s, p=None, g=1, act=True):
        super().__init__()
        self.conv = Snn_Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = batch_norm_2d(c2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            out = AG.StemFn.apply(self, x, *self.parameters())
            return AG.ExpandT.apply(out, x.shape[0]) if out.shape[0] != x.shape[0] else out
        a = Act.from_ref(x)
        if self.bn.bn.training:
            y = self.conv.conv_real(a)
            sc, sh = self.bn.scale_shift(y)
            return F_.affine_add(y, sc, sh).to_ref()
        sc, sh = self.bn.scale_shift(None)
        return self.conv.conv_real(a, sc, sh).to_ref()

    def forward_fuse(self, x):
        return self.conv.conv_real(Act.from_ref(x)).to_ref()


class Conv_B(nn.Module):
    """LIF -> conv -> tdBN (models/common.py:393-406)."""

    def __init__(self, c1, c2, k, s=1, p=None, g=1, act=True):
        super().__init__()
        self.act = mem_update(act=False)
        self.conv = Snn_Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = batch_norm_2d(c2)
        self.act._init_spread(c1)

    def run(self, a: Act, defer_affine: bool = False):
        """-> Act, or (raw Act, (scale, shift)) in training when the caller applies the tdBN affine itself."""
        y, aff = _lif_conv_bn(self.act, self.conv, self.bn, a)
        if defer_affine:
            return y, aff
        return y if aff is None else F_.affine_add(y, aff[0], aff[1])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            return AG.ConvBFn.apply(self, x, *self.parameters())
        return self.run(Act.from_ref(x)).to_ref()


class Conv_2(Conv_B):
    """models/common.py:428-440 (same chain, explicit stride argument)."""

    def __init__(self, c1, c2, k, s, p=None, g=1):
        nn.Module.__init__(self)   # registration order of the reference (conv, bn, act) = state_dict key order
        self.conv = Snn_Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = batch_norm_2d(c2)
        self.act = mem_update(act=False)
        self.act._init_spread(c1)


class Conv(nn.Module):
    """conv on REAL input -> tdBN -> mem_update(act=True) (models/common.py:362-375)."""

    def __init__(self, c1, c2, k, s, p=None, g=1, ECS=False):
        super().__init__()
        self.conv = Snn_Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = batch_norm_2d(c2)
        self.act = mem_update(act=True, ECS=ECS)
        self.act._init_spread(c2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import autograd as AG
        if AG.wants_grad(self, x):
            return AG.ConvSiluFn.apply(self, x, *self.parameters())
        a = Act.from_ref(x)
        if self.bn.bn.training:
            y = self.conv.conv_real(a)
            return self.act.analog(y, self.bn.scale_shift(y)).to_ref()
        sc, sh = self.bn.scale_shift(None)
        return self.act.analog(self.conv.conv_real(a, sc, sh)).to_ref()


class Conv_7(nn.Module):
    """Learned T -> 1 fusion, a bias-free Conv3d(T, 1, 1) (models/common.py:549-562)."""

    def __init__(self, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv3d(time_window, 1, k, s, autopad(k, p), bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        a = Act.from_ref(x)
        if self.conv.weight.numel() != a.T:   # the reference's Conv3d raises a shape error here (models/common.py:549-562)
            raise RuntimeError(f"Conv_7: built for time_window={self.conv.weight.numel()} but the input has T={a.T} steps")
        out = F_.tsum(a, self.conv.weight.detach().reshape(-1).contiguous(), 1.0)
        return out.permute(0, 3, 1, 2)


class Sample(nn.Module):
    """Per-timestep nearest up-sampling (models/common.py:844-868)."""

    def __init__(self, size=None, scale_factor=None, mode='nearset'):
        super().__init__()
        self.scale_factor = scale_factor
        self.mode = mode
        self.size = size
        self.up = nn.Upsample(self.size, self.scale_factor, mode=self.mode)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self.mode != 'nearest' or self.size is not None:
            raise NotImplementedError("Sample: only nearest up-sampling by an integer factor")
        if torch.is_grad_enabled() and input.requires_grad:
            from . import autograd as AG
            return AG.SampleFn.apply(input, int(self.scale_factor))
        return F_.upsample(Act.from_ref(input), int(self.scale_factor)).to_ref()


class Concat(nn.Module):
    """Channel concat of 5-D tensors (models/common.py:1758-1765; the YAMLs pass dimension 2)."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        if self.d == 2 and all(t.dim() == 5 and t.is_cuda for t in x):
            if torch.is_grad_enabled() and any(t.requires_grad for t in x):
                from . import autograd as AG
                return AG.ConcatFn.apply(*x)
            return F_.concat_channels([Act.from_ref(t) for t in x]).to_ref()
        return torch.cat(x, self.d)


class DFL(nn.Module):
    """Distribution-focal-loss expectation (models/common.py:312-323); DDetect decodes in-kernel."""

    def __init__(self, c1=17):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = nn.Parameter(torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1))
        self.c1 = c1

    def forward(self, x):
        b, c, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)
