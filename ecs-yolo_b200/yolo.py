"""Stack A model assembly: YAML -> modules, T-replication of the image, anchor-based Detect head.

Mirrors the reference's ``models/yolo.py`` public surface for the hot path -- ``Detect`` (:50-161),
``Model`` (:165-312) and ``parse_model`` (:434-553) -- with the same constructor signatures,
attributes read by the loss / validation code (``stride``, ``names``, ``yaml``, ``save``, ``nc``,
``Detect.{nl,na,no,nc,anchors,stride}``, per-layer ``i f type np``) and state_dict keys.
"""
from __future__ import annotations

import math
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn

from . import common
from . import functional as F_
from .common import *  # noqa: F401,F403  -- the YAML "registry" is this namespace, as in the reference
from .common import Act, Conv_7, Snn_Conv2d, _cached


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


class Detect(nn.Module):
    stride = None
    onnx_dynamic = False

    def __init__(self, nc=80, anchors=(), ch=(), inplace=True, use_cupy=False):
        super().__init__()
        self.nc = nc
        self.no = nc + 5
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        self.grid = [torch.zeros(1)] * self.nl
        self.anchor_grid = [torch.zeros(1)] * self.nl
        self.register_buffer('anchors', torch.tensor(anchors).float().view(self.nl, -1, 2))
        self.m = nn.ModuleList(Snn_Conv2d(x, self.no * self.na, 1) for x in ch)
        self.inplace = inplace
        self.w = nn.ModuleList(Conv_7(1, 1) for x in ch)

    def forward(self, x):
        """x: list of REAL [T,N,C,H,W].  Conv_7's weighted sum over T commutes with the per-step 1x1
        conv + bias, so the features are reduced over T first (4x less head work), then one 1x1 conv
        with bias * sum(w) and the view/permute/decode kernel (models/yolo.py:85-146)."""
        x = list(x)
        if self.training and torch.is_grad_enabled() and any(t.requires_grad for t in x):
            return self._forward_train_autograd(x)
        z_rows = sum(self.na * xi.shape[3] * xi.shape[4] for xi in x)
        z = None
        if not self.training:
            z = torch.empty(x[0].shape[1], z_rows, self.no, device=x[0].device, dtype=torch.float32)
        off = 0
        for i in range(self.nl):
            a = Act.from_ref(x[i])
            wt = self.w[i].conv.weight
            if wt.numel() != a.T:   # the reference's Conv3d(T -> 1) raises a shape error (models/common.py:549-562)
                raise RuntimeError(f"Detect: built for time_window={wt.numel()} but the input has T={a.T} steps")
            tw = _cached(self.w[i], "tw", (wt,), lambda: wt.detach().reshape(-1).float().contiguous())
            feat = F_.tsum(a, tw, 1.0)                                   # [N,H,W,C]
            conv = self.m[i]
            if conv.in_channels % 64 == 0:
                # tensor-core path: output channels zero-padded to the 64-column tile, bias * sum(w) as the epilogue's shift,
                # ALWAYS the three-term bf16 split (hi*hi + lo*hi + hi*lo, ~1e-6): the head stays at fp32 accuracy in fast
                # precision too.  (The one-thread-per-output SIMT conv took 0.84 ms per batch-64 forward for 3.5 GFLOP.)
                cw, sc1, sh1 = _cached(conv, "detw_umma", (conv.weight, conv.bias, wt), lambda: _detect_conv_w(conv, wt))
                y = F_.real_conv(Act(feat.unsqueeze(0), 1), cw, sc1, sh1).data[0][..., :conv.out_channels].contiguous()
            else:
                cw = _cached(conv, "detw", (conv.weight, conv.bias, wt), lambda: F_.make_conv_w(
                    conv.weight, conv.bias * wt.detach().sum(), 1, 0, 1, False, True))
                y = F_.real_conv(Act(feat.unsqueeze(0), 1), cw).data[0]      # [N,H,W,na*no]
            st = getattr(self, "_strides", None)
            stride_i = (st[i] if st else float(self.stride[i])) if z is not None else 1.0
            x[i] = F_.detect_decode(y, self.na, self.no, self.anchors[i].contiguous(), stride_i, z, off)
            off += self.na * y.shape[1] * y.shape[2]
        return x if self.training else (z, x)


    def _forward_train_autograd(self, x):
        """Training with gradients: the head (< 0.1 % of the model's FLOPs, tensors of N*(ny*nx) rows) runs as
        differentiable torch ops -- T-fusion, 1x1 conv + bias, view/permute (models/yolo.py:95-104)."""
        import torch.nn.functional as tF
        out = []
        for i in range(self.nl):
            xi = x[i]
            if self.w[i].conv.weight.numel() != xi.shape[0]:
                raise RuntimeError(f"Detect: built for time_window={self.w[i].conv.weight.numel()} but the input has "
                                   f"T={xi.shape[0]} steps")
            wt = self.w[i].conv.weight.reshape(-1, 1, 1, 1, 1)
            feat = (xi * wt).sum(0)
            y = tF.conv2d(feat, self.m[i].weight, None) + self.m[i].bias.view(1, -1, 1, 1) * wt.sum()
            bs, _, ny, nx = y.shape
            out.append(y.view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous())
        return out


def _detect_conv_w(conv, wt):
    """Detect's 1x1 conv (+ bias * sum of the T-fusion weights) for the dense tcgen05 GEMM (functional.make_head_conv_w)."""
    return F_.make_head_conv_w(conv.weight, conv.bias, 1, wt.detach().float().sum())


class Model(nn.Module):
    def __init__(self, cfg='resnet34.yaml', ch=3, nc=None, anchors=None, use_cupy=False):
        super().__init__()
        if isinstance(cfg, dict):
            self.yaml = cfg
        else:
            import yaml
            self.yaml_file = Path(cfg).name
            with open(cfg, encoding='ascii', errors='ignore') as f:
                self.yaml = yaml.safe_load(f)
        ch = self.yaml['ch'] = self.yaml.get('ch', ch)
        if nc and nc != self.yaml['nc']:
            self.yaml['nc'] = nc
        if anchors:
            self.yaml['anchors'] = round(anchors)
        self.model, self.save, reductions = parse_model(deepcopy(self.yaml), ch=[ch], use_cupy=use_cupy)
        self.names = [str(i) for i in range(self.yaml['nc'])]
        self.inplace = self.yaml.get('inplace', True)
        m = self.model[-1]
        if isinstance(m, Detect):
            m.inplace = self.inplace
            # the reference measures strides with a 256x256 CPU probe forward (yolo.py:228); every
            # in-scope layer changes resolution by its stride argument, so they follow from the plan
            m.stride = torch.tensor([float(reductions[j]) for j in m.f])
            m.anchors /= m.stride.view(-1, 1, 1)
            a = m.anchors.prod(-1).view(-1)
            if (a[-1] - a[0]).sign() != (m.stride[-1] - m.stride[0]).sign():
                m.anchors[:] = m.anchors.flip(0)
            self.stride = m.stride
            m._strides = [float(v) for v in m.stride]
            self._initialize_biases()
        for mod in self.modules():  # utils/torch_utils.py:157-166 initialize_weights
            if isinstance(mod, nn.SiLU):
                mod.inplace = True

    def forward(self, x, augment=False, profile=False, visualize=False):
        """x: [N,3,H,W] image (direct coding: the same frame every timestep, yolo.py:248-251 -- kept as
        ONE frame plus a broadcast flag instead of T copies) or [T,N,C,H,W] event frames."""
        if augment:
            raise NotImplementedError("test-time augmentation is outside the hot path")
        if x.dim() == 4:
            x = x.unsqueeze(0).expand(common.time_window, -1, -1, -1, -1)
        return self._forward_once(x)

    def _forward_once(self, x, profile=False, visualize=False):
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
        return x

    def _initialize_biases(self, cf=None):
        m = self.model[-1]
        for mi, s in zip(m.m, m.stride):
            b = mi.bias.view(m.na, -1)
            b.data[:, 4] += math.log(8 / (640 / s) ** 2)
            b.data[:, 5:] += math.log(0.6 / (m.nc - 0.999999)) if cf is None else torch.log(cf / cf.sum())
            mi.bias = torch.nn.Parameter(b.view(-1), requires_grad=True)

    def _apply(self, fn):
        self = super()._apply(fn)
        m = self.model[-1]
        if isinstance(m, Detect):
            m.stride = fn(m.stride)
        return self


_CHANNEL_MODULES = ("Conv", "Conv_1", "Conv_2", "Conv_B", "BasicBlock_1", "BasicBlock_2", "Concat_res2",
                    "BasicBlock_ms", "ConcatBlock_ms")


def parse_model(d, ch, use_cupy=False, heads=None):
    """YAML rows [from, number, module, args] -> nn.Sequential (models/yolo.py:434-553 for the in-scope
    module types).  Also returns each layer's total stride for the Detect strides."""
    anchors, nc, gd, gw = d['anchors'], d['nc'], d['depth_multiple'], d['width_multiple']
    na = (len(anchors[0]) // 2) if isinstance(anchors, list) else anchors
    no = na * (nc + 5)
    ns = {k: getattr(common, k) for k in dir(common)}
    ns.update(Detect=Detect, nn=nn)
    ns.update(heads or {})
    head_types = tuple((heads or {}).values())
    layers, save, c2 = [], [], ch[-1]
    red = []
    for i, (f, n, m, args) in enumerate(d['backbone'] + d['head']):
        name = m
        if isinstance(m, str):
            if m not in ns:
                raise NotImplementedError(f"module '{m}' is outside the ported hot path")
            m = ns[m]
        args = list(args)
        for j, a in enumerate(args):
            if isinstance(a, str):
                args[j] = {"nc": nc, "anchors": anchors, "None": None}.get(a, a)
        n = n_ = max(round(n * gd), 1) if n > 1 else n
        src = (f if isinstance(f, int) else f[0])
        base = 1 if i == 0 else red[src]
        r = base
        if name in _CHANNEL_MODULES:
            c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)
            args = [c1, c2, *args[1:]]
            if name == "BasicBlock_1":
                r = base * (args[2] if len(args) > 2 else 1)
            elif name in ("Conv_B",):
                r = base * (args[3] if len(args) > 3 else 1)
            else:
                r = base * (args[3] if len(args) > 3 else 1)
        elif m is common.Concat:
            c2 = sum(ch[x] for x in f)
        elif m is Detect:
            args.append([ch[x] for x in f])
            if isinstance(args[1], int):
                args[1] = [list(range(args[1] * 2))] * len(f)
        elif head_types and m in head_types:
            args.append([ch[x] for x in f])
        elif m is common.Sample:
            c2 = ch[f]
            r = base // int(args[1])
        else:
            c2 = ch[f]
        if n > 1:
            # repeats of a block: the first maps c1 -> c2, the rest c2 -> c2 would need new args; the
            # reference passes identical args to every repeat (yolo.py:539), which is only valid when c1 == c2
            m_ = nn.Sequential(*(m(*args) for _ in range(n)))
        else:
            m_ = m(*args)
        t = name if isinstance(name, str) else m.__name__
        np_ = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type, m_.np = i, f, f"models.common.{t}", np_
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
        red.append(r)
    return nn.Sequential(*layers), sorted(save), red
