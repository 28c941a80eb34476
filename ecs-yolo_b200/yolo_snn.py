"""Stack B model assembly: anchor-free DDetect head (mean over T, DFL decode) and DetectionModel.

Mirrors the hot-path surface of the reference's ``models/yolo_snn.py``: ``DDetect`` (:83-139),
``DetectionModel`` (:687-749) and ``parse_model`` (:829-914, shared with Stack A here), with the same
constructor signatures, attributes (``stride``, ``names``, ``yaml``, ``save``, ``nc``,
``DDetect.{nc,nl,no,reg_max,stride}``) and state_dict keys.
"""
from __future__ import annotations

import math
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn

from . import common
from . import functional as F_
from .common import Act, Conv_B, DFL, Snn_Conv2d, _cached
from .yolo import make_divisible, parse_model as _parse_model


class DDetect(nn.Module):
    dynamic = False
    export = False
    shape = None

    def __init__(self, nc=80, ch=(), inplace=True):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.inplace = inplace
        self.stride = torch.zeros(self.nl)
        c2 = make_divisible(max((ch[0] // 4, self.reg_max * 4, 16)), 4)
        c3 = max((ch[0], min((self.nc * 2, 128))))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv_B(x, c2, 3), Conv_B(c2, c2, 3, g=4), Snn_Conv2d(c2, 4 * self.reg_max, 1, groups=4))
            for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(Conv_B(x, c3, 3), Conv_B(c3, c3, 3), Snn_Conv2d(c3, self.nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max) if self.reg_max > 1 else nn.Identity()
        # The reference evaluates every branch twice per forward (yolo_snn.py:115-116: once for the sum, once
        # for `.size()[0]`), so in training each tdBN's running statistics receive the momentum update twice.
        # The second evaluation is not recomputed here; its only side effect is reproduced.
        for m in self.modules():
            if isinstance(m, common._tdbn):
                m.stat_updates = 2

    def _branch(self, seq: nn.Sequential, a: Act) -> torch.Tensor:
        """Conv_B -> Conv_B -> 1x1 conv (+bias, real input), then mean over T.  The mean commutes with the
        last (linear) conv, so the features are averaged first and convolved once."""
        y1 = seq[0].run(a)
        y2 = seq[1].run(y1)
        feat = F_.tsum(y2, None, float(y2.T))          # [N,H,W,C]
        conv = seq[2]
        if conv.in_channels % 64 == 0 and conv.kernel_size[0] == 1 and not self.training:
            # inference: the last 1x1 (+ bias, grouped in the box branch) on the dense tcgen05 GEMM instead of the SIMT conv
            cw, sc1, sh1 = _cached(conv, "headw_umma", (conv.weight, conv.bias),
                                   lambda: F_.make_head_conv_w(conv.weight, conv.bias, conv.groups))
            return F_.real_conv(Act(feat.unsqueeze(0), 1), cw, sc1, sh1).data[0][..., :conv.out_channels].contiguous()
        out = conv.conv_real(Act(feat.unsqueeze(0), 1))
        return out.data[0]                              # [N,H,W,Cout]

    def _forward_train_autograd(self, x):
        """Training with gradients: the two Conv_B trunks of each branch run through the manual backward chains
        (autograd.ConvBChainFn); the mean over T, the last 1x1 conv (+bias) and the concat -- N*H*W-row tensors,
        < 0.1 % of the FLOPs -- are differentiable torch ops (models/yolo_snn.py:115-116)."""
        import torch.nn.functional as tF
        from . import autograd as AG
        out = []
        for i in range(self.nl):
            parts = []
            for seq in (self.cv2[i], self.cv3[i]):
                params = list(seq[0].parameters()) + list(seq[1].parameters())
                y2n = AG.ConvBChainFn.apply(seq[0], seq[1], x[i], *params)
                feat = y2n.sum(dim=0) / y2n.shape[0]
                parts.append(tF.conv2d(feat, seq[2].weight, seq[2].bias, groups=seq[2].groups))
            out.append(torch.cat(parts, 1))
        return out

    def forward(self, x):
        x = list(x)
        if self.training and torch.is_grad_enabled() and (any(t.requires_grad for t in x) or
                                                          any(p.requires_grad for p in self.parameters())):
            return self._forward_train_autograd(x)
        a_total = sum(xi.shape[3] * xi.shape[4] for xi in x)
        N = x[0].shape[1]
        y = None
        if not self.training:
            y = torch.empty(N, 4 + self.nc, a_total, device=x[0].device, dtype=torch.float32)
        strides = getattr(self, "_strides", None) or [float(s) for s in self.stride]
        off = 0
        for i in range(self.nl):
            a = Act.from_ref(x[i])
            box = self._branch(self.cv2[i], a)
            cls = self._branch(self.cv3[i], a)
            x[i] = F_.ddetect_decode(box, cls, strides[i], y, off)
            off += box.shape[1] * box.shape[2]
        if self.training:
            return x
        return y if self.export else (y, x)

    def bias_init(self):
        m = self
        for a, b, s in zip(m.cv2, m.cv3, m.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:m.nc] = math.log(5 / m.nc / (640 / s) ** 2)


class DetectionModel(nn.Module):
    def __init__(self, cfg='resnet18.yaml', ch=3, nc=None, anchors=None):
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
        self.model, self.save, reductions = _parse_model(deepcopy(self.yaml), ch=[ch], heads={"DDetect": DDetect})
        self.names = [str(i) for i in range(self.yaml['nc'])]
        self.inplace = self.yaml.get('inplace', True)
        m = self.model[-1]
        if isinstance(m, DDetect):
            m.inplace = self.inplace
            m.stride = torch.tensor([float(reductions[j]) for j in m.f])
            m._strides = [float(reductions[j]) for j in m.f]
            self.stride = m.stride
            m.bias_init()
        # utils/torch_utils.py:157-166 initialize_weights: activations become in-place, which changes what
        # mem_update(act=True) stores as mem_old (common.py:280)
        for mod in self.modules():
            if isinstance(mod, nn.SiLU):
                mod.inplace = True

    def forward(self, x, augment=False, profile=False, visualize=False):
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

    def _apply(self, fn):
        self = super()._apply(fn)
        m = self.model[-1]
        if isinstance(m, DDetect):
            m.stride = fn(m.stride)
        return self
