"""CPU oracle for the ECS-YOLO spiking hot path -- TEST INFRASTRUCTURE ONLY.

A functional, state-dict-driven restatement (plain torch fp32 on CPU) of the reference's
time-unrolled spiking backbone + detection heads.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product
package (``ecs-yolo_b200/``) never does and fails loudly when its CUDA library is missing.

Pinning: every function here is checked (tests/test_oracle_golden.py, ``-m "not gpu"``) against
fixtures in ``tests/golden/`` that were produced by importing the UNMODIFIED reference from
/root/reference in the build container (``oracle/gen_golden.py``; the reference ships no golden
vectors of its own).  On CPU the oracle is bit-identical to the reference for
the pinned cases because it issues the same aten calls in the same order.

All tensors are reference-shaped: activations ``[T, N, C, H, W]`` fp32, weights under the
reference's ``state_dict`` key names.  ``sd`` arguments are plain ``dict[str, Tensor]``; BN
running statistics are updated in place when ``training=True``.

Reference citations are to /root/reference (mowanggui/ECS-YOLO).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# models/common.py:37-39
THRESH = 0.5
LENS = 0.5
DECAY = 0.25
# nn.BatchNorm3d default momentum; the calibration recipe sets it to 1.0
# for one train-mode pass before switching to eval.
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# a1  ActFun  (models/common.py:56-82)
# --------------------------------------------------------------------------------------
class _Spike(torch.autograd.Function):
    """Heaviside forward, rectangular surrogate backward (models/common.py:61-79)."""

    @staticmethod
    def forward(ctx, mem):
        ctx.save_for_backward(mem)
        return mem.gt(THRESH).float()

    @staticmethod
    def backward(ctx, g):
        (mem,) = ctx.saved_tensors
        win = (abs(mem - THRESH) < LENS) / (2 * LENS)
        return g.clone() * win.float()


spike_fn = _Spike.apply


# --------------------------------------------------------------------------------------
# a2  mem_update  (models/common.py:236-309)
# --------------------------------------------------------------------------------------
def ecs_lif(x: torch.Tensor, dw_w, dw_b, pw_w, pw_b, act: bool = False, ecs_tau: float = 5.0,
            alpha: float = 0.75, beta: float = 0.25, silu_inplace: bool = False,
            record: Optional[dict] = None) -> torch.Tensor:
    """ECS-LIF scan over T = x.shape[0] steps.

    x: [T,N,C,H,W].  spread = pw 1x1 (C->C, bias) o depthwise 3x3 (pad 1, bias), common.py:285-298.
    Evaluation order follows common.py:263-281 and the scripted charge at :306-309.
    ``silu_inplace`` reproduces the reference quirk that ``initialize_weights`` turns the
    ``nn.SiLU`` of ``mem_update(act=True)`` into an in-place op (utils/torch_utils.py:165-166),
    so ``mem_old`` becomes silu(mem) for the SiLU variant.
    """
    T = x.shape[0]
    C = x.shape[2]
    out = torch.zeros_like(x)
    mem_old = 0
    spike = torch.zeros_like(x[0])
    ecs = 0.0
    fecs = 0.0
    for t in range(T):
        if t >= 1:
            mem = mem_old * DECAY * (1 - spike.detach()) + x[t] + fecs
        else:
            mem = x[t] + fecs
        if record is not None:
            record.setdefault("mem", []).append(mem.detach().clone())
        if act:
            spike = F.silu(mem, inplace=silu_inplace)
        else:
            spike = spike_fn(mem)
        s = F.conv2d(spike, dw_w, dw_b, 1, 1, 1, C)
        s = F.conv2d(s, pw_w, pw_b)
        ecs = alpha * s + (1.0 - 1.0 / ecs_tau) * ecs
        fecs = beta * torch.tanh(ecs)
        mem_old = mem.clone()
        out[t] = spike
    return out


def lif_from_sd(sd: Dict[str, torch.Tensor], prefix: str, x, act=False, silu_inplace=False, record=None):
    return ecs_lif(x, sd[prefix + "spread.0.weight"], sd[prefix + "spread.0.bias"],
                   sd[prefix + "spread.1.weight"], sd[prefix + "spread.1.bias"], act=act,
                   silu_inplace=silu_inplace, record=record)


# --------------------------------------------------------------------------------------
# a3  Snn_Conv2d  (models/common.py:593-624)
# --------------------------------------------------------------------------------------
def snn_conv2d(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, stride: int = 1,
               padding: int = 0, groups: int = 1) -> torch.Tensor:
    """Per-timestep conv2d into a zero-initialised [T,N,Co,Ho,Wo] buffer (common.py:619-624)."""
    T, N = x.shape[0], x.shape[1]
    k = w.shape[-1]
    ho = (x.shape[3] - (k - 1) + 2 * padding - 1) // stride + 1
    wo = (x.shape[4] - (k - 1) + 2 * padding - 1) // stride + 1
    out = torch.zeros(T, N, w.shape[0], ho, wo, dtype=x.dtype, device=x.device)
    for t in range(T):
        out[t] = F.conv2d(x[t], w, b, stride, padding, 1, groups)
    return out


# --------------------------------------------------------------------------------------
# a4  tdBN  (models/common.py:668-700, 753-758)
# --------------------------------------------------------------------------------------
def tdbn(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, training: bool,
         eps: float = 1e-5, momentum: Optional[float] = None) -> torch.Tensor:
    """BatchNorm3d over (N,T,H,W) per channel after the reference's two transposing copies
    (common.py:674-679).  ``prefix`` ends in ``'bn.'``.  Updates running stats in place when training."""
    momentum = BN_MOMENTUM if momentum is None else momentum
    y = x.transpose(0, 2).contiguous().transpose(0, 1).contiguous()  # [N,C,T,H,W]
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    if training and (prefix + "num_batches_tracked") in sd:
        sd[prefix + "num_batches_tracked"] += 1
    y = F.batch_norm(y, rm, rv, sd[prefix + "weight"], sd[prefix + "bias"], training, momentum, eps)
    return y.contiguous().transpose(0, 1).contiguous().transpose(0, 2)


def maxpool_hw(x: torch.Tensor, s: int) -> torch.Tensor:
    """nn.MaxPool3d((1,s,s), stride=(1,s,s)) on [T,N,C,H,W] (common.py:1209, 1481): the pooled
    'depth' axis is C (kernel 1), so this is a spatial s x s max-pool."""
    if s == 1:
        return x
    return F.max_pool3d(x, (1, s, s), (1, s, s))


# --------------------------------------------------------------------------------------
# a5  blocks  (models/common.py:1049-1079, 1182-1219, 1454-1488, 362-425)
# --------------------------------------------------------------------------------------
def _residual(sd, p, x, stride, pad, training, rec):
    """[LIF -> conv(k, stride) -> tdBN(gamma0=.5) -> LIF -> conv(k) -> tdBN(gamma0=.1)]"""
    r = p + "residual_function."
    s = lif_from_sd(sd, r + "0.", x)
    if rec is not None:
        rec[r + "0"] = s
    y = snn_conv2d(s, sd[r + "1.weight"], None, stride, pad)
    y = tdbn(y, sd, r + "2.bn.", training)
    s = lif_from_sd(sd, r + "3.", y)
    if rec is not None:
        rec[r + "3"] = s
    y = snn_conv2d(s, sd[r + "4.weight"], None, 1, pad)
    return tdbn(y, sd, r + "5.bn.", training)


def basic_block(sd, p, x, cin, cout, k=3, stride=1, training=False, rec=None):
    """BasicBlock_2 (common.py:1182-1219) and BasicBlock_1 (:1049-1079; hidden width comes from
    the weight shapes, the only difference)."""
    pad = 1 if k == 3 else 0
    out = _residual(sd, p, x, stride, pad, training, rec)
    if stride != 1 or cin != cout:
        q = p + "shortcut."
        z = maxpool_hw(x, stride)
        s = lif_from_sd(sd, q + "1.", z)
        if rec is not None:
            rec[q + "1"] = s
        z = snn_conv2d(s, sd[q + "2.weight"], None, 1, 0)
        sc = tdbn(z, sd, q + "3.bn.", training)
    else:
        sc = x
    return out + sc


def concat_res2(sd, p, x, cin, cout, k=3, stride=1, training=False, rec=None):
    """Concat_res2 (common.py:1454-1488): shortcut = cat(LIF->1x1(cout-cin)->BN, x) -> max-pool."""
    pad = 1 if k == 3 else 0
    if cin < cout:
        q = p + "shortcut."
        s = lif_from_sd(sd, q + "0.", x)
        if rec is not None:
            rec[q + "0"] = s
        z = snn_conv2d(s, sd[q + "1.weight"], None, 1, 0)
        temp = tdbn(z, sd, q + "2.bn.", training)
    else:
        temp = x
    out = torch.cat((temp, x), dim=2) if cin < cout else torch.cat((x, x), dim=2)
    out = maxpool_hw(out, stride)
    return _residual(sd, p, x, stride, pad, training, rec) + out


def basic_block_ms(sd, p, x, cin, cout, k=3, stride=1, training=False, rec=None):
    """BasicBlock_ms (common.py:1658-1687, res*-ee.yaml): hidden width 0.5*cout (implicit in the weight
    shapes); the shortcut has NO neuron: max-pool -> 1x1 conv on the real input -> tdBN."""
    pad = 1 if k == 3 else 0
    out = _residual(sd, p, x, stride, pad, training, rec)
    if stride != 1 or cin != cout:
        q = p + "shortcut."
        z = snn_conv2d(maxpool_hw(x, stride), sd[q + "1.weight"], None, 1, 0)
        sc = tdbn(z, sd, q + "2.bn.", training)
    else:
        sc = x
    return out + sc


def concat_block_ms(sd, p, x, cin, cout, k=3, stride=1, training=False, rec=None):
    """ConcatBlock_ms (common.py:1690-1725): Concat_res2's data flow; only the hidden width (0.5*cout,
    implicit in the weight shapes) differs."""
    return concat_res2(sd, p, x, cin, cout, k, stride, training, rec)


def conv_1(sd, p, x, k, s, training=False):
    """Conv_1 (common.py:409-425): conv on the REAL input then tdBN, no neuron."""
    y = snn_conv2d(x, sd[p + "conv.weight"], None, s, k // 2)
    return tdbn(y, sd, p + "bn.bn.", training)


def conv_b(sd, p, x, k, s=1, g=1, training=False, rec=None):
    """Conv_B / Conv_2 (common.py:393-406, 428-440): LIF -> conv -> tdBN."""
    sp = lif_from_sd(sd, p + "act.", x)
    if rec is not None:
        rec[p + "act"] = sp
    y = snn_conv2d(sp, sd[p + "conv.weight"], None, s, k // 2, g)
    return tdbn(y, sd, p + "bn.bn.", training)


def conv_silu(sd, p, x, k, s, g=1, training=False, silu_inplace=True):
    """Conv (common.py:362-375): conv on REAL input -> tdBN -> mem_update(act=True) (SiLU 'analog spikes')."""
    y = snn_conv2d(x, sd[p + "conv.weight"], None, s, k // 2, g)
    y = tdbn(y, sd, p + "bn.bn.", training)
    return lif_from_sd(sd, p + "act.", y, act=True, silu_inplace=silu_inplace)


# --------------------------------------------------------------------------------------
# a6  Sample / Concat  (models/common.py:844-868, 1758-1765)
# --------------------------------------------------------------------------------------
def sample_nearest(x: torch.Tensor, scale: int) -> torch.Tensor:
    out = torch.zeros(x.shape[0], x.shape[1], x.shape[2], x.shape[3] * scale, x.shape[4] * scale, dtype=x.dtype,
                      device=x.device)
    for t in range(x.shape[0]):
        out[t] = F.interpolate(x[t], scale_factor=float(scale), mode="nearest")
    return out


# --------------------------------------------------------------------------------------
# a8  Detect (Stack A)  (models/yolo.py:50-161) + Conv_7 (models/common.py:549-562)
# --------------------------------------------------------------------------------------
def detect_a(sd, p, feats: Sequence[torch.Tensor], nc: int, anchors: torch.Tensor, stride: torch.Tensor,
             training: bool):
    """feats: list of REAL [T,N,C,H,W]; anchors: [nl,na,2] in grid units (already divided by stride,
    yolo.py:230); returns list of [N,na,ny,nx,no] (train) or (z [N,sum,no], list) (eval)."""
    nl, na = anchors.shape[0], anchors.shape[1]
    no = nc + 5
    xs, z = [], []
    for i in range(nl):
        y = snn_conv2d(feats[i], sd[f"{p}m.{i}.weight"], sd[f"{p}m.{i}.bias"])  # 1x1 + bias per step
        y = y.permute(1, 0, 2, 3, 4)  # [N,T,C,H,W]  (Conv_7, common.py:558-562)
        y = F.conv3d(y, sd[f"{p}w.{i}.conv.weight"]).squeeze(dim=1)
        bs, _, ny, nx = y.shape
        y = y.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        xs.append(y)
        if not training:
            yv, xv = torch.meshgrid([torch.arange(ny, device=y.device), torch.arange(nx, device=y.device)], indexing="ij")
            grid = torch.stack((xv, yv), 2).expand((1, na, ny, nx, 2)).float()
            ag = (anchors[i].clone() * stride[i]).view((1, na, 1, 1, 2)).expand((1, na, ny, nx, 2)).float()
            o = y.sigmoid()
            o[..., 0:2] = (o[..., 0:2] * 2 - 0.5 + grid) * stride[i]
            o[..., 2:4] = (o[..., 2:4] * 2) ** 2 * ag
            z.append(o.view(bs, -1, no))
    return xs if training else (torch.cat(z, 1), xs)


# --------------------------------------------------------------------------------------
# a9  DDetect (Stack B)  (models/yolo_snn.py:83-139) + DFL (common.py:312-323) + anchors
#     (utils/tal/anchor_generator.py:8-32)
# --------------------------------------------------------------------------------------
def _ddetect_branch(sd, p, x, groups_mid: int, groups_last: int, training: bool, rec=None):
    y = conv_b(sd, p + "0.", x, 3, 1, 1, training, rec)
    y = conv_b(sd, p + "1.", y, 3, 1, groups_mid, training, rec)
    return snn_conv2d(y, sd[p + "2.weight"], sd[p + "2.bias"], 1, 0, groups_last)


def ddetect(sd, p, feats: Sequence[torch.Tensor], nc: int, stride: torch.Tensor, training: bool,
            double_eval: bool = True, rec=None):
    """Mean over T of the box (cv2) and class (cv3) branches.  The reference evaluates every
    branch twice per forward (yolo_snn.py:115-116), which in training mode applies the tdBN
    momentum update twice; ``double_eval`` reproduces that."""
    reg_max = 16
    no = nc + 4 * reg_max
    xs = []
    for i, f in enumerate(feats):
        a = _ddetect_branch(sd, f"{p}cv2.{i}.", f, 4, 4, training, rec)
        if double_eval:
            a2 = _ddetect_branch(sd, f"{p}cv2.{i}.", f, 4, 4, training)
            a = a.sum(dim=0) / a2.size()[0]
        else:
            a = a.sum(dim=0) / a.size()[0]
        b = _ddetect_branch(sd, f"{p}cv3.{i}.", f, 1, 1, training, rec)
        if double_eval:
            b2 = _ddetect_branch(sd, f"{p}cv3.{i}.", f, 1, 1, training)
            b = b.sum(dim=0) / b2.size()[0]
        else:
            b = b.sum(dim=0) / b.size()[0]
        xs.append(torch.cat((a, b), 1))
    if training:
        return xs
    pts, strs = [], []
    for i, s in enumerate(stride):
        _, _, h, w = xs[i].shape
        sx = torch.arange(end=w, dtype=torch.float32, device=xs[i].device) + 0.5
        sy = torch.arange(end=h, dtype=torch.float32, device=xs[i].device) + 0.5
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        strs.append(torch.full((h * w, 1), float(s), dtype=torch.float32, device=xs[i].device))
    anchors = torch.cat(pts).transpose(0, 1)
    strides = torch.cat(strs).transpose(0, 1)
    bs = xs[0].shape[0]
    box, cls = torch.cat([xi.view(bs, no, -1) for xi in xs], 2).split((reg_max * 4, nc), 1)
    b, _, a = box.shape
    proj = torch.arange(reg_max, dtype=torch.float32, device=box.device).view(1, reg_max, 1, 1)
    dist = F.conv2d(box.view(b, 4, reg_max, a).transpose(2, 1).softmax(1), proj).view(b, 4, a)
    lt, rb = torch.split(dist, 2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
    return torch.cat((dbox, cls.sigmoid()), 1), xs


# --------------------------------------------------------------------------------------
# a7 / a10  model assembly  (models/yolo.py:247-312, 434-553; models/yolo_snn.py:632-646, 741-749, 829-914)
# --------------------------------------------------------------------------------------
def _make_divisible(x, d):
    return math.ceil(x / d) * d


def plan_model(cfg: dict, ch: int = 3) -> Tuple[List[dict], List[int]]:
    """Channel bookkeeping of parse_model for the in-scope layer types.  Returns a list of
    layer records {i, f, n, type, args(c1,c2,...)} and the save list."""
    gd, gw = cfg["depth_multiple"], cfg["width_multiple"]
    chs: List[int] = [ch]
    layers, save = [], []
    for i, (f, n, m, args) in enumerate(cfg["backbone"] + cfg["head"]):
        args = [cfg["nc"] if a == "nc" else (cfg["anchors"] if a == "anchors" else a) for a in args]
        args = [None if a == "None" else a for a in args]
        n = max(round(n * gd), 1) if n > 1 else n
        if m in ("Conv", "Conv_1", "Conv_2", "Conv_B", "BasicBlock_1", "BasicBlock_2", "Concat_res2",
                 "BasicBlock_ms", "ConcatBlock_ms"):
            c1, c2 = chs[f], _make_divisible(args[0] * gw, 8)
            args = [c1, c2, *args[1:]]
        elif m == "Concat":
            c2 = sum(chs[x] for x in f)
        elif m in ("Detect", "DDetect"):
            args = [*args, [chs[x] for x in f]]
            c2 = chs[f[-1]]
        else:
            c2 = chs[f]
        layers.append(dict(i=i, f=f, n=n, type=m, args=args))
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        if i == 0:
            chs = []
        chs.append(c2)
    return layers, sorted(save)


def _run_layer(sd, L, p, x, training, T_dedupe, rec):
    t, a = L["type"], L["args"]
    if t == "Conv_1":
        return conv_1(sd, p, x, a[2], a[3], training)
    if t == "BasicBlock_2":
        k = a[2] if len(a) > 2 else 3
        s = a[3] if len(a) > 3 else 1
        return basic_block(sd, p, x, a[0], a[1], k, s, training, rec)
    if t == "BasicBlock_1":
        s = a[2] if len(a) > 2 else 1
        return basic_block(sd, p, x, a[0], a[1], 3, s, training, rec)
    if t == "Concat_res2":
        k = a[2] if len(a) > 2 else 3
        s = a[3] if len(a) > 3 else 1
        return concat_res2(sd, p, x, a[0], a[1], k, s, training, rec)
    if t in ("BasicBlock_ms", "ConcatBlock_ms"):
        k = a[2] if len(a) > 2 else 3
        s = a[3] if len(a) > 3 else 1
        fn = basic_block_ms if t == "BasicBlock_ms" else concat_block_ms
        return fn(sd, p, x, a[0], a[1], k, s, training, rec)
    if t == "Conv":
        return conv_silu(sd, p, x, a[2], a[3], 1, training)
    if t in ("Conv_B", "Conv_2"):
        return conv_b(sd, p, x, a[2], a[3] if len(a) > 3 else 1, 1, training, rec)
    if t == "Sample":
        return sample_nearest(x, a[1])
    if t == "Concat":
        return torch.cat(x, a[0])
    raise NotImplementedError(t)


def forward(cfg: dict, sd: Dict[str, torch.Tensor], x: torch.Tensor, T: int, training: bool,
            stride: Optional[torch.Tensor] = None, anchors: Optional[torch.Tensor] = None,
            rec: Optional[dict] = None, ddetect_double_eval: bool = True):
    """Model.forward / DetectionModel.forward.  x: [N,3,H,W] (replicated T times, yolo.py:248-251)
    or [T,N,C,H,W] (event frames straight into _forward_once).  ``anchors``: grid-unit anchors for
    Detect ([nl,na,2]), i.e. the module's ``anchors`` buffer; ``stride``: the head strides."""
    layers, save = plan_model(cfg, x.shape[-3])
    if x.dim() == 4:
        inp = torch.zeros(T, *x.shape, dtype=x.dtype, device=x.device)
        for t in range(T):
            inp[t] = x
        x = inp
    ys = []
    for L in layers:
        f = L["f"]
        if f != -1:
            x = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
        p = f"model.{L['i']}."
        if L["type"] == "Detect":
            x = detect_a(sd, p, list(x), L["args"][0], anchors if anchors is not None else sd[p + "anchors"],
                         stride, training)
        elif L["type"] == "DDetect":
            x = ddetect(sd, p, list(x), L["args"][0], stride, training, ddetect_double_eval, rec)
        elif L["n"] > 1:
            for j in range(L["n"]):
                x = _run_layer(sd, L, f"{p}{j}.", x, training, False, rec)
        else:
            x = _run_layer(sd, L, p, x, training, False, rec)
        ys.append(x if L["i"] in save else None)
        if rec is not None and torch.is_tensor(x):
            rec[f"layer{L['i']}"] = x
    return x


# --------------------------------------------------------------------------------------
# state-dict construction (default init of the reference classes, for benches with no checkpoint)
# --------------------------------------------------------------------------------------
def init_state_dict(cfg: dict, T: int, ch: int = 3, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's initialisers (kaiming-uniform convs via
    nn.Conv2d.reset_parameters; tdBN gamma = thresh or 0.2*thresh, common.py:694-700,753-758).
    Key names and shapes equal the reference model's state_dict; values are NOT bit-equal to a
    reference build under the same seed (module construction order differs) -- parity tests copy
    a state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k, bias, groups=1):
        fan_in = (ci // groups) * k * k
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        sd[name + "weight"] = (torch.rand(co, ci // groups, k, k, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + "bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def lif(name, c):
        conv(name + "spread.0.", c, c, 3, True, c)
        conv(name + "spread.1.", c, c, 1, True)

    def bn(name, c, gamma):
        sd[name + "weight"] = torch.full((c,), gamma)
        sd[name + "bias"] = torch.zeros(c)
        sd[name + "running_mean"] = torch.zeros(c)
        sd[name + "running_var"] = torch.ones(c)
        sd[name + "num_batches_tracked"] = torch.tensor(0)

    def residual(p, cin, hid, cout, k):
        r = p + "residual_function."
        lif(r + "0.", cin)
        conv(r + "1.", hid, cin, k, False)
        bn(r + "2.bn.", hid, THRESH)
        lif(r + "3.", hid)
        conv(r + "4.", cout, hid, k, False)
        bn(r + "5.bn.", cout, 0.2 * THRESH)

    def convb(p, c1, c2, k, g=1):
        lif(p + "act.", c1)
        conv(p + "conv.", c2, c1, k, False, g)
        bn(p + "bn.bn.", c2, THRESH)

    layers, _ = plan_model(cfg, ch)
    for L in layers:
        t, a = L["type"], L["args"]
        for j in range(L["n"]):
            p = f"model.{L['i']}." + (f"{j}." if L["n"] > 1 else "")
            if t == "Conv_1":
                conv(p + "conv.", a[1], a[0], a[2], False)
                bn(p + "bn.bn.", a[1], THRESH)
            elif t in ("BasicBlock_2", "BasicBlock_1"):
                k = (a[2] if len(a) > 2 else 3) if t == "BasicBlock_2" else 3
                s = (a[3] if len(a) > 3 else 1) if t == "BasicBlock_2" else (a[2] if len(a) > 2 else 1)
                cin = a[0] if j == 0 else a[1]
                hid = 1024 if t == "BasicBlock_1" else a[1]
                residual(p, cin, hid, a[1], k)
                if s != 1 or cin != a[1]:
                    lif(p + "shortcut.1.", cin)
                    conv(p + "shortcut.2.", a[1], cin, 1, False)
                    bn(p + "shortcut.3.bn.", a[1], THRESH)
            elif t == "Concat_res2":
                k = a[2] if len(a) > 2 else 3
                residual(p, a[0], a[1], a[1], k)
                if a[0] < a[1]:
                    lif(p + "shortcut.0.", a[0])
                    conv(p + "shortcut.1.", a[1] - a[0], a[0], 1, False)
                    bn(p + "shortcut.2.bn.", a[1] - a[0], THRESH)
            elif t == "BasicBlock_ms":
                k = a[2] if len(a) > 2 else 3
                s = a[3] if len(a) > 3 else 1
                residual(p, a[0], int(a[1] * 0.5), a[1], k)
                if s != 1 or a[0] != a[1]:
                    conv(p + "shortcut.1.", a[1], a[0], 1, False)
                    bn(p + "shortcut.2.bn.", a[1], THRESH)
            elif t == "ConcatBlock_ms":
                k = a[2] if len(a) > 2 else 3
                residual(p, a[0], int(a[1] * 0.5), a[1], k)
                if a[0] < a[1]:
                    lif(p + "shortcut.0.", a[0])
                    conv(p + "shortcut.1.", a[1] - a[0], a[0], 1, False)
                    bn(p + "shortcut.2.bn.", a[1] - a[0], THRESH)
            elif t == "Conv":
                conv(p + "conv.", a[1], a[0], a[2], False)
                bn(p + "bn.bn.", a[1], THRESH)
                lif(p + "act.", a[1])
            elif t in ("Conv_B", "Conv_2"):
                convb(p, a[0], a[1], a[2])
            elif t == "Detect":
                nc, anchors, chs = a[0], a[1], a[2]
                na = len(anchors[0]) // 2
                for i, c in enumerate(chs):
                    conv(f"{p}m.{i}.", na * (nc + 5), c, 1, True)
                    bound = 1.0 / math.sqrt(T)
                    sd[f"{p}w.{i}.conv.weight"] = (torch.rand(1, T, 1, 1, 1, generator=g) * 2 - 1) * bound
                sd[p + "anchors"] = torch.tensor(anchors).float().view(len(anchors), -1, 2)
            elif t == "DDetect":
                nc, chs = a[0], a[1]
                c2 = _make_divisible(max((chs[0] // 4, 64, 16)), 4)
                c3 = max((chs[0], min((nc * 2, 128))))
                for i, c in enumerate(chs):
                    convb(f"{p}cv2.{i}.0.", c, c2, 3)
                    convb(f"{p}cv2.{i}.1.", c2, c2, 3, 4)
                    conv(f"{p}cv2.{i}.2.", 64, c2, 1, True, 4)
                    convb(f"{p}cv3.{i}.0.", c, c3, 3)
                    convb(f"{p}cv3.{i}.1.", c3, c3, 3)
                    conv(f"{p}cv3.{i}.2.", nc, c3, 1, True)
                sd[p + "dfl.conv.weight"] = torch.arange(16, dtype=torch.float).view(1, 16, 1, 1)
    return sd


def detect_strides(cfg: dict, ch: int = 3, s: int = 256) -> torch.Tensor:
    """Strides the reference measures with a 256x256 probe forward (yolo.py:228): with the in-scope
    blocks every stride-2 layer halves the map, so the stride follows from the layer plan."""
    layers, _ = plan_model(cfg, ch)
    red = []
    for L in layers:
        f = L["f"]
        src = f if isinstance(f, int) else f[0]
        base = 1 if L["i"] == 0 else (red[src] if src != -1 else red[-1])
        t, a = L["type"], L["args"]
        if t == "Conv_1" or t == "Conv" or t in ("Conv_B", "Conv_2"):
            r = base * a[3] if len(a) > 3 else base
        elif t in ("BasicBlock_2", "Concat_res2", "BasicBlock_ms", "ConcatBlock_ms"):
            r = base * (a[3] if len(a) > 3 else 1)
        elif t == "BasicBlock_1":
            r = base * (a[2] if len(a) > 2 else 1)
        elif t == "Sample":
            r = base // a[1]
        elif t in ("Detect", "DDetect"):
            return torch.tensor([float(red[x]) for x in f])
        else:
            r = base
        red.append(r)
    raise ValueError("no detection head in cfg")


def firing_rates(rec: dict) -> Dict[str, float]:
    return {k: float(v.mean()) for k, v in rec.items() if not k.startswith("layer")}
