"""Deterministic inputs / weights for the golden parity cases.

Shared by oracle/gen_golden.py (which runs the UNMODIFIED reference on these tensors in the build
container and stores its outputs under tests/golden/*.pt) and by the tests (which rebuild the same
tensors from the seed and compare oracle / CUDA results with the stored reference outputs).
Only seeds and reference OUTPUTS are committed; a checksum of the regenerated inputs guards against
RNG drift between torch versions.
"""
from __future__ import annotations

import math
import os
from typing import Dict

import torch

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))


def gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def randn(g, *shape, scale=1.0, shift=0.0):
    return torch.randn(*shape, generator=g) * scale + shift


def uniform(g, *shape, lo=-1.0, hi=1.0):
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def reseed_state_dict(sd: Dict[str, torch.Tensor], seed: int) -> Dict[str, torch.Tensor]:
    """Fill a reference-shaped state dict with seeded, non-degenerate values (sorted key order)."""
    g = gen(seed)
    out = {}
    for k in sorted(sd):
        v = sd[k]
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(3)
        elif k.endswith("running_mean"):
            out[k] = uniform(g, *v.shape, lo=-0.2, hi=0.2)
        elif k.endswith("running_var"):
            out[k] = uniform(g, *v.shape, lo=0.5, hi=1.5)
        elif ".bn." in k and k.endswith("weight"):
            out[k] = uniform(g, *v.shape, lo=0.3, hi=0.7)
        elif ".bn." in k and k.endswith("bias"):
            out[k] = uniform(g, *v.shape, lo=-0.1, hi=0.1)
        elif k.endswith("anchors") or "dfl" in k:
            out[k] = v.clone()
        elif k.endswith("w.0.conv.weight") or (".w." in k and k.endswith("conv.weight")):
            out[k] = uniform(g, *v.shape, lo=0.1, hi=0.5)
        elif v.dim() == 4:  # conv weight [co, ci/g, k, k]
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            b = 1.5 / math.sqrt(fan_in)
            out[k] = uniform(g, *v.shape, lo=-b, hi=b)
        elif v.dim() == 1:  # conv bias
            out[k] = uniform(g, *v.shape, lo=-0.1, hi=0.1)
        else:
            out[k] = uniform(g, *v.shape)
    return out


def checksum(*tensors) -> float:
    s = 0.0
    for t in tensors:
        t = t.double().flatten()
        s += float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0)).sum())
    return s


def sd_checksum(sd) -> float:
    return checksum(*[sd[k].float() for k in sorted(sd)])


def pack_spikes(s: torch.Tensor) -> torch.Tensor:
    """{0,1} float tensor -> uint8 bit-packed (storage only)."""
    import numpy as np
    return torch.from_numpy(np.packbits(s.numpy().astype("uint8").reshape(-1)))


def unpack_spikes(p: torch.Tensor, shape) -> torch.Tensor:
    import numpy as np
    n = 1
    for d in shape:
        n *= d
    return torch.from_numpy(np.unpackbits(p.numpy())[:n].astype("float32")).reshape(*shape)


def block_cfg(kind: str, cout: int, k: int = 3, s: int = 1):
    """Single-layer model dict understood by the oracle's plan_model/init_state_dict."""
    args = [cout, s] if kind == "BasicBlock_1" else [cout, k, s]
    return dict(nc=3, depth_multiple=1.0, width_multiple=1.0, anchors=2,
                backbone=[[-1, 1, kind, args]], head=[])


# name -> spec.  T, N small; channel counts are multiples of 64 like every in-scope layer.
LIF_CASES = {
    "lif_c64_t4": dict(T=4, N=2, C=64, H=10, W=12, seed=101),
    "lif_c128_t4": dict(T=4, N=1, C=128, H=7, W=9, seed=102),
    "lif_c64_t1": dict(T=1, N=2, C=64, H=6, W=6, seed=103),
    "lif_c64_t5": dict(T=5, N=1, C=64, H=8, W=8, seed=104),
    "lif_c64_t8": dict(T=8, N=1, C=64, H=5, W=7, seed=105),
}
CONV_CASES = {
    # binary (spike) inputs
    "conv3_s1_64_128": dict(T=4, N=2, ci=64, co=128, k=3, s=1, p=1, H=10, W=12, spikes=True, seed=201),
    "conv3_s2_64_64": dict(T=4, N=2, ci=64, co=64, k=3, s=2, p=1, H=12, W=10, spikes=True, seed=202),
    "conv1_s1_128_64": dict(T=4, N=2, ci=128, co=64, k=1, s=1, p=0, H=6, W=8, spikes=True, seed=203),
    "conv3_s1_192_256": dict(T=2, N=1, ci=192, co=256, k=3, s=1, p=1, H=9, W=9, spikes=True, seed=204),
    # real inputs
    "stem7_s2_3_64": dict(T=4, N=2, ci=3, co=64, k=7, s=2, p=3, H=32, W=32, spikes=False, seed=205),
    "head1_128_24_bias": dict(T=4, N=2, ci=128, co=24, k=1, s=1, p=0, H=8, W=8, spikes=False, bias=True, seed=206),
}
BN_CASES = {
    "bn_c64": dict(T=4, N=2, C=64, H=6, W=8, seed=301),
}
BLOCK_CASES = {
    "bb2_64_128_s2": dict(kind="BasicBlock_2", cin=64, cout=128, k=3, s=2, T=4, N=2, H=12, W=12, seed=401),
    "bb2_128_128_s1": dict(kind="BasicBlock_2", cin=128, cout=128, k=3, s=1, T=4, N=2, H=8, W=8, seed=402),
    "bb2_128_64_k1": dict(kind="BasicBlock_2", cin=128, cout=64, k=1, s=1, T=4, N=2, H=8, W=8, seed=403),
    "cr2_64_128_s2": dict(kind="Concat_res2", cin=64, cout=128, k=3, s=2, T=4, N=2, H=12, W=12, seed=404),
    "bb1_128_64_s1": dict(kind="BasicBlock_1", cin=128, cout=64, k=3, s=1, T=4, N=1, H=6, W=6, seed=405),
}
# res*-ee.yaml blocks (the original EMS-YOLO topology): narrow 3 / 32-channel front, hidden width 0.5*cout,
# neuron-free shortcut conv
MS_BLOCK_CASES = {
    "conv2_3_32_s2": dict(kind="Conv_2", cin=3, cout=32, k=3, s=2, T=4, N=2, H=16, W=16, seed=411),
    "cb_ms_32_64_s2": dict(kind="ConcatBlock_ms", cin=32, cout=64, k=3, s=2, T=4, N=2, H=12, W=12, seed=412),
    "cb_ms_128_256_s1": dict(kind="ConcatBlock_ms", cin=128, cout=256, k=3, s=1, T=4, N=1, H=8, W=8, seed=413),
    "bb_ms_64_128_s2": dict(kind="BasicBlock_ms", cin=64, cout=128, k=3, s=2, T=4, N=2, H=12, W=12, seed=414),
    "bb_ms_128_128_s1": dict(kind="BasicBlock_ms", cin=128, cout=128, k=3, s=1, T=4, N=2, H=8, W=8, seed=415),
    "bb_ms_256_128_k1": dict(kind="BasicBlock_ms", cin=256, cout=128, k=1, s=1, T=4, N=2, H=6, W=6, seed=416),
}
MODEL_CASES = {
    "tiny_64": dict(cfg="tiny", T=4, N=2, H=64, W=64, seed=501),
    "tiny_ee_64": dict(cfg="tiny_ee", T=4, N=2, H=64, W=64, seed=502),
}
# Stack B
SILU_CASES = {
    "silu_c64_t4_inplace": dict(T=4, N=2, C=64, H=6, W=8, seed=601, inplace=True),
    "silu_c64_t4": dict(T=4, N=1, C=64, H=5, W=5, seed=602, inplace=False),
}
CONVSILU_CASES = {
    "convsilu_128_64": dict(T=4, N=2, cin=128, cout=64, k=3, s=1, H=6, W=6, seed=611),
}
DDETECT_CASES = {
    "ddetect_128": dict(T=4, N=2, ch=(128, 128), nc=3, H=8, W=8, seed=621),
}
MODEL_B_CASES = {
    "tiny_b_64": dict(cfg="tiny_b", T=4, N=2, H=64, W=64, seed=631),
}


def lif_inputs(spec):
    g = gen(spec["seed"])
    C = spec["C"]
    x = randn(g, spec["T"], spec["N"], C, spec["H"], spec["W"], scale=0.5, shift=0.1)
    dw_w = uniform(g, C, 1, 3, 3, lo=-1 / 3, hi=1 / 3)
    dw_b = uniform(g, C, lo=-1 / 3, hi=1 / 3)
    b = 1.0 / math.sqrt(C)
    pw_w = uniform(g, C, C, 1, 1, lo=-b, hi=b)
    pw_b = uniform(g, C, lo=-b, hi=b)
    gout = randn(g, *x.shape)
    return dict(x=x, dw_w=dw_w, dw_b=dw_b, pw_w=pw_w, pw_b=pw_b, gout=gout)


def conv_inputs(spec):
    g = gen(spec["seed"])
    shape = (spec["T"], spec["N"], spec["ci"], spec["H"], spec["W"])
    if spec["spikes"]:
        x = (torch.rand(*shape, generator=g) < 0.2).float()
    else:
        x = torch.rand(*shape, generator=g)
    fan_in = spec["ci"] * spec["k"] ** 2
    b = 1.0 / math.sqrt(fan_in)
    w = uniform(g, spec["co"], spec["ci"], spec["k"], spec["k"], lo=-b, hi=b)
    bias = uniform(g, spec["co"], lo=-b, hi=b) if spec.get("bias") else None
    return dict(x=x, w=w, b=bias)


def bn_inputs(spec):
    g = gen(spec["seed"])
    C = spec["C"]
    x = randn(g, spec["T"], spec["N"], C, spec["H"], spec["W"], scale=1.3, shift=0.4)
    sd = {"bn.weight": uniform(g, C, lo=0.3, hi=0.7), "bn.bias": uniform(g, C, lo=-0.1, hi=0.1),
          "bn.running_mean": uniform(g, C, lo=-0.2, hi=0.2), "bn.running_var": uniform(g, C, lo=0.5, hi=1.5),
          "bn.num_batches_tracked": torch.tensor(3)}
    return dict(x=x, sd=sd)


def block_inputs(spec, oracle):
    """oracle: the ecs_oracle module (for key/shape bookkeeping only)."""
    cfg = block_cfg(spec["kind"], spec["cout"], spec["k"], spec["s"])
    sd = reseed_state_dict(oracle.init_state_dict(cfg, spec["T"], ch=spec["cin"]), spec["seed"])
    g = gen(spec["seed"] + 7)
    x = randn(g, spec["T"], spec["N"], spec["cin"], spec["H"], spec["W"], scale=0.5, shift=0.1)
    return dict(cfg=cfg, sd=sd, x=x)


def model_inputs(spec, oracle, cfg):
    sd = reseed_state_dict(oracle.init_state_dict(cfg, spec["T"]), spec["seed"])
    stride = oracle.detect_strides(cfg)
    for k in sd:  # Detect keeps its anchors in grid units (models/yolo.py:230)
        if k.endswith("anchors"):
            sd[k] = sd[k] / stride.view(-1, 1, 1)
    g = gen(spec["seed"] + 7)
    x = torch.rand(spec["N"], 3, spec["H"], spec["W"], generator=g)
    return dict(sd=sd, x=x, stride=stride)


def convsilu_inputs(spec, oracle):
    cfg = dict(nc=3, depth_multiple=1.0, width_multiple=1.0, anchors=2,
               backbone=[[-1, 1, "Conv", [spec["cout"], spec["k"], spec["s"]]]], head=[])
    sd = reseed_state_dict(oracle.init_state_dict(cfg, spec["T"], ch=spec["cin"]), spec["seed"])
    g = gen(spec["seed"] + 7)
    x = randn(g, spec["T"], spec["N"], spec["cin"], spec["H"], spec["W"], scale=0.6, shift=0.1)
    return dict(cfg=cfg, sd=sd, x=x)


def ddetect_inputs(spec, oracle):
    """State dict of a bare DDetect head under the prefix 'model.0.' plus real-valued P4/P5 features."""
    cfg = dict(nc=spec["nc"], depth_multiple=1.0, width_multiple=1.0, anchors=2, backbone=[],
               head=[[[-1, -1], 1, "DDetect", ["nc"]]])
    # build the key set by hand: plan_model needs channel bookkeeping for the 'from' layers
    full = dict(nc=spec["nc"], depth_multiple=1.0, width_multiple=1.0, anchors=2,
                backbone=[[-1, 1, "Conv_1", [spec["ch"][0], 1, 1]], [-1, 1, "Conv_1", [spec["ch"][1], 1, 1]]],
                head=[[[0, 1], 1, "DDetect", ["nc"]]])
    sd_all = oracle.init_state_dict(full, spec["T"], ch=3)
    sd = {k.replace("model.2.", "model.0."): v for k, v in sd_all.items() if k.startswith("model.2.")}
    sd = reseed_state_dict(sd, spec["seed"])
    g = gen(spec["seed"] + 7)
    feats = [randn(g, spec["T"], spec["N"], spec["ch"][0], spec["H"], spec["W"], scale=0.6, shift=0.1),
             randn(g, spec["T"], spec["N"], spec["ch"][1], spec["H"] // 2, spec["W"] // 2, scale=0.6, shift=0.1)]
    return dict(sd=sd, feats=feats, stride=torch.tensor([8.0, 16.0]))


# ---------------------------------------------------------------------------------------------
# post-decode NMS, optimizer + EMA step
# ---------------------------------------------------------------------------------------------
NMS_CASES = {
    "nms_basic": dict(N=3, R=600, nc=3, conf=0.25, iou=0.45, seed=701),
    "nms_multi": dict(N=2, R=500, nc=4, conf=0.001, iou=0.6, multi_label=True, seed=702),
    "nms_agnostic_classes": dict(N=2, R=400, nc=5, conf=0.2, iou=0.5, agnostic=True, classes=[0, 2, 3], seed=703),
    "nms_maxdet": dict(N=2, R=800, nc=2, conf=0.05, iou=0.3, max_det=20, seed=704),
    "nms_empty": dict(N=3, R=200, nc=3, conf=0.6, iou=0.45, seed=705, empty_image=1),
    "nms_nc1": dict(N=2, R=300, nc=1, conf=0.1, iou=0.45, multi_label=True, seed=706),
    "nms_big": dict(N=1, R=6000, nc=13, conf=0.001, iou=0.6, multi_label=True, seed=707),   # > max_nms candidates
}


def nms_inputs(spec) -> torch.Tensor:
    """Detect-shaped decoded predictions [N, R, 5 + nc]: jittered boxes around a few objects per image (so that
    suppression actually happens) plus background boxes; objectness / class scores in (0, 1)."""
    g = gen(spec["seed"])
    N, R, nc = spec["N"], spec["R"], spec["nc"]
    n_obj = 6
    centers = torch.rand(N, n_obj, 2, generator=g) * 520 + 60
    sizes = torch.rand(N, n_obj, 2, generator=g) * 150 + 30
    which = torch.randint(0, n_obj + 2, (N, R), generator=g)          # the last two ids = background
    pred = torch.zeros(N, R, 5 + nc)
    bg = which >= n_obj
    idx = which.clamp(max=n_obj - 1)
    c = torch.gather(centers, 1, idx.unsqueeze(-1).expand(-1, -1, 2))
    s = torch.gather(sizes, 1, idx.unsqueeze(-1).expand(-1, -1, 2))
    jitter = torch.randn(N, R, 2, generator=g) * 6
    scale = 1 + torch.randn(N, R, 2, generator=g) * 0.08
    pred[..., 0:2] = torch.where(bg.unsqueeze(-1), torch.rand(N, R, 2, generator=g) * 640, c + jitter)
    pred[..., 2:4] = torch.where(bg.unsqueeze(-1), torch.rand(N, R, 2, generator=g) * 100 + 8, s * scale)
    obj = torch.rand(N, R, generator=g)
    pred[..., 4] = torch.where(bg, obj * 0.3, obj)
    pred[..., 5:] = torch.rand(N, R, nc, generator=g)
    if spec.get("empty_image") is not None:
        pred[spec["empty_image"], :, 4] = 0.0
    return pred


OPT_CASE = dict(lr=0.01, momentum=0.937, weight_decay=0.0005, steps=3, seed=801)


class _OptNet(torch.nn.Module):
    """The three parameter kinds train.py:259-270 sorts into groups: BatchNorm3d weights (g0, no decay), other
    weights (g1, decay), biases (g2); plus BN buffers, which only the EMA touches."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(4, 8, 3, bias=True)
        self.bn = torch.nn.BatchNorm3d(8)
        self.head = torch.nn.Conv2d(8, 5, 1, bias=True)
        self.big = torch.nn.Conv2d(16, 33, 3, bias=False)     # 4752 elements: more than one chunk, odd tail


def opt_inputs(spec):
    torch.manual_seed(spec["seed"])
    m = _OptNet()
    g = gen(spec["seed"] + 1)
    with torch.no_grad():
        m.bn.weight.copy_(uniform(g, 8, lo=0.3, hi=0.7))
        m.bn.running_var.copy_(uniform(g, 8, lo=0.5, hi=1.5))
    grads = [[randn(g, *p.shape, scale=0.1) for p in m.parameters()] for _ in range(spec["steps"])]
    return m, grads


def opt_groups(model):
    """train.py:259-270"""
    g0, g1, g2 = [], [], []
    for v in model.modules():
        if hasattr(v, 'bias') and isinstance(v.bias, torch.nn.Parameter):
            g2.append(v.bias)
        if isinstance(v, torch.nn.BatchNorm3d):
            g0.append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, torch.nn.Parameter):
            g1.append(v.weight)
    return g0, g1, g2


# Gen1 event-camera input (g1-resnet): T bins of (x, y, p) events of a 304 x 240 sensor per sample
EVENT_CASES = {
    "ev_t5_320": dict(N=3, T=5, n_events=40000, out=320, seed=901),
    "ev_t5_640": dict(N=2, T=5, n_events=15000, out=640, seed=902),
    "ev_t4_sparse": dict(N=2, T=4, n_events=300, out=320, seed=903, empty_bin=(1, 2)),
}
EV_W, EV_H = 304, 240


def event_inputs(spec):
    """-> list over samples of list over T bins of dict(x, y, p) int64 arrays, in sensor (time) order.  Events
    cluster around a few moving blobs so that many pixels receive several events (last write wins)."""
    g = gen(spec["seed"])
    samples = []
    for n in range(spec["N"]):
        bins = []
        for t in range(spec["T"]):
            k = int(torch.randint(spec["n_events"] // 2, spec["n_events"], (1,), generator=g))
            if spec.get("empty_bin") == (n, t):
                k = 0
            c = torch.rand(4, 2, generator=g) * torch.tensor([EV_W - 40.0, EV_H - 40.0]) + 20
            which = torch.randint(0, 5, (k,), generator=g)
            pos = torch.where((which < 4).unsqueeze(-1), c[which.clamp(max=3)] + torch.randn(k, 2, generator=g) * 9,
                              torch.rand(k, 2, generator=g) * torch.tensor([float(EV_W), float(EV_H)]))
            x = pos[:, 0].round().long().clamp(0, EV_W - 1)
            y = pos[:, 1].round().long().clamp(0, EV_H - 1)
            p = torch.randint(0, 2, (k,), generator=g)
            bins.append(dict(x=x, y=y, p=p))
        samples.append(bins)
    return samples


def zpack(t: torch.Tensor) -> dict:
    """uint8 tensor -> zlib-compressed bytes (event frames are mostly grey: ~50x smaller fixtures)."""
    import zlib
    assert t.dtype == torch.uint8
    return dict(z=zlib.compress(t.contiguous().numpy().tobytes(), 9), shape=tuple(t.shape))


def zunpack(d: dict) -> torch.Tensor:
    import zlib
    return torch.frombuffer(bytearray(zlib.decompress(d["z"])), dtype=torch.uint8).reshape(*d["shape"]).clone()


# ---------------------------------------------------------------------------------------------
# Stack-A training loss (utils/loss.py ComputeLoss, SIoU box term)
# ---------------------------------------------------------------------------------------------
_ANCH2 = [[[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]], [[1.875, 3.8125], [3.875, 2.8125], [3.6875, 7.4375]]]
_ANCH3 = _ANCH2 + [[[3.625, 2.8125], [4.875, 6.1875], [11.65625, 10.1875]]]
_HYP = dict(box=0.05, obj=1.0, cls=0.5, cls_pw=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0,
            label_smoothing=0.0)
LOSS_CASES = {
    "loss_basic": dict(N=4, nc=13, grids=[(12, 16), (6, 8)], anchors=_ANCH2, nt=24, seed=801, hyp=_HYP),
    "loss_empty": dict(N=2, nc=13, grids=[(8, 8), (4, 4)], anchors=_ANCH2, nt=0, seed=802, hyp=_HYP),
    "loss_dups": dict(N=2, nc=4, grids=[(8, 8), (4, 4)], anchors=_ANCH2, nt=40, seed=803, hyp=_HYP, cluster=True),
    "loss_nc1": dict(N=3, nc=1, grids=[(10, 10), (5, 5)], anchors=_ANCH2, nt=15, seed=804, hyp=_HYP),
    "loss_nl3_smooth": dict(N=2, nc=6, grids=[(16, 16), (8, 8), (4, 4)], anchors=_ANCH3, nt=30, seed=805,
                            hyp=dict(_HYP, box=0.0375, obj=0.7, cls=0.3, cls_pw=1.5, obj_pw=0.8, anchor_t=3.0,
                                     label_smoothing=0.1)),
    "loss_border": dict(N=2, nc=3, grids=[(8, 12), (4, 6)], anchors=_ANCH2, nt=32, seed=806, hyp=_HYP, border=True),
    # the wrapped criteria (utils/loss.py:38-106): FocalLoss, and the stateful SlideLoss over two consecutive calls
    "loss_focal": dict(N=3, nc=5, grids=[(10, 10), (5, 5)], anchors=_ANCH2, nt=18, seed=808, hyp=dict(_HYP, fl_gamma=1.5)),
    "loss_slide": dict(N=3, nc=5, grids=[(10, 10), (5, 5)], anchors=_ANCH2, nt=18, seed=809, hyp=dict(_HYP, slide_ratio=1.0),
                       calls=2),
    "loss_slide_mixed": dict(N=2, nc=4, grids=[(8, 8), (4, 4)], anchors=[_ANCH2[0], [[30.0, 30.0], [40.0, 40.0], [50.0, 50.0]]],
                             nt=14, seed=810, hyp=dict(_HYP, slide_ratio=1.0), calls=3),   # level 1 never matches: auto_iou 0.5
    "loss_unmatched": dict(N=2, nc=3, grids=[(8, 8), (4, 4)], anchors=_ANCH2, nt=6, seed=807, hyp=_HYP, tiny_boxes=True),
}


def loss_inputs(spec):
    """Raw Detect training outputs p[i] [N, na, ny, nx, 5 + nc] (models/yolo.py:141-145 layout), anchors in grid units
    (models/yolo.py:230) and targets [nt, 6] = (image, class, cx, cy, w, h) normalised (utils/datasets.py:615-630)."""
    g = gen(spec["seed"])
    N, nc = spec["N"], spec["nc"]
    anchors = torch.tensor(spec["anchors"], dtype=torch.float32)
    na = anchors.shape[1]
    p = [randn(g, N, na, ny, nx, 5 + nc, scale=1.5) for ny, nx in spec["grids"]]
    nt = spec["nt"]
    img = torch.randint(0, N, (nt,), generator=g).float()
    cls = torch.randint(0, nc, (nt,), generator=g).float()
    if spec.get("cluster"):          # a few centres shared by many boxes: duplicate (image, anchor, cell) indices
        centres = torch.rand(4, 2, generator=g) * 0.6 + 0.2
        cxy = centres[torch.randint(0, 4, (nt,), generator=g)] + (torch.rand(nt, 2, generator=g) - 0.5) * 0.02
        img = (img % 2)
    elif spec.get("border"):         # centres in the outermost cells and just beyond the half-cell offsets
        cxy = torch.rand(nt, 2, generator=g)
        edge = torch.rand(nt, 2, generator=g)
        cxy = torch.where(edge < 0.35, cxy * 0.06, torch.where(edge > 0.65, 1.0 - cxy * 0.06, cxy))
    else:
        cxy = torch.rand(nt, 2, generator=g) * 0.9 + 0.05
    if spec.get("tiny_boxes"):       # no anchor within anchor_t: every level ends up without a match
        wh = torch.rand(nt, 2, generator=g) * 0.004 + 0.001
    else:
        wh = torch.exp(torch.rand(nt, 2, generator=g) * 3.2 - 3.6)          # 0.027 .. 0.67, log-uniform
    targets = torch.cat([img[:, None], cls[:, None], cxy, wh], 1).reshape(nt, 6)
    gout = float(torch.rand(1, generator=g)) + 0.5                          # upstream gradient of the scalar loss
    return dict(p=p, anchors=anchors, targets=targets, gout=gout)


# ---------------------------------------------------------------------------------------------
# SURVEY section 8f rank 1, Stack B: TAL loss (utils/loss_tal.py ComputeLoss + utils/tal/assigner.py)
# ---------------------------------------------------------------------------------------------
TAL_CASES = {
    "tal_basic": dict(N=3, nc=5, grids=[(8, 8), (4, 4)], strides=[16.0, 32.0], nt=12, seed=901, wh=(0.3, 0.7)),
    "tal_empty": dict(N=2, nc=5, grids=[(8, 8), (4, 4)], strides=[16.0, 32.0], nt=0, seed=902, wh=(0.3, 0.7)),
    "tal_dense": dict(N=2, nc=3, grids=[(10, 10), (5, 5)], strides=[16.0, 32.0], nt=30, seed=903, wh=(0.35, 0.8)),
    "tal_rect_nc1": dict(N=2, nc=1, grids=[(6, 10), (3, 5)], strides=[16.0, 32.0], nt=9, seed=904, wh=(0.4, 0.8)),
    "tal_small_boxes": dict(N=2, nc=4, grids=[(16, 16), (8, 8)], strides=[16.0, 32.0], nt=14, seed=905, wh=(0.04, 0.5)),
    "tal_focal": dict(N=2, nc=4, grids=[(8, 8), (4, 4)], strides=[16.0, 32.0], nt=10, seed=907, wh=(0.3, 0.7), fl_gamma=1.5),
    "tal_assigner_hp": dict(N=2, nc=4, grids=[(10, 10), (5, 5)], strides=[16.0, 32.0], nt=16, seed=908, wh=(0.3, 0.8),
                            assigner=(13, 1.0, 4.0)),           # YOLOM / YOLOA / YOLOB (utils/loss_tal.py:134-137)
    "tal_nl3_uneven": dict(N=4, nc=6, grids=[(16, 16), (8, 8), (4, 4)], strides=[8.0, 16.0, 32.0], nt=20, seed=906,
                           wh=(0.3, 0.9), skip_image=2, smooth=0.0, cls_pw=1.3),
}


def tal_inputs(spec):
    """Raw DDetect training outputs feats[i] [N, 64 + nc, ny, nx] (models/yolo_snn.py:117-119) and targets [nt, 6] =
    (image, class, cx, cy, w, h) normalised.  The box logits are scaled so that the decoded boxes are a few cells wide
    (positive CIoU with the synthetic boxes, i.e. a non-trivial assignment)."""
    g = gen(spec["seed"])
    N, nc = spec["N"], spec["nc"]
    feats = []
    for ny, nx in spec["grids"]:
        f = randn(g, N, 64 + nc, ny, nx, scale=1.0)
        f[:, :64] = f[:, :64] * 1.5 - torch.arange(16.0).repeat(4).view(1, 64, 1, 1) * 0.45   # mass on the low bins
        feats.append(f)
    nt = spec["nt"]
    img = torch.randint(0, N, (nt,), generator=g).float()
    if "skip_image" in spec and nt:
        img = torch.where(img == spec["skip_image"], torch.zeros_like(img), img)     # one image without labels
    cls = torch.randint(0, nc, (nt,), generator=g).float()
    lo, hi = spec["wh"]
    wh = torch.rand(nt, 2, generator=g) * (hi - lo) + lo
    cxy = torch.rand(nt, 2, generator=g) * (1 - wh) + wh / 2                          # boxes inside the image
    targets = torch.cat([img[:, None], cls[:, None], cxy, wh], 1).reshape(nt, 6)
    gout = float(torch.rand(1, generator=g)) + 0.5
    return dict(feats=feats, targets=targets, strides=torch.tensor(spec["strides"]), gout=gout)


# ---------------------------------------------------------------------------------------------
# pickled reference checkpoints (models/experimental.py:87-127)
# ---------------------------------------------------------------------------------------------
CKPT_PLANS = {
    "micro_a": dict(nc=3, depth_multiple=1.0, width_multiple=1.0, anchors=[[10, 14, 23, 27, 37, 58], [81, 82, 135, 169, 344, 319]],
                    backbone=[[-1, 1, "Conv_1", [64, 7, 2]], [-1, 1, "BasicBlock_2", [64, 3, 2]],
                              [-1, 1, "BasicBlock_2", [64, 3, 2]]],
                    head=[[[1, 2], 1, "Detect", ["nc", "anchors"]]]),
    "micro_b": dict(nc=3, depth_multiple=1.0, width_multiple=1.0, anchors=2,
                    backbone=[[-1, 1, "Conv_1", [64, 7, 2]], [-1, 1, "BasicBlock_2", [64, 3, 2]],
                              [-1, 1, "BasicBlock_2", [64, 3, 2]]],
                    head=[[[1, 2], 1, "DDetect", ["nc"]]]),
}


# ---------------------------------------------------------------------------------------------
# Whole model + training loss (north_star: "head outputs and loss agree within 1e-3 relative")
# ---------------------------------------------------------------------------------------------
MODEL_LOSS_HYP = dict(box=0.05 * 3 / 2, cls=0.5 * 3 / 80 * 3 / 2, obj=1.0 * (64 / 640) ** 2 * 3 / 2, cls_pw=1.0, obj_pw=1.0,
                      anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)     # train.py:427-433 scaling

# tier-4 training-curve fixtures (oracle/gen_golden_trajectory.py): the reference's SGD recipe on one fixed batch
TRAJECTORY_HYP = dict(steps=6, lr=0.01, momentum=0.937, weight_decay=5e-4)


def model_targets(spec, nc=3):
    """[nt, 6] labels for the whole-model loss cases: 3 boxes per image, sized to match the tiny plans' anchors."""
    g = gen(spec["seed"] + 31)
    N = spec["N"]
    img = torch.arange(N).repeat_interleave(3).float()
    n = img.numel()
    return torch.cat([img[:, None], torch.randint(0, nc, (n, 1), generator=g).float(),
                      torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.4 + 0.25], 1)
