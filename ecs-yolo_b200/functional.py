"""Tensor-level host wrappers over the C ABI (include/ecsy.h).

PyTorch is plumbing here: device memory (caching allocator), streams, tiny [C]-vector math.  Every
heavy operation is one C-ABI call into libecsy.so; nothing falls back to torch ops or the CPU.

Internal layout: real activations are fp32 NHWC ``[Tp, N, H, W, C]`` with Tp == T or
Tp == 1 (a T-broadcast tensor: the direct-coded image repeats every timestep, models/yolo.py:248-251);
spikes are bit-packed int32 ``[T, N, H, W, C/32]``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _cabi

# models/common.py:37-39 -- module-level "config" of the reference, read at call time
thresh = 0.5
lens = 0.5
decay = 0.25

# conv_ts: "auto" (measured dispatch rule), "all" (wherever supported) or "off"
_state = {"splits": 2, "conv_ts": os.environ.get("ECSY_CONV_TS", "auto"),
          "lif_fused": os.environ.get("ECSY_LIF_FUSED", "0") == "1",
          "lif_wave": {"1": "auto", "0": "off"}.get(os.environ.get("ECSY_LIF_WAVE", "auto"), os.environ.get("ECSY_LIF_WAVE", "auto")),
          "lif_store": os.environ.get("ECSY_LIF_STORE", "1") == "1",
          "dispatch": os.environ.get("ECSY_DISPATCH", "ops"),
          "stem_kernel": os.environ.get("ECSY_STEM_KERNEL", "1") == "1"}

# ---- launch accounting / per-operator CUDA-event timing (used by bench.py) ----
launches = {"n": 0}
launches_by_op = {}
_prof = {"on": False, "events": []}
flops = {"spike_conv": 0.0, "ecs_pw": 0.0, "real_conv": 0.0, "lif_fwd_elems": 0.0, "lif_fwd_in_elems": 0.0, "lif_elems": 0.0,
         "tdbn_elems": 0.0}


def profile_begin():
    _prof["on"], _prof["events"] = True, []


def profile_end():
    """-> {op: milliseconds}; call after a synchronize."""
    _prof["on"] = False
    out = {}
    for name, a, b in _prof["events"]:
        out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    _prof["events"] = []
    return out


class _timed:
    def __init__(self, name, n_launch):
        self.name, self.n = name, n_launch

    def __enter__(self):
        launches["n"] += self.n
        launches_by_op[self.name] = launches_by_op.get(self.name, 0) + self.n
        if _prof["on"]:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _prof["on"]:
            self.b.record()
            _prof["events"].append((self.name, self.a, self.b))
        return False


def bump_weights_epoch() -> None:
    """Called by anything that rewrites parameters behind torch's back (the fused optimizer writes through raw pointers):
    part of every derived-weight cache key (common._cached, ops._cached)."""
    _state["weights_epoch"] = _state.get("weights_epoch", 0) + 1


def weights_epoch() -> int:
    return _state.get("weights_epoch", 0)


def set_precision(mode: str) -> None:
    """'parity': weights as bf16 hi+lo pairs (2 MMAs, ~fp32 weights; spikes are exact in bf16);
    'fast': single bf16 plane."""
    if mode not in ("parity", "fast"):
        raise ValueError(mode)
    _state["splits"] = 2 if mode == "parity" else 1


def get_splits() -> int:
    return _state["splits"]


def set_conv_ts(mode) -> None:
    """Where spike convs stage their A operand: "auto" = tensor memory (tcgen05.mma TS form) except on the wide
    layers where 256-column tiles with a shared-memory operand measured faster; "all" / True = tensor memory
    wherever supported; "off" / False = shared memory."""
    if mode is True:
        mode = "all"
    elif mode is False:
        mode = "off"
    if mode not in ("auto", "all", "off", "pair"):
        raise ValueError(mode)
    _state["conv_ts"] = mode


def set_lif_fused(on: bool) -> None:
    """Fast mode, C == 64: run the ECS-LIF forward as ONE kernel with all T steps on chip.  Parity-green but
    measured slower than the per-timestep pipeline so far (csrc/lif_fused.cu header), hence off by default."""
    _state["lif_fused"] = bool(on)


def set_dispatch(mode: str) -> None:
    """How the drop-in modules reach the kernels on the inference path: "ops" (default) = through the PyTorch custom
    operators torch.ops.ecsy.lif_ecs / torch.ops.ecsy.spike_conv (ops.py: torch.library ops with fake implementations, so
    the modules can be traced / exported), which call this module's C-ABI wrappers; "direct" = the wrappers themselves."""
    if mode not in ("ops", "direct"):
        raise ValueError(mode)
    _state["dispatch"] = mode


def set_lif_wave(mode) -> None:
    """Fast precision, C == 64, 2 <= T <= 4: the ECS-LIF forward as ONE wavefront kernel with the state of all T steps on
    chip (csrc/lif_wave.cu) instead of the per-timestep pipeline.  "auto" (default): where it measured faster
    (ecsy_lif_ecs_wave_prefers); "all" / True: wherever supported; "off" / False: never."""
    mode = {True: "all", False: "off"}.get(mode, mode)
    if mode not in ("auto", "all", "off"):
        raise ValueError(mode)
    _state["lif_wave"] = mode


def set_lif_store(on: bool) -> None:
    """Training: keep the membranes / ECS traces the forward pass computes for the BPTT backward (8 B per element-step,
    ~35 GB for resnet34 at batch 32 -- sized for the B200's 180 GB) instead of recomputing the forward inside the
    backward.  Falls back to recomputation per layer when free device memory is short."""
    _state["lif_store"] = bool(on)
    _store_decisions.clear()


_store_decisions = {}


def lif_store_ok(x: "Act") -> bool:
    if not _state["lif_store"] or x.C % 64:
        return False
    # decided once per (device, layer shape), in forward order of the first training step: cudaMemGetInfo is a slow,
    # driver-serialised call (measured: ~3 ms each with two ranks on one box -- 170 ms per step when asked per layer)
    key = (x.data.device.index, x.T, x.N, x.H, x.W, x.C)
    ok = _store_decisions.get(key)
    if ok is None:
        need = (2 * x.T - 1) * x.N * x.H * x.W * x.C * 4
        free, _ = torch.cuda.mem_get_info(x.data.device)
        # the caching allocator may hold reusable blocks on top of `free`; keep a wide margin for the backward's workspaces
        spare = free + torch.cuda.memory_reserved(x.data.device) - torch.cuda.memory_allocated(x.data.device)
        ok = _store_decisions[key] = bool(spare > 3 * need + (8 << 30))
    return ok


def conv_ts_enabled() -> bool:
    return _state["conv_ts"] != "off"


def _use_ts(ci: int, co: int, splits: int) -> bool:
    m = _state["conv_ts"]
    if m in ("off", "pair"):     # "pair": shared-memory operand kernel on tensor-memory-layout weights everywhere
        return False
    L = _cabi.lib()
    return bool(L.ecsy_spike_conv_ts_supported(ci, co) if m == "all" else L.ecsy_spike_conv_prefers_ts(ci, co, splits))


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda):
            raise RuntimeError("ecs-yolo_b200 operators need CUDA tensors: there is no CPU fallback")


class Act:
    """Real-valued activation in the internal NHWC layout."""
    __slots__ = ("data", "T")

    def __init__(self, data: torch.Tensor, T: int):
        assert data.dim() == 5 and data.dtype == torch.float32 and data.is_contiguous()
        assert data.shape[0] in (1, T)
        self.data, self.T = data, T

    Tp = property(lambda s: s.data.shape[0])
    N = property(lambda s: s.data.shape[1])
    H = property(lambda s: s.data.shape[2])
    W = property(lambda s: s.data.shape[3])
    C = property(lambda s: s.data.shape[4])
    imgs = property(lambda s: s.T * s.data.shape[1])
    src_imgs = property(lambda s: s.data.shape[0] * s.data.shape[1])

    @property
    def tstride(self) -> int:
        return 0 if (self.Tp == 1 and self.T > 1) else self.N * self.H * self.W * self.C

    def to_ref(self) -> torch.Tensor:
        """Reference-shaped logical view [T, N, C, H, W] over the NHWC memory (no copy)."""
        v = self.data.permute(0, 1, 4, 2, 3)
        return v.expand(self.T, -1, -1, -1, -1) if self.Tp != self.T else v

    @staticmethod
    def from_ref(x: torch.Tensor) -> "Act":
        """Accepts a reference tensor [T, N, C, H, W]; zero-copy when it already is an NHWC view."""
        _chk_cuda(x)
        if x.dim() != 5:
            raise ValueError(f"expected [T,N,C,H,W], got {tuple(x.shape)}")
        T = x.shape[0]
        x = x.float() if x.dtype != torch.float32 else x
        src = x[:1] if (x.stride(0) == 0 and T > 1) else x
        nhwc = src.permute(0, 1, 3, 4, 2)
        if nhwc.is_contiguous():
            return Act(nhwc, T)
        src = src.contiguous()
        Tp, N, Cc, H, W = src.shape
        out = torch.empty(Tp, N, H, W, Cc, device=x.device, dtype=torch.float32)
        with _timed("layout", 1):
            _cabi.check(_cabi.lib().ecsy_nchw_to_nhwc_f32(_p(src), _p(out), Tp * N, Cc, H, W, _st()), "nchw_to_nhwc")
        return Act(out, T)

    def full(self) -> "Act":
        """Materialise a T-broadcast tensor."""
        if self.Tp == self.T:
            return self
        return Act(self.data.expand(self.T, -1, -1, -1, -1).contiguous(), self.T)


class Spikes:
    """Bit-packed spike train [T, N, H, W, C/32] (int32 words).  ``Cr`` <= C is the number of channels that
    carry data: layers narrower than the kernels' 64-channel granule (the 3- and 32-channel front of
    res*-ee.yaml) run zero-padded to C, and the padded channels never fire."""
    __slots__ = ("bits", "C", "Cr")

    def __init__(self, bits: torch.Tensor, C: int, Cr: Optional[int] = None):
        self.bits, self.C = bits, C
        self.Cr = C if Cr is None else Cr

    T = property(lambda s: s.bits.shape[0])
    N = property(lambda s: s.bits.shape[1])
    H = property(lambda s: s.bits.shape[2])
    W = property(lambda s: s.bits.shape[3])

    def to_act(self) -> Act:
        T, N, H, W, _ = self.bits.shape
        out = torch.empty(T, N, H, W, self.C, device=self.bits.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().ecsy_spikes_unpack(_p(self.bits), _p(out), T * N * H * W, self.C, _st()), "spikes_unpack")
        if self.Cr != self.C:
            out = out[..., :self.Cr].contiguous()
        return Act(out, T)

    @staticmethod
    def from_act(a: Act, th: Optional[float] = None) -> "Spikes":
        a = a.full()
        bits = torch.empty(a.T, a.N, a.H, a.W, a.C // 32, device=a.data.device, dtype=torch.int32)
        _cabi.check(_cabi.lib().ecsy_spikes_pack(_p(a.data), _p(bits), a.T * a.N * a.H * a.W, a.C,
                                                 thresh if th is None else th, _st()), "spikes_pack")
        return Spikes(bits, a.C)


# ------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------
def pack_conv_weight(w: torch.Tensor, splits: int) -> torch.Tensor:
    """[Co, Ci, kh, kw] fp32 -> [splits, Co, Kpad] bf16 (hi, lo planes), k = (ky*kw+kx)*Ci + ci."""
    _chk_cuda(w)
    w = w.detach().float().contiguous()
    Co, Ci, kh, kw = w.shape
    Kpad = (Ci * kh * kw + 63) // 64 * 64
    out = torch.empty(splits, Co, Kpad, device=w.device, dtype=torch.bfloat16)
    _cabi.check(_cabi.lib().ecsy_pack_conv_weight(_p(w), _p(out), Co, Ci, kh, kw, Kpad, splits, _st()),
                "pack_conv_weight")
    return out


@dataclass
class ConvW:
    """Device-side forms of one Snn_Conv2d weight."""
    packed: Optional[torch.Tensor]  # [splits, Co, Kpad] bf16 for the tcgen05 paths
    simt: Optional[torch.Tensor]    # [kh, kw, Ci/g, Co] fp32 for the SIMT path
    bias: Optional[torch.Tensor]
    co: int
    ci: int
    k: int
    stride: int
    pad: int
    groups: int
    splits: int
    dense_groups: bool = False   # `packed` holds the block-diagonal dense form of a grouped weight
    packed_ts: Optional[torch.Tensor] = None   # spike-conv form for the tensor-memory path (ecsy_pack_spike_conv_weight)
    stem: Optional[torch.Tensor] = None        # real-image stem form for ecsy_stem_conv (pack_stem_weight)


def densify_grouped(weight: torch.Tensor, groups: int) -> torch.Tensor:
    """[Co, Ci/g, kh, kw] grouped weight -> equivalent dense [Co, Ci, kh, kw] (zeros across groups), so a
    grouped spike conv (DDetect cv2, models/yolo_snn.py:100-103) runs on the tensor-core path unchanged."""
    Co, Cig, kh, kw = weight.shape
    dense = torch.zeros(Co, Cig * groups, kh, kw, device=weight.device, dtype=torch.float32)
    cog = Co // groups
    for g in range(groups):
        dense[g * cog:(g + 1) * cog, g * Cig:(g + 1) * Cig] = weight[g * cog:(g + 1) * cog].detach().float()
    return dense


def pack_spike_conv_weight(w: torch.Tensor, splits: int) -> torch.Tensor:
    """[Co, Ci, kh, kw] fp32 -> [splits, Co, kh*kw*Ci] bf16 in the operand order of the tensor-memory spike conv."""
    _chk_cuda(w)
    w = w.detach().float().contiguous()
    Co, Ci, kh, kw = w.shape
    out = torch.empty(splits, Co, kh * kw * Ci, device=w.device, dtype=torch.bfloat16)
    _cabi.check(_cabi.lib().ecsy_pack_spike_conv_weight(_p(w), _p(out), Co, Ci, kh, kw, splits, _st()),
                "pack_spike_conv_weight")
    return out


def pack_stem_weight(weight: torch.Tensor) -> torch.Tensor:
    """[64, Ci <= 4, k, k] -> bf16 [64][ceil(k/2) * 64] for ecsy_stem_conv: K entry kb*64 + part*32 + kx*4 + ci holds
    W[co][ci][2*kb + part][kx] (a kernel ROW is 8 pixels x 4 channels of the NHWC4 image = 32 K entries), zeros elsewhere."""
    Co, Ci, k, _ = weight.shape
    KB = (k + 1) // 2
    t = torch.zeros(Co, 2 * KB, 8, 4, device=weight.device, dtype=torch.float32)
    t[:, :k, :k, :Ci] = weight.detach().float().permute(0, 2, 3, 1)
    return t.reshape(Co, KB * 64).to(torch.bfloat16).contiguous()


def make_conv_w(weight: torch.Tensor, bias, stride: int, pad: int, groups: int, umma: bool, simt: bool,
                densify: bool = False) -> ConvW:
    splits = get_splits()
    Co, Cig, kh, kw = weight.shape
    dense = densify_grouped(weight, groups) if (densify and groups > 1) else weight
    packed = pack_conv_weight(dense, splits) if umma else None
    sw = weight.detach().float().permute(2, 3, 1, 0).contiguous() if simt else None
    b = bias.detach().float().contiguous() if bias is not None else None
    cw = ConvW(packed, sw, b, Co, Cig * groups, kh, stride, pad, groups, splits)
    cw.dense_groups = bool(densify and groups > 1)
    if (packed is not None and bias is None and (groups == 1 or densify)
            and _cabi.lib().ecsy_spike_conv_ts_supported(Cig * groups, Co)):
        cw.packed_ts = pack_spike_conv_weight(dense, splits)
    if umma and bias is None and groups == 1 and kh == kw and _cabi.lib().ecsy_stem_conv_supported(Cig, Co, kh, splits):
        cw.stem = pack_stem_weight(weight)
    return cw


def make_head_conv_w(weight: torch.Tensor, bias: Optional[torch.Tensor], groups: int = 1, bias_mul=1.0):
    """A detection head's 1x1 conv (+ bias) on real features for the dense tcgen05 GEMM: a grouped weight becomes its
    block-diagonal dense form, the output channels are zero-padded to the 64-column tile, and the weight is ALWAYS packed as
    bf16 hi + lo planes (three-term split hi*hi + lo*hi + hi*lo, ~1e-6: head outputs keep fp32 accuracy in fast precision).
    -> (ConvW, scale = 1, shift = bias * bias_mul) for real_conv; the caller slices the first `co` output channels."""
    dense = densify_grouped(weight, groups) if groups > 1 else weight.detach().float()
    co, ci = dense.shape[0], dense.shape[1]
    cop = pad64(co)
    w = torch.nn.functional.pad(dense, (0, 0, 0, 0, 0, 0, 0, cop - co))
    cw = ConvW(pack_conv_weight(w, 2), None, None, cop, ci, 1, 1, 0, 1, 2)
    shift = torch.zeros(cop, device=w.device, dtype=torch.float32)
    if bias is not None:
        shift[:co] = bias.detach().float() * bias_mul
    return cw, torch.ones(cop, device=w.device, dtype=torch.float32), shift


@dataclass
class LifW:
    dw_w: torch.Tensor   # [9, C]
    dw_b: torch.Tensor   # [C]
    pw: torch.Tensor     # [splits, C, C] bf16
    pw_b: torch.Tensor   # [C]
    splits: int
    w_eff: Optional[torch.Tensor] = None    # fused kernel: folded 3x3 spread weight (ecsy_pack_spike_conv_weight form)
    bconst: Optional[torch.Tensor] = None   # fused kernels: pw @ dw_b + pw_b
    w_wave: Optional[torch.Tensor] = None   # wavefront kernel: [9][64][64] bf16 folded spread weight (pack_lif_wave_weight)


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def pad_channels(t: torch.Tensor, Cp: int) -> torch.Tensor:
    """Zero-pad the last (channel) axis of an NHWC tensor / a [C] vector to Cp entries."""
    return t if t.shape[-1] == Cp else torch.nn.functional.pad(t, (0, Cp - t.shape[-1]))


def pack_lif_wave_weight(w_eff: torch.Tensor) -> torch.Tensor:
    """[64 co, 64 ci, 3, 3] fp32 folded spread weight -> [9 taps][64 co][64 kk] bf16 for ecsy_lif_ecs_wave_fwd.  kk is the
    position of channel c in the operand row a thread of the kernel writes: it holds channels 8k + 2q + e (k = 0..7, e =
    0, 1) of a pixel for its q = lane % 4 and stores them at kk = 16q + 2k + e (csrc/lif_wave.cu)."""
    kk = torch.arange(64, device=w_eff.device)
    c_of_kk = 8 * ((kk % 16) // 2) + 2 * (kk // 16) + (kk % 2)
    w = w_eff.detach().float()[:, c_of_kk]                       # [co, kk, ky, kx]
    return w.permute(2, 3, 0, 1).reshape(9, 64, 64).contiguous().to(torch.bfloat16)


def make_lif_w(dw_w, dw_b, pw_w, pw_b, Cp: Optional[int] = None) -> LifW:
    """Cp > C: spread weights zero-padded to Cp channels (the extra channels see zero current, a zero trace
    and never fire), for layers narrower than the 64-channel granule."""
    splits = get_splits()
    C = dw_w.shape[0]
    if Cp is not None and Cp != C:
        dw_w = torch.nn.functional.pad(dw_w.detach().float(), (0, 0, 0, 0, 0, 0, 0, Cp - C))
        dw_b = pad_channels(dw_b.detach().float(), Cp)
        pw_w = torch.nn.functional.pad(pw_w.detach().float(), (0, 0, 0, 0, 0, Cp - C, 0, Cp - C))
        pw_b = pad_channels(pw_b.detach().float(), Cp)
        C = Cp
    w = LifW(dw_w.detach().float().reshape(C, 9).t().contiguous(), dw_b.detach().float().contiguous(),
             pack_conv_weight(pw_w, splits), pw_b.detach().float().contiguous(), splits)
    if splits == 1 and C == 64:
        pw2 = pw_w.detach().float().reshape(C, C)
        w_eff = pw2.reshape(C, C, 1, 1) * dw_w.detach().float().reshape(1, C, 3, 3)
        w.w_eff = pack_spike_conv_weight(w_eff, 1)
        w.bconst = (pw2 @ dw_b.detach().float() + pw_b.detach().float()).contiguous()
        w.w_wave = pack_lif_wave_weight(w_eff)
    return w


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------
_wave_ws_cache = {}


def _wave_ws(dev, nbytes: int) -> torch.Tensor:
    """Membrane scratch of the wavefront LIF kernel: one buffer per (device, stream), reused by every layer (the kernels of
    a stream run one after the other), so it stays resident in the L2."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    t = _wave_ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = _wave_ws_cache[key] = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    return t


def lif_ecs(x: Act, w: Optional[LifW], affine: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
            ecs_tau: float = 5.0, alpha: float = 0.75, beta: float = 0.25, save_mem: bool = False, allow_wave: bool = True):
    """mem_update.forward (models/common.py:252-283) -> bit-packed spikes; with save_mem also the membranes
    m_t [T,...] and ECS traces e_t [T-1,...] (the backward's recompute pass).  C % 64 != 0: the input current
    is zero-padded to the next multiple of 64 (``w`` must come from make_lif_w(..., Cp)); the result carries Cr = C."""
    Cr = x.C
    if Cr % 64:
        Cp = pad64(Cr)     # with save_mem the membranes / traces come back padded too (lif_ecs_bwd slices its results)
        if w is not None and w.dw_b.numel() != Cp:
            raise RuntimeError("lif_ecs: spread weights must be padded to %d channels" % Cp)
        x = Act(pad_channels(x.data, Cp), x.T)
        if affine is not None:
            affine = (pad_channels(affine[0], Cp), pad_channels(affine[1], Cp))
    T, N, H, W, C = x.T, x.N, x.H, x.W, x.C
    dev = x.data.device
    # algorithmic traffic of the neuron: every element-step of input current once (a T-broadcast input: one frame), 1 bit out
    flops["lif_fwd_elems"] += float(T) * N * H * W * Cr
    flops["lif_fwd_in_elems"] += float(x.Tp) * N * H * W * Cr
    bits = torch.empty(T, N, H, W, C // 32, device=dev, dtype=torch.int32)
    # The wavefront kernel is the inference path.  The BPTT chains (autograd.chain_fwd) keep, or their backward recomputes,
    # membranes and traces with the per-timestep pipeline, and both passes must see the SAME spikes (the folded-spread
    # arithmetic of the wavefront kernel differs from the pipeline's in the last bits): they pass allow_wave=False.
    if (not save_mem and allow_wave and w is not None and w.w_wave is not None and _state["lif_wave"] != "off"
            and not _state["lif_fused"]
            and (_cabi.lib().ecsy_lif_ecs_wave_supported(T, C, H, W) if _state["lif_wave"] == "all"
                 else _cabi.lib().ecsy_lif_ecs_wave_prefers(T, C, H, W))):
        sc, sh = affine if affine is not None else (None, None)
        flops["ecs_pw"] += 2.0 * (T - 1) * N * H * W * C * C
        ws = _wave_ws(dev, _cabi.lib().ecsy_lif_ecs_wave_ws_bytes(T, N, H, W, C))
        with _timed("lif_ecs", 1):
            _cabi.check(_cabi.lib().ecsy_lif_ecs_wave_fwd(
                _p(x.data), x.tstride, _p(sc), _p(sh), _p(w.w_wave), _p(w.bconst), _p(bits), T, N, H, W, C,
                float(thresh), float(decay), float(alpha), float(beta), float(1.0 - 1.0 / ecs_tau), _p(ws), ws.numel(),
                _st()), "lif_ecs_wave_fwd")
        return Spikes(bits, C, Cr)
    if (not save_mem and w is not None and w.w_eff is not None and _state["lif_fused"]
            and _cabi.lib().ecsy_lif_ecs_fused_supported(T, C)):
        sc, sh = affine if affine is not None else (None, None)
        flops["ecs_pw"] += 2.0 * (T - 1) * N * H * W * C * C
        with _timed("lif_ecs", 1):
            _cabi.check(_cabi.lib().ecsy_lif_ecs_fused_fwd(
                _p(x.data), x.tstride, _p(sc), _p(sh), _p(w.w_eff), _p(w.bconst), _p(bits), T, N, H, W, C,
                float(thresh), float(decay), float(alpha), float(beta), float(1.0 - 1.0 / ecs_tau), _st()),
                "lif_ecs_fused_fwd")
        return Spikes(bits, C, Cr)
    mem = torch.empty(T, N, H, W, C, device=dev, dtype=torch.float32) if save_mem else None
    ecs = torch.empty(max(T - 1, 1), N, H, W, C, device=dev, dtype=torch.float32) if save_mem else None
    L = _cabi.lib()
    splits = w.splits if w is not None else 1
    nws = L.ecsy_lif_ecs_ws_bytes(T, N, H, W, C, splits) if T > 1 else 0
    ws = torch.empty(max(nws, 16), device=dev, dtype=torch.uint8)
    sc, sh = affine if affine is not None else (None, None)
    if T > 1 and w is None:
        raise RuntimeError("lif_ecs: spread weights required for T > 1")
    flops["ecs_pw"] += 2.0 * (T - 1) * N * H * W * C * C
    with _timed("lif_ecs", 1 + 3 * (T - 1)):
        _cabi.check(L.ecsy_lif_ecs_fwd(_p(x.data), x.tstride, _p(sc), _p(sh),
                                   _p(w.dw_w) if w else None, _p(w.dw_b) if w else None,
                                   _p(w.pw) if w else None, _p(w.pw_b) if w else None, splits,
                                   _p(bits), _p(mem), _p(ecs), T, N, H, W, C, float(thresh), float(decay), float(alpha),
                                   float(beta), float(1.0 - 1.0 / ecs_tau), _p(ws), ws.numel(), _st()), "lif_ecs_fwd")
    sp = Spikes(bits, C, Cr)
    return (sp, mem, ecs) if save_mem else sp


def spread_dw(sp: Spikes, t: int, w: LifW, lo: bool = False, version: int = 0):
    """Depth-wise half of mem_update.spread (models/common.py:289-294) on the spikes of step t -> bf16 rows
    [N*H*W, C] (hi plane, and the residual plane when `lo`): the A operand of the point-wise spread GEMM."""
    N, H, W, C = sp.N, sp.H, sp.W, sp.C
    a_hi = torch.empty(N * H * W, C, device=sp.bits.device, dtype=torch.bfloat16)
    a_lo = torch.empty_like(a_hi) if lo else None
    with _timed("spread_dw", 1):
        _cabi.check(_cabi.lib().ecsy_spread_dw(_p(sp.bits[t]), _p(w.dw_w), _p(w.dw_b), _p(a_hi), _p(a_lo), N, H, W, C,
                                               int(version), _st()), "spread_dw")
    return (a_hi, a_lo) if lo else a_hi


def lif_ecs_bwd(gout: torch.Tensor, x: Act, w: LifW, pw_weight: torch.Tensor, affine=None, ecs_tau: float = 5.0,
                alpha: float = 0.75, beta: float = 0.25, saved=None):
    """Surrogate-gradient BPTT of lif_ecs.  gout: [T,N,H,W,C] dL/dspikes.  Re-runs the forward to recompute the
    membranes / ECS traces, then the reverse scan.  Returns (g_in [T,N,H,W,C] wrt the affine-applied input
    current, g_dw_w [C,1,3,3], g_dw_b [C], g_pw_w [C,C,1,1], g_pw_b [C])."""
    Cr = x.C
    if Cr % 64:
        # narrow layer (res*-ee.yaml front): run zero-padded to the 64-channel granule, slice the results
        Cp = pad64(Cr)
        x = Act(pad_channels(x.data, Cp), x.T)
        if affine is not None:
            affine = (pad_channels(affine[0], Cp), pad_channels(affine[1], Cp))
        gout = pad_channels(gout, Cp)
        pw_weight = torch.nn.functional.pad(pw_weight.detach().float(), (0, 0, 0, 0, 0, Cp - Cr, 0, Cp - Cr))
        gx, gdw, gdb, gpw, gpb = lif_ecs_bwd(gout, x, w, pw_weight, affine, ecs_tau, alpha, beta, saved)
        return (gx[..., :Cr].contiguous(), gdw[:Cr].contiguous(), gdb[:Cr].contiguous(),
                gpw[:Cr, :Cr].contiguous(), gpb[:Cr].contiguous())
    T, N, H, W, C = x.T, x.N, x.H, x.W, x.C
    dev = x.data.device
    flops["lif_elems"] += float(T) * N * H * W * C
    # `saved` = (spikes, membranes, traces) kept by a forward run with save_mem (set_lif_store); otherwise recompute
    sp, mem, ecs = saved if saved is not None else lif_ecs(x, w, affine, ecs_tau, alpha, beta, save_mem=True)
    gout = gout.contiguous()
    gx = torch.empty(T, N, H, W, C, device=dev, dtype=torch.float32)
    g_dw_w = torch.zeros(9, C, device=dev, dtype=torch.float32)
    g_dw_b = torch.zeros(C, device=dev, dtype=torch.float32)
    g_pw_w = torch.zeros(C, C, device=dev, dtype=torch.float32)
    g_pw_b = torch.zeros(C, device=dev, dtype=torch.float32)
    pwT = pack_conv_weight(pw_weight.detach().reshape(C, C).t().contiguous().reshape(C, C, 1, 1), w.splits)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_lif_ecs_bwd_ws_bytes(T, N, H, W, C, w.splits), device=dev, dtype=torch.uint8)
    with _timed("lif_ecs_bwd", 1 + 6 * (T - 1)):
        _cabi.check(L.ecsy_lif_ecs_bwd(_p(gout), _p(sp.bits), _p(mem), _p(ecs), _p(w.dw_w), _p(w.dw_b), _p(pwT), w.splits,
                                       _p(gx), _p(g_dw_w), _p(g_dw_b), _p(g_pw_w), _p(g_pw_b), T, N, H, W, C,
                                       float(thresh), float(lens), float(decay), float(alpha), float(beta),
                                       float(1.0 - 1.0 / ecs_tau), _p(ws), ws.numel(), _st()), "lif_ecs_bwd")
    return gx, g_dw_w.t().reshape(C, 1, 3, 3).contiguous(), g_dw_b, g_pw_w.reshape(C, C, 1, 1), g_pw_b


def pack_dgrad_weight(weight: torch.Tensor, splits: int) -> torch.Tensor:
    """Forward-form packing of the flipped, transposed weight: W'[ci][co][ky][kx] = W[co][ci][k-1-ky][k-1-kx]."""
    wt = weight.detach().float().flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return pack_conv_weight(wt, splits)


def conv_dgrad(gy: torch.Tensor, wT_packed: torch.Tensor, splits: int, H: int, W: int, Cin: int, k: int, stride: int,
               pad: int) -> torch.Tensor:
    """gy: [T,N,Ho,Wo,Co] fp32 NHWC -> gx [T,N,H,W,Cin] (input gradient of Snn_Conv2d)."""
    T, N, Ho, Wo, Co = gy.shape
    gy = gy.contiguous()
    gx = torch.empty(T, N, H, W, Cin, device=gy.device, dtype=torch.float32)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_conv_dgrad_ws_bytes(T * N, H, W, Co, k, stride, pad, splits), device=gy.device, dtype=torch.uint8)
    flops["conv_bwd"] = flops.get("conv_bwd", 0.0) + 2.0 * T * N * H * W * Cin * Co * k * k
    with _timed("conv_dgrad", 2):
        _cabi.check(L.ecsy_conv_dgrad(_p(gy), _p(wT_packed), splits, _p(gx), T * N, H, W, Cin, Co, k, stride, pad,
                                      _p(ws), ws.numel(), _st()), "conv_dgrad")
    return gx


def spike_conv_wgrad(gy: torch.Tensor, s: Spikes, k: int, stride: int, pad: int) -> torch.Tensor:
    """gy: [T,N,Ho,Wo,Co] fp32, s: the conv's input spikes -> dW [Co, Ci, k, k]."""
    T, N, Ho, Wo, Co = gy.shape
    gy = gy.contiguous()
    Ci = s.C
    dw = torch.zeros(Co, k * k * Ci, device=gy.device, dtype=torch.float32)
    L = _cabi.lib()
    splits = get_splits()
    ws = torch.empty(L.ecsy_spike_conv_wgrad_ws_bytes(T * N, Ho, Wo, Co, splits), device=gy.device, dtype=torch.uint8)
    flops["conv_bwd"] = flops.get("conv_bwd", 0.0) + 2.0 * T * N * Ho * Wo * Ci * Co * k * k
    with _timed("conv_wgrad", 2):
        _cabi.check(L.ecsy_spike_conv_wgrad(_p(gy), _p(s.bits), _p(dw), T * N, s.H, s.W, Ci, Co, k, stride, pad, splits,
                                            _p(ws), ws.numel(), _st()), "spike_conv_wgrad")
    return dw.reshape(Co, k, k, Ci).permute(0, 3, 1, 2).contiguous()


def spike_conv_bwd(g: torch.Tensor, y: Optional[torch.Tensor], coef, s: Spikes, wT_packed: torch.Tensor, k: int, stride: int,
                   pad: int, Cin: int):
    """Backward of Snn_Conv2d (+ tdBN) on spikes in one call.  g: [T,N,Ho,Wo,Co] gradient w.r.t. the normalised
    output with coef = (A, B, C) per-channel vectors and y the raw conv output (g_y = A*g + B*y + C), or the plain
    output gradient with coef = None.  -> (g_s [T,N,H,W,Cin], dW [Co, Cin, k, k])."""
    T, N, Ho, Wo, Co = g.shape
    g = g.contiguous()
    H, W = s.H, s.W
    dev = g.device
    splits = get_splits()
    gx = torch.empty(T, N, H, W, Cin, device=dev, dtype=torch.float32)
    dw = torch.zeros(Co, k * k * Cin, device=dev, dtype=torch.float32)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_spike_conv_bwd_ws_bytes(T * N, H, W, Co, k, stride, pad, splits), device=dev, dtype=torch.uint8)
    A, B, Cv = coef if coef is not None else (None, None, None)
    # same accounting as conv_dgrad + spike_conv_wgrad (the dgrad of a strided conv runs over the zero-inserted gradient)
    flops["conv_bwd"] = flops.get("conv_bwd", 0.0) + 2.0 * T * N * (H * W + Ho * Wo) * Cin * Co * k * k
    with _timed("conv_bwd", 3 if stride == 1 else 4):
        _cabi.check(L.ecsy_spike_conv_bwd(_p(g), _p(y), _p(A), _p(B), _p(Cv), _p(s.bits), _p(wT_packed), splits, _p(gx), _p(dw),
                                          T * N, H, W, Cin, Co, k, stride, pad, _p(ws), ws.numel(), _st()), "spike_conv_bwd")
    return gx, dw.reshape(Co, k, k, Cin).permute(0, 3, 1, 2).contiguous()


def real_conv_wgrad(gy: torch.Tensor, x: Act, k: int, stride: int, pad: int) -> torch.Tensor:
    """gy: [Tp,N,Ho,Wo,Co] fp32 (already summed over T for a T-broadcast input), x: the conv's real input."""
    Tp, N, Ho, Wo, Co = gy.shape
    assert Tp == x.Tp
    gy = gy.contiguous()
    Ci = x.C
    Kpad = (k * k * Ci + 63) // 64 * 64
    dw = torch.zeros(Co, Kpad, device=gy.device, dtype=torch.float32)
    L = _cabi.lib()
    splits = get_splits()
    ws = torch.empty(L.ecsy_real_conv_wgrad_ws_bytes(Tp * N, x.H, x.W, Ci, Co, k, stride, pad, splits), device=gy.device,
                     dtype=torch.uint8)
    with _timed("conv_wgrad", 3):
        _cabi.check(L.ecsy_real_conv_wgrad(_p(gy), _p(x.data), Tp * N, _p(dw), Tp * N, x.H, x.W, Ci, Co, k, stride, pad,
                                           splits, _p(ws), ws.numel(), _st()), "real_conv_wgrad")
    return dw[:, :k * k * Ci].reshape(Co, k, k, Ci).permute(0, 3, 1, 2).contiguous()


def colsum2(g: torch.Tensor, x: torch.Tensor, C: int):
    """sum_r g[r,c], sum_r g[r,c]*x[r mod x_rows, c] for [rows, C]-flattened NHWC tensors."""
    rows, xr = g.numel() // C, x.numel() // C
    sg = torch.empty(C, device=g.device, dtype=torch.float32)
    sgx = torch.empty(C, device=g.device, dtype=torch.float32)
    ws = torch.empty(16 * C + 512, device=g.device, dtype=torch.uint8)
    with _timed("colsum2", 3):
        _cabi.check(_cabi.lib().ecsy_colsum2(_p(g), _p(x), rows, xr, C, _p(sg), _p(sgx), _p(ws), ws.numel(), _st()),
                    "colsum2")
    return sg, sgx


def lif_silu(x: Act, w: Optional[LifW], affine=None, ecs_tau: float = 5.0, alpha: float = 0.75, beta: float = 0.25,
             inplace: bool = True, save: bool = False):
    """mem_update(act=True).forward (models/common.py:252-283): real-valued silu(mem_t) outputs; with `save`
    also the raw membranes [T,...] and ECS traces [T-1,...] (backward recompute)."""
    T, N, H, W, C = x.T, x.N, x.H, x.W, x.C
    dev = x.data.device
    out = torch.empty(T, N, H, W, C, device=dev, dtype=torch.float32)
    mem = torch.empty(T, N, H, W, C, device=dev, dtype=torch.float32) if save else None
    ecs = torch.empty(max(T - 1, 1), N, H, W, C, device=dev, dtype=torch.float32) if save else None
    L = _cabi.lib()
    splits = w.splits if w is not None else 1
    nws = L.ecsy_lif_silu_ws_bytes(T, N, H, W, C, splits) if T > 1 else 0
    ws = torch.empty(max(nws, 16), device=dev, dtype=torch.uint8)
    sc, sh = affine if affine is not None else (None, None)
    if T > 1 and w is None:
        raise RuntimeError("lif_silu: spread weights required for T > 1")
    flops["ecs_pw"] += 2.0 * (T - 1) * N * H * W * C * C
    with _timed("lif_silu", 1 + 3 * (T - 1)):
        _cabi.check(L.ecsy_lif_silu_fwd(_p(x.data), x.tstride, _p(sc), _p(sh),
                                        _p(w.dw_w) if w else None, _p(w.dw_b) if w else None,
                                        _p(w.pw) if w else None, _p(w.pw_b) if w else None, splits, _p(out),
                                        _p(mem), _p(ecs), 1 if inplace else 0, T, N, H, W, C, float(decay),
                                        float(alpha), float(beta), float(1.0 - 1.0 / ecs_tau), _p(ws), ws.numel(),
                                        _st()), "lif_silu_fwd")
    return (Act(out, T), mem, ecs) if save else Act(out, T)


def lif_silu_bwd(gout: torch.Tensor, x: Act, w: LifW, pw_weight: torch.Tensor, affine=None, ecs_tau: float = 5.0,
                 alpha: float = 0.75, beta: float = 0.25):
    """BPTT of the in-place SiLU neuron (class Conv inside a model).  Same returns as lif_ecs_bwd."""
    T, N, H, W, C = x.T, x.N, x.H, x.W, x.C
    dev = x.data.device
    out, mem, ecs = lif_silu(x, w, affine, ecs_tau, alpha, beta, inplace=True, save=True)
    gout = gout.contiguous()
    gx = torch.empty(T, N, H, W, C, device=dev, dtype=torch.float32)
    g_dw_w = torch.zeros(9, C, device=dev, dtype=torch.float32)
    g_dw_b = torch.zeros(C, device=dev, dtype=torch.float32)
    g_pw_w = torch.zeros(C, C, device=dev, dtype=torch.float32)
    g_pw_b = torch.zeros(C, device=dev, dtype=torch.float32)
    pwT = pack_conv_weight(pw_weight.detach().reshape(C, C).t().contiguous().reshape(C, C, 1, 1), w.splits)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_lif_ecs_bwd_ws_bytes(T, N, H, W, C, w.splits), device=dev, dtype=torch.uint8)
    with _timed("lif_silu_bwd", 1 + 7 * (T - 1)):
        _cabi.check(L.ecsy_lif_silu_bwd(_p(gout), _p(out.data), _p(mem), _p(ecs), _p(w.dw_w), _p(w.dw_b), _p(pwT),
                                        w.splits, _p(gx), _p(g_dw_w), _p(g_dw_b), _p(g_pw_w), _p(g_pw_b), T, N, H, W, C,
                                        float(decay), float(alpha), float(beta), float(1.0 - 1.0 / ecs_tau), _p(ws),
                                        ws.numel(), _st()), "lif_silu_bwd")
    return gx, g_dw_w.t().reshape(C, 1, 3, 3).contiguous(), g_dw_b, g_pw_w.reshape(C, C, 1, 1), g_pw_b


def spike_conv(s: Spikes, w: ConvW, scale=None, shift=None, residual: Optional[Act] = None) -> Act:
    """Snn_Conv2d on spikes (models/common.py:609-624) with optional folded tdBN and shortcut add."""
    T, N, H, W = s.T, s.N, s.H, s.W
    if (w.groups != 1 and not w.dense_groups) or w.packed is None:
        raise RuntimeError("spike_conv: grouped / unpacked weights go through real_conv on unpacked spikes")
    Ho = (H + 2 * w.pad - w.k) // w.stride + 1
    Wo = (W + 2 * w.pad - w.k) // w.stride + 1
    out = torch.empty(T, N, Ho, Wo, w.co, device=s.bits.device, dtype=torch.float32)
    if residual is not None:
        assert (residual.N, residual.H, residual.W, residual.C) == (N, Ho, Wo, w.co), "residual shape"
    flops["spike_conv"] += 2.0 * T * N * Ho * Wo * w.co * s.C * w.k * w.k
    ts = w.packed_ts is not None and _use_ts(s.C, w.co, w.splits)
    pair = (not ts) and w.packed_ts is not None and _state["conv_ts"] != "off"   # smem operand, tensor-memory weight layout
    L = _cabi.lib()
    fn = L.ecsy_spike_conv_ts_fwd if ts else (L.ecsy_spike_conv_pair_fwd if pair else L.ecsy_spike_conv_fwd)
    with _timed("spike_conv", 1):
        _cabi.check(fn(
            _p(s.bits), _p(w.packed_ts if (ts or pair) else w.packed), w.splits, _p(out), _p(scale), _p(shift),
            _p(residual.data) if residual is not None else None, residual.src_imgs if residual is not None else 0,
            T * N, H, W, s.C, w.co, w.k, w.stride, w.pad, _st()), "spike_conv_fwd")
    return Act(out, T)


def real_conv(x: Act, w: ConvW, scale=None, shift=None, bias_mul: float = 1.0) -> Act:
    """Snn_Conv2d on a real-valued input.  A T-broadcast input is convolved once."""
    Tp, N, H, W = x.Tp, x.N, x.H, x.W
    Ho = (H + 2 * w.pad - w.k) // w.stride + 1
    Wo = (W + 2 * w.pad - w.k) // w.stride + 1
    out = torch.empty(Tp, N, Ho, Wo, w.co, device=x.data.device, dtype=torch.float32)
    L = _cabi.lib()
    if w.stem is not None and w.splits == 1 and _state["stem_kernel"]:
        # the image stem (3 -> 64, 7x7 / 2) in fast precision: dedicated tcgen05 kernel on the NHWC4 bf16 image
        ws = torch.empty(L.ecsy_stem_conv_ws_bytes(Tp * N, H, W), device=x.data.device, dtype=torch.uint8)
        flops["real_conv"] += 2.0 * Tp * N * Ho * Wo * w.co * w.ci * w.k * w.k
        with _timed("real_conv", 2):
            _cabi.check(L.ecsy_stem_conv(_p(x.data), Tp * N, H, W, w.ci, _p(w.stem), _p(out), _p(scale), _p(shift), w.co, w.k,
                                         w.stride, w.pad, _p(ws), ws.numel(), _st()), "stem_conv")
        return Act(out, x.T)
    use_umma = w.packed is not None and w.groups == 1 and w.co % 64 == 0 and w.bias is None
    if w.splits == 2 and w.ci < 64 and w.simt is not None:
        use_umma = False  # parity mode: the tiny-K stem runs in plain fp32 (the 3-term bf16 split leaves ~6e-6)
    nws = L.ecsy_real_conv_ws_bytes(Tp * N, H, W, w.ci, w.co, w.k, w.stride, w.pad, w.groups, w.splits) if use_umma else 0
    ws = torch.empty(max(nws, 16), device=x.data.device, dtype=torch.uint8)
    if not use_umma and w.simt is None:
        raise RuntimeError("real_conv: SIMT weight layout missing")
    flops["real_conv"] += 2.0 * Tp * N * Ho * Wo * w.co * (w.ci // w.groups) * w.k * w.k
    with _timed("real_conv", 2 if use_umma else 1):
        _cabi.check(L.ecsy_real_conv_fwd(_p(x.data), Tp * N, _p(w.packed) if use_umma else None, _p(w.simt), w.splits,
                                         _p(w.bias), float(bias_mul), _p(scale), _p(shift), _p(out), Tp * N, H, W, w.ci,
                                         w.co, w.k, w.stride, w.pad, w.groups, _p(ws), ws.numel(), _st()), "real_conv_fwd")
    return Act(out, x.T)


def bn_stats(y: Act) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel mean / biased variance over (T, N, H, W)."""
    rows = y.Tp * y.N * y.H * y.W
    C = y.C
    dev = y.data.device
    flops["tdbn_elems"] += float(rows) * C
    mean = torch.empty(C, device=dev, dtype=torch.float32)
    var = torch.empty(C, device=dev, dtype=torch.float32)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_tdbn_stats_ws_bytes(rows, C), device=dev, dtype=torch.uint8)
    with _timed("tdbn_stats", 2):
        _cabi.check(L.ecsy_tdbn_stats(_p(y.data), rows, C, _p(mean), _p(var), _p(ws), ws.numel(), _st()), "tdbn_stats")
    return mean, var


def affine_add(a: Act, sa=None, ba=None, b: Optional[Act] = None, sb=None, bb=None) -> Act:
    T = a.T
    Tp = T if (a.Tp == T or (b is not None and b.Tp == T)) else 1
    out = torch.empty(Tp, a.N, a.H, a.W, a.C, device=a.data.device, dtype=torch.float32)
    if b is not None:
        assert (b.N, b.H, b.W, b.C) == (a.N, a.H, a.W, a.C), "affine_add shapes"
    with _timed("affine_add", 1):
        _cabi.check(_cabi.lib().ecsy_affine_add(_p(a.data), a.src_imgs, _p(sa), _p(ba),
                                                _p(b.data) if b is not None else None, b.src_imgs if b is not None else 0,
                                                _p(sb), _p(bb), _p(out), Tp * a.N, a.H * a.W, a.C, _st()), "affine_add")
    return Act(out, T)


def resample_into(x: Act, out: torch.Tensor, coff: int, pool: int = 1, up: int = 1, scale=None, shift=None) -> None:
    """max-pool / nearest-up / copy `x` into channels [coff, coff + x.C) of out [Tp, N, Ho, Wo, Ctot]."""
    imgs = out.shape[0] * out.shape[1]
    with _timed("resample", 1):
        _cabi.check(_cabi.lib().ecsy_resample(_p(x.data), x.src_imgs, _p(scale), _p(shift), _p(out), imgs, x.H, x.W, x.C,
                                              out.shape[4], coff, pool, up, _st()), "resample")


def maxpool(x: Act, s: int) -> Act:
    if s == 1:
        return x
    out = torch.empty(x.Tp, x.N, x.H // s, x.W // s, x.C, device=x.data.device, dtype=torch.float32)
    resample_into(x, out, 0, pool=s)
    return Act(out, x.T)


def upsample(x: Act, s: int) -> Act:
    out = torch.empty(x.Tp, x.N, x.H * s, x.W * s, x.C, device=x.data.device, dtype=torch.float32)
    resample_into(x, out, 0, up=s)
    return Act(out, x.T)


def concat_channels(xs, pool: int = 1) -> Act:
    T = xs[0].T
    Tp = T if any(a.Tp == T for a in xs) else 1
    Ct = sum(a.C for a in xs)
    a0 = xs[0]
    out = torch.empty(Tp, a0.N, a0.H // pool, a0.W // pool, Ct, device=a0.data.device, dtype=torch.float32)
    off = 0
    for a in xs:
        resample_into(a, out, off, pool=pool)
        off += a.C
    return Act(out, T)


def maxpool_bwd(x: Act, g_pooled: torch.Tensor, s: int, gcoff: int = 0) -> torch.Tensor:
    """x: the pool input; g_pooled: [T,N,Ho,Wo,gC] (channels [gcoff, gcoff+x.C) are this input's) -> gx [T,N,H,W,C]."""
    T, N, Ho, Wo, gC = g_pooled.shape
    gx = torch.empty(T, N, Ho * s, Wo * s, x.C, device=g_pooled.device, dtype=torch.float32)
    with _timed("resample_bwd", 1):
        _cabi.check(_cabi.lib().ecsy_maxpool_bwd(_p(x.data), x.src_imgs, _p(g_pooled), _p(gx), T * N, Ho, Wo, x.C, gC,
                                                 gcoff, s, _st()), "maxpool_bwd")
    return gx


def sumpool_slice(g: torch.Tensor, C: int, coff: int, s: int) -> torch.Tensor:
    """[T,N,H,W,inC] -> [T,N,H/s,W/s,C]: block sums of channels [coff, coff+C) (Sample / Concat backward)."""
    T, N, H, W, inC = g.shape
    out = torch.empty(T, N, H // s, W // s, C, device=g.device, dtype=torch.float32)
    with _timed("resample_bwd", 1):
        _cabi.check(_cabi.lib().ecsy_sumpool_slice(_p(g), _p(out), T * N, H // s, W // s, C, inC, coff, s, _st()),
                    "sumpool_slice")
    return out


def tsum(x: Act, w: Optional[torch.Tensor], div: float) -> torch.Tensor:
    """[T, N, H, W, C] -> [N, H, W, C]: (sum_t w[t] x[t]) / div."""
    x = x.full()
    out = torch.empty(x.N, x.H, x.W, x.C, device=x.data.device, dtype=torch.float32)
    with _timed("tsum", 1):
        _cabi.check(_cabi.lib().ecsy_tsum(_p(x.data), _p(w), float(div), _p(out), x.T, out.numel(), _st()), "tsum")
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """[imgs, H, W, C] -> [imgs, C, H, W] contiguous."""
    n, H, W, C = x.shape
    out = torch.empty(n, C, H, W, device=x.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().ecsy_nhwc_to_nchw_f32(_p(x), _p(out), n, C, H, W, _st()), "nhwc_to_nchw")
    return out


def detect_decode(y: torch.Tensor, na: int, no: int, anchors: torch.Tensor, stride_px: float, z: Optional[torch.Tensor],
                  row_off: int) -> torch.Tensor:
    """y: [N, H, W, na*no] -> raw [N, na, H, W, no]; fills z rows if given."""
    N, H, W, _ = y.shape
    raw = torch.empty(N, na, H, W, no, device=y.device, dtype=torch.float32)
    with _timed("detect_decode", 1):
        _cabi.check(_cabi.lib().ecsy_detect_decode(_p(y), _p(raw), _p(z), _p(anchors), float(stride_px), N, H, W, na, no,
                                                   z.shape[1] if z is not None else 0, row_off, _st()), "detect_decode")
    return raw


def ddetect_decode(box: torch.Tensor, cls: torch.Tensor, stride_px: float, y: Optional[torch.Tensor], a_off: int):
    N, H, W, _ = box.shape
    nc = cls.shape[3]
    xs = torch.empty(N, 64 + nc, H, W, device=box.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().ecsy_ddetect_decode(_p(box), _p(cls), _p(xs), _p(y), float(stride_px), N, H, W, nc,
                                                y.shape[2] if y is not None else 0, a_off, _st()), "ddetect_decode")
    return xs
