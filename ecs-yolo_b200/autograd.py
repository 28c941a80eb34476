"""Manual backward of the fused train-mode chains, exposed as torch.autograd.Functions per block.

The reference gets its backward from PyTorch autograd through every aten op of the time loop
(SURVEY section 8 row a11).  Here each block's backward is an explicit chain of C-ABI calls:

    tdBN backward   : ecsy_colsum2 (sum g, sum g*y) + ecsy_affine_add (g_y = A*g + B*y + C per channel)
    conv backward   : ecsy_spike_conv_wgrad / ecsy_real_conv_wgrad, ecsy_conv_dgrad          (tcgen05)
    neuron backward : ecsy_lif_ecs_bwd -- surrogate-gradient BPTT, forward recomputed, not stored
    resampling      : ecsy_maxpool_bwd, ecsy_sumpool_slice

torch.autograd only connects the blocks (and runs the tiny Detect head, < 0.1 % of the FLOPs).
Gradients between chains are "w.r.t. the normalised tensor", so a chain's tdBN backward consumes
exactly what the next chain's neuron backward produces.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import functional as F_
from .functional import Act


def _params(mod) -> list:
    return [p for p in mod.parameters()]


def _grad_tuple(mod, grads: Dict[int, torch.Tensor]):
    out = []
    for p in _params(mod):
        g = grads.get(id(p))
        if g is not None and g.shape != p.shape:
            g = g.reshape(p.shape)
        if g is not None and g.is_contiguous() and g.stride() != p.stride() and p.is_contiguous():
            g = g.as_strided(p.shape, p.stride())   # size-1 dims: give DDP's bucket views the canonical strides
        out.append(g)
    return tuple(out)


def _grad_tuple_params(params, grads: Dict[int, torch.Tensor]):
    out = []
    for p in params:
        g = grads.get(id(p))
        if g is not None and g.shape != p.shape:
            g = g.reshape(p.shape)
        if g is not None and g.is_contiguous() and g.stride() != p.stride() and p.is_contiguous():
            g = g.as_strided(p.shape, p.stride())
        out.append(g)
    return tuple(out)


def _acc(grads, p, g):
    grads[id(p)] = g if id(p) not in grads else grads[id(p)] + g


# ------------------------------------------------------------------------------------------------
# tdBN (train mode)
# ------------------------------------------------------------------------------------------------
def bn_train_fwd(bn_mod, y: Act):
    """-> (scale, shift, mean, rstd); updates running statistics like _tdbn.scale_shift.  One launch after the batch
    statistics (ecsy_tdbn_finish) instead of ~11 element-wise torch kernels on [C] vectors."""
    bn = bn_mod.bn
    mean, var = F_.bn_stats(y)
    C = y.C
    n = float(y.T * y.N * y.H * y.W)
    if isinstance(bn, torch.nn.SyncBatchNorm):      # the reference's --sync-bn (train.py:359-360): statistics of all ranks
        from . import dist as D
        mean, var, world = D.sync_bn_stats(mean, var)
        n *= world
    track = bool(bn.track_running_stats) and bn.running_mean is not None
    if (track and bn.momentum is None) or bn.weight is None:
        return _bn_train_fwd_torch(bn_mod, y, mean, var, n)     # cumulative average: needs the step count on the host
    scale = torch.empty(C, device=mean.device, dtype=torch.float32)
    shift, rstd = torch.empty_like(scale), torch.empty_like(scale)
    with torch.no_grad():
        F_._cabi.check(F_._cabi.lib().ecsy_tdbn_finish(
            F_._p(mean), F_._p(var), F_._p(bn.weight), F_._p(bn.bias),
            F_._p(bn.running_mean) if track else None, F_._p(bn.running_var) if track else None,
            F_._p(bn.num_batches_tracked) if track else None, float(bn.momentum or 0.0), n / max(n - 1.0, 1.0), float(bn.eps),
            int(bn_mod.stat_updates) if track else 0, F_._p(scale), F_._p(shift), F_._p(rstd), C, F_._st()), "tdbn_finish")
        if track:   # the kernel wrote the buffers through raw pointers
            torch._C._increment_version([bn.running_mean, bn.running_var, bn.num_batches_tracked])
    return scale, shift, mean, rstd


def _bn_train_fwd_torch(bn_mod, y: Act, mean, var, n):
    bn = bn_mod.bn
    with torch.no_grad():
        if bn.track_running_stats:
            for _ in range(bn_mod.stat_updates):
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1.0 - m).add_(mean, alpha=m)
                bn.running_var.mul_(1.0 - m).add_(var, alpha=m * n / max(n - 1.0, 1.0))
        rstd = torch.rsqrt(var + bn.eps)
        scale = (bn.weight * rstd).contiguous()
        shift = (bn.bias - mean * scale).contiguous()
    return scale, shift, mean, rstd


def bn_train_bwd_coeffs(bn_mod, y: Act, mean, rstd, g_yn: torch.Tensor, grads):
    """Batch sums of the tdBN backward -> per-channel (A, B, C) with g_y = A*g_yn + B*y + C; accumulates the
    gradients of the tdBN weight / bias.  colsum2 + ONE launch for the [C]-vector algebra (ecsy_tdbn_bwd_coef)."""
    bn = bn_mod.bn
    C = y.C
    n = float(y.T * y.N * y.H * y.W)
    tfac = float(y.T) / float(y.Tp)
    sg, sgy = F_.colsum2(g_yn, y.data, C)
    sg_l, sgy_l, world = sg, sgy, 1
    if isinstance(bn, torch.nn.SyncBatchNorm):      # input gradient from the sums over all ranks, parameter gradients local
        from . import dist as D
        sg, sgy, world = D.sync_bn_sums(sg, sgy)
        n *= world
    A = torch.empty(C, device=sg.device, dtype=torch.float32)
    B, Cc, sgx = torch.empty_like(A), torch.empty_like(A), torch.empty_like(A)
    with torch.no_grad():
        F_._cabi.check(F_._cabi.lib().ecsy_tdbn_bwd_coef(F_._p(sg), F_._p(sgy), F_._p(mean), F_._p(rstd), F_._p(bn.weight), n, tfac,
                                                         F_._p(A), F_._p(B), F_._p(Cc), F_._p(sgx), C, F_._st()), "tdbn_bwd_coef")
        if world > 1:
            sgx = rstd * (sgy_l - mean * sg_l)
    _acc(grads, bn.weight, sgx)
    _acc(grads, bn.bias, sg_l)
    return A, B, Cc


def bn_train_bwd(bn_mod, y: Act, mean, rstd, g_yn: torch.Tensor, grads):
    """g_yn: [Tp,N,H,W,C] gradient w.r.t. the normalised output (already summed over T when y is a
    T-broadcast tensor).  Returns g_y with the same shape."""
    A, B, Cc = bn_train_bwd_coeffs(bn_mod, y, mean, rstd, g_yn, grads)
    zeros = torch.zeros_like(B)
    return F_.affine_add(Act(g_yn, g_yn.shape[0]), A, Cc, Act(y.data, y.Tp), B, zeros).data


# ------------------------------------------------------------------------------------------------
# LIF -> spike conv -> tdBN
# ------------------------------------------------------------------------------------------------
class _Saved:
    __slots__ = ("x", "aff", "sp", "y", "mean", "rstd", "scale", "shift", "state")


def chain_fwd(lif, conv, bn, x: Act, aff):
    sv = _Saved()
    sv.x, sv.aff = x, aff
    if F_.lif_store_ok(x):
        # keep membranes + traces for the backward (memory for time on a 180 GB part): no forward recompute there
        if lif.spread is None:
            lif._init_spread(x.C, x.data.device)
        sv.sp, mem, ecs = F_.lif_ecs(x, lif._weights(), aff, lif.ecs_tau, lif.alpha, lif.beta, save_mem=True)
        sv.state = (sv.sp, mem, ecs)
    else:
        # the backward recomputes membranes / traces with the per-timestep pipeline: the forward must produce ITS spikes
        if lif.spread is None:
            lif._init_spread(x.C, x.data.device)
        sv.sp = F_.lif_ecs(x, lif._weights(), aff, lif.ecs_tau, lif.alpha, lif.beta, allow_wave=False)
        sv.state = None
    sv.y = conv.conv_spikes(sv.sp)
    sv.scale, sv.shift, sv.mean, sv.rstd = bn_train_fwd(bn, sv.y)
    return sv


def conv_spikes_bwd(conv, sp, g_y: torch.Tensor, grads):
    """-> g_s [T,N,H,W,Ci]; accumulates dW."""
    from .common import _cached
    k, s, p = conv._geom()
    co, ci = conv.out_channels, conv.in_channels
    if (sp.C != ci or co % 64) and conv.groups == 1:
        # narrow layer (res*-ee.yaml front) on the padded 64-channel granule: zero-padded output gradient and weights,
        # the real block of the weight gradient; the input gradient keeps the padded width of the spike tensor
        cop = F_.pad64(co)
        g_yp = F_.pad_channels(g_y, cop)
        _acc(grads, conv.weight, F_.spike_conv_wgrad(g_yp, sp, k, s, p)[:co, :ci].contiguous())
        splits = F_.get_splits()

        def build_p():
            w = torch.nn.functional.pad(conv.weight.detach().float(), (0, 0, 0, 0, 0, sp.C - ci, 0, cop - co))
            return F_.pack_dgrad_weight(w, splits)
        wT = _cached(conv, "dgradw_pad", (conv.weight,), build_p)
        g_s = F_.conv_dgrad(g_yp, wT, splits, sp.H, sp.W, sp.C, k, s, p)
        return g_s[..., :ci].contiguous() if sp.C != ci else g_s
    dw = F_.spike_conv_wgrad(g_y, sp, k, s, p)
    if conv.groups > 1:  # block-diagonal part of the dense gradient
        cog, cig = conv.out_channels // conv.groups, conv.in_channels // conv.groups
        dw = torch.cat([dw[g * cog:(g + 1) * cog, g * cig:(g + 1) * cig] for g in range(conv.groups)], 0)
    _acc(grads, conv.weight, dw)
    splits = F_.get_splits()

    def build():
        w = conv.weight
        if conv.groups > 1:
            w = F_.densify_grouped(w, conv.groups)
        return F_.pack_dgrad_weight(w, splits)
    wT = _cached(conv, "dgradw", (conv.weight,), build)
    return F_.conv_dgrad(g_y, wT, splits, sp.H, sp.W, conv.in_channels, k, s, p)


def chain_bwd(lif, conv, bn, sv: _Saved, g_yn: torch.Tensor, grads):
    """g_yn: gradient w.r.t. the chain's normalised output -> gradient w.r.t. its (normalised) input."""
    if conv.groups == 1 and conv.out_channels % 64 == 0 and sv.sp.C == conv.in_channels and sv.y.Tp == sv.y.T:
        # one call: the output gradient (tdBN backward folded in) is formed once as bf16 planes for wgrad and dgrad
        from .common import _cached
        coef = bn_train_bwd_coeffs(bn, sv.y, sv.mean, sv.rstd, g_yn, grads)
        k, s_, p_ = conv._geom()
        splits = F_.get_splits()
        wT = _cached(conv, "dgradw", (conv.weight,), lambda: F_.pack_dgrad_weight(conv.weight, splits))
        g_s, dw = F_.spike_conv_bwd(g_yn, sv.y.data, coef, sv.sp, wT, k, s_, p_, conv.in_channels)
        _acc(grads, conv.weight, dw)
    else:
        g_y = bn_train_bwd(bn, sv.y, sv.mean, sv.rstd, g_yn, grads)
        g_s = conv_spikes_bwd(conv, sv.sp, g_y, grads)
    g_x, gdw, gdb, gpw, gpb = F_.lif_ecs_bwd(g_s, sv.x, lif._weights(), lif.spread[1].weight, sv.aff, lif.ecs_tau,
                                             lif.alpha, lif.beta, saved=sv.state)
    sv.state = None
    _acc(grads, lif.spread[0].weight, gdw)
    _acc(grads, lif.spread[0].bias, gdb)
    _acc(grads, lif.spread[1].weight, gpw)
    _acc(grads, lif.spread[1].bias, gpb)
    return g_x


def _to_nhwc(g: torch.Tensor) -> torch.Tensor:
    """Incoming reference-shaped gradient [T,N,C,H,W] (any strides) -> contiguous NHWC data [T,N,H,W,C]."""
    a = Act.from_ref(g)
    return a.full().data if a.Tp != a.T else a.data


def _add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return F_.affine_add(Act(a, a.shape[0]), None, None, Act(b, b.shape[0])).data


class BasicBlockFn(torch.autograd.Function):
    """BasicBlock_1 / BasicBlock_2 in training mode (models/common.py:1049-1079, 1182-1219)."""

    @staticmethod
    def forward(ctx, block, x_ref, *params):
        a = Act.from_ref(x_ref)
        l1, c1, b1, l2, c2, b2 = block.residual_function
        s1 = chain_fwd(l1, c1, b1, a, None)
        s2 = chain_fwd(l2, c2, b2, s1.y, (s1.scale, s1.shift))
        ctx.block, ctx.a, ctx.s1, ctx.s2 = block, a, s1, s2
        if len(block.shortcut) == 0:
            ctx.s3 = None
            out = F_.affine_add(s2.y, s2.scale, s2.shift, a, None, None)
        else:
            pool, l3, c3, b3 = block.shortcut
            ctx.pool = pool.stride[1]
            ctx.xp = F_.maxpool(a, ctx.pool)
            s3 = chain_fwd(l3, c3, b3, ctx.xp, None)
            ctx.s3 = s3
            out = F_.affine_add(s2.y, s2.scale, s2.shift, s3.y, s3.scale, s3.shift)
        return out.to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        block = ctx.block
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        l1, c1, b1, l2, c2, b2 = block.residual_function
        g_y1n = chain_bwd(l2, c2, b2, ctx.s2, g, grads)
        g_x = chain_bwd(l1, c1, b1, ctx.s1, g_y1n, grads)
        if ctx.s3 is None:
            g_x = _add(g_x, g)
        else:
            pool, l3, c3, b3 = block.shortcut
            g_p = chain_bwd(l3, c3, b3, ctx.s3, g, grads)
            g_sc = F_.maxpool_bwd(ctx.a, g_p, ctx.pool) if ctx.pool > 1 else g_p
            g_x = _add(g_x, g_sc)
        return (None, g_x.permute(0, 1, 4, 2, 3)) + _grad_tuple(block, grads)


class BasicBlockMsFn(torch.autograd.Function):
    """BasicBlock_ms in training mode (models/common.py:1658-1687): the shortcut is max-pool -> 1x1 conv on the REAL
    input -> tdBN, no neuron."""

    @staticmethod
    def forward(ctx, block, x_ref, *params):
        a = Act.from_ref(x_ref)
        l1, c1, b1, l2, c2, b2 = block.residual_function
        s1 = chain_fwd(l1, c1, b1, a, None)
        s2 = chain_fwd(l2, c2, b2, s1.y, (s1.scale, s1.shift))
        ctx.block, ctx.a, ctx.s1, ctx.s2 = block, a, s1, s2
        if len(block.shortcut) == 0:
            ctx.z = None
            out = F_.affine_add(s2.y, s2.scale, s2.shift, a, None, None)
        else:
            pool, conv, bn = block.shortcut
            ctx.pool = pool.stride[1]
            ctx.xp = F_.maxpool(a, ctx.pool).full()
            ctx.z = conv.conv_real(ctx.xp)
            sc, sh, ctx.mean, ctx.rstd = bn_train_fwd(bn, ctx.z)
            out = F_.affine_add(s2.y, s2.scale, s2.shift, ctx.z, sc, sh)
        return out.to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        from .common import _cached
        block = ctx.block
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        l1, c1, b1, l2, c2, b2 = block.residual_function
        g_y1n = chain_bwd(l2, c2, b2, ctx.s2, g, grads)
        g_x = chain_bwd(l1, c1, b1, ctx.s1, g_y1n, grads)
        if ctx.z is None:
            g_x = _add(g_x, g)
        else:
            pool, conv, bn = block.shortcut
            g_z = bn_train_bwd(bn, ctx.z, ctx.mean, ctx.rstd, g, grads)
            k, s, p = conv._geom()
            _acc(grads, conv.weight, F_.real_conv_wgrad(g_z, ctx.xp, k, s, p))
            splits = F_.get_splits()
            wT = _cached(conv, "dgradw", (conv.weight,), lambda: F_.pack_dgrad_weight(conv.weight, splits))
            g_xp = F_.conv_dgrad(g_z, wT, splits, ctx.xp.H, ctx.xp.W, conv.in_channels, k, s, p)
            g_sc = F_.maxpool_bwd(ctx.a, g_xp, ctx.pool) if ctx.pool > 1 else g_xp
            g_x = _add(g_x, g_sc)
        return (None, g_x.permute(0, 1, 4, 2, 3)) + _grad_tuple(block, grads)


class ConvBFn(torch.autograd.Function):
    """One Conv_B / Conv_2 (LIF -> conv -> tdBN) in training mode (models/common.py:393-406, 428-440)."""

    @staticmethod
    def forward(ctx, mod, x_ref, *params):
        a = Act.from_ref(x_ref)
        ctx.mod, ctx.need_gx = mod, x_ref.requires_grad
        ctx.s1 = chain_fwd(mod.act, mod.conv, mod.bn, a, None)
        return F_.affine_add(ctx.s1.y, ctx.s1.scale, ctx.s1.shift).to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        mod = ctx.mod
        g_x = chain_bwd(mod.act, mod.conv, mod.bn, ctx.s1, g, grads)
        gx_ref = g_x.permute(0, 1, 4, 2, 3) if ctx.need_gx else None
        return (None, gx_ref) + _grad_tuple(mod, grads)


class ConcatRes2Fn(torch.autograd.Function):
    """Concat_res2 in training mode (models/common.py:1454-1488)."""

    @staticmethod
    def forward(ctx, block, x_ref, *params):
        a = Act.from_ref(x_ref)
        l1, c1, b1, l2, c2, b2 = block.residual_function
        s1 = chain_fwd(l1, c1, b1, a, None)
        s2 = chain_fwd(l2, c2, b2, s1.y, (s1.scale, s1.shift))
        pool = block.pools.stride[1]
        ctx.block, ctx.a, ctx.s1, ctx.s2, ctx.pool = block, a, s1, s2, pool
        if len(block.shortcut) == 0:
            ctx.s3, ctx.temp = None, a
        else:
            l3, c3, b3 = block.shortcut
            ctx.s3 = chain_fwd(l3, c3, b3, a, None)
            ctx.temp = F_.affine_add(ctx.s3.y, ctx.s3.scale, ctx.s3.shift)
        sc = F_.concat_channels([ctx.temp, a], pool=pool)
        out = F_.affine_add(s2.y, s2.scale, s2.shift, sc, None, None)
        return out.to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        block = ctx.block
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        l1, c1, b1, l2, c2, b2 = block.residual_function
        g_y1n = chain_bwd(l2, c2, b2, ctx.s2, g, grads)
        g_x = chain_bwd(l1, c1, b1, ctx.s1, g_y1n, grads)
        ct = ctx.temp.C
        # shortcut = max-pool(cat(temp, x)): the pool acts per channel, so each source gets its own slice back
        g_temp = F_.maxpool_bwd(ctx.temp, g, ctx.pool, 0) if ctx.pool > 1 else F_.sumpool_slice(g, ct, 0, 1)
        g_x2 = F_.maxpool_bwd(ctx.a, g, ctx.pool, ct) if ctx.pool > 1 else F_.sumpool_slice(g, ctx.a.C, ct, 1)
        g_x = _add(g_x, g_x2)
        if ctx.s3 is None:
            g_x = _add(g_x, g_temp)
        else:
            l3, c3, b3 = block.shortcut
            g_x = _add(g_x, chain_bwd(l3, c3, b3, ctx.s3, g_temp, grads))
        return (None, g_x.permute(0, 1, 4, 2, 3)) + _grad_tuple(block, grads)


class ExpandT(torch.autograd.Function):
    """[1,N,C,H,W] -> the T-broadcast view [T,N,C,H,W] of a stem output.  torch's own expand backward sums over T with a
    generic reduction (2 ms per 320x320x64 frame batch at 1.4 TB/s); this one is the streaming ecsy_tsum kernel."""

    @staticmethod
    def forward(ctx, x, T):
        return x.expand(T, -1, -1, -1, -1)

    @staticmethod
    def backward(ctx, g_ref):
        g = Act.from_ref(g_ref)
        return F_.tsum(g, None, 1.0).unsqueeze(0).permute(0, 1, 4, 2, 3), None


class StemFn(torch.autograd.Function):
    """Conv_1 in training mode: real-input conv + tdBN (models/common.py:409-425).  The image needs no
    gradient.  For a T-broadcast input the output is ONE frame; the caller expands it, so autograd delivers
    the sum over T."""

    @staticmethod
    def forward(ctx, mod, x_ref, *params):
        a = Act.from_ref(x_ref)
        y = mod.conv.conv_real(a)
        scale, shift, mean, rstd = bn_train_fwd(mod.bn, y)
        ctx.mod, ctx.a, ctx.y, ctx.mean, ctx.rstd = mod, a, y, mean, rstd
        out = F_.affine_add(y, scale, shift)
        return out.data.permute(0, 1, 4, 2, 3)        # [Tp,N,C,H,W]

    @staticmethod
    def backward(ctx, g_ref):
        mod = ctx.mod
        g = Act.from_ref(g_ref).data                  # [Tp,N,H,W,C]
        grads: Dict[int, torch.Tensor] = {}
        g_y = bn_train_bwd(mod.bn, ctx.y, ctx.mean, ctx.rstd, g, grads)
        k, s, p = mod.conv._geom()
        _acc(grads, mod.conv.weight, F_.real_conv_wgrad(g_y, ctx.a, k, s, p))
        return (None, None) + _grad_tuple(mod, grads)


class ConvSiluFn(torch.autograd.Function):
    """class Conv in training mode: real-input conv -> tdBN -> SiLU neuron (models/common.py:362-375)."""

    @staticmethod
    def forward(ctx, mod, x_ref, *params):
        if not bool(getattr(mod.act.actFun, "inplace", False)):
            # ecsy_lif_silu_bwd differentiates the in-place recurrence (mem <- silu(mem) feeds the next step, what
            # Model.__init__ sets up, models/yolo.py:237-239); the out-of-place variant has a different gradient
            raise NotImplementedError("Conv (mem_update(act=True)): the BPTT backward is implemented for SiLU(inplace=True) "
                                      "only, the setting every reference model runs with")
        a = Act.from_ref(x_ref).full()
        y = mod.conv.conv_real(a)
        scale, shift, mean, rstd = bn_train_fwd(mod.bn, y)
        ctx.mod, ctx.a, ctx.y, ctx.mean, ctx.rstd, ctx.aff = mod, a, y, mean, rstd, (scale, shift)
        return mod.act.analog(y, (scale, shift)).to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        from .common import _cached
        mod = ctx.mod
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        lif = mod.act
        g_yn, gdw, gdb, gpw, gpb = F_.lif_silu_bwd(g, ctx.y, lif._weights(), lif.spread[1].weight, ctx.aff, lif.ecs_tau,
                                                   lif.alpha, lif.beta)
        _acc(grads, lif.spread[0].weight, gdw)
        _acc(grads, lif.spread[0].bias, gdb)
        _acc(grads, lif.spread[1].weight, gpw)
        _acc(grads, lif.spread[1].bias, gpb)
        g_y = bn_train_bwd(mod.bn, ctx.y, ctx.mean, ctx.rstd, g_yn, grads)
        k, s, p = mod.conv._geom()
        _acc(grads, mod.conv.weight, F_.real_conv_wgrad(g_y, ctx.a, k, s, p))
        splits = F_.get_splits()
        wT = _cached(mod.conv, "dgradw", (mod.conv.weight,), lambda: F_.pack_dgrad_weight(mod.conv.weight, splits))
        g_x = F_.conv_dgrad(g_y, wT, splits, ctx.a.H, ctx.a.W, mod.conv.in_channels, k, s, p)
        return (None, g_x.permute(0, 1, 4, 2, 3)) + _grad_tuple(mod, grads)


class ConvBChainFn(torch.autograd.Function):
    """Two chained Conv_B modules (LIF -> conv -> tdBN, twice) in training mode: the trunk of a DDetect branch
    (models/yolo_snn.py:100-107).  Returns the normalised output of the second one."""

    @staticmethod
    def forward(ctx, cb0, cb1, x_ref, *params):
        a = Act.from_ref(x_ref).full()
        s1 = chain_fwd(cb0.act, cb0.conv, cb0.bn, a, None)
        s2 = chain_fwd(cb1.act, cb1.conv, cb1.bn, s1.y, (s1.scale, s1.shift))
        ctx.cb0, ctx.cb1, ctx.s1, ctx.s2 = cb0, cb1, s1, s2
        return F_.affine_add(s2.y, s2.scale, s2.shift).to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        g = _to_nhwc(g_ref)
        grads: Dict[int, torch.Tensor] = {}
        cb0, cb1 = ctx.cb0, ctx.cb1
        g_y1n = chain_bwd(cb1.act, cb1.conv, cb1.bn, ctx.s2, g, grads)
        g_x = chain_bwd(cb0.act, cb0.conv, cb0.bn, ctx.s1, g_y1n, grads)
        params = list(cb0.parameters()) + list(cb1.parameters())
        return (None, None, g_x.permute(0, 1, 4, 2, 3)) + _grad_tuple_params(params, grads)


class SampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_ref, s):
        ctx.s = s
        return F_.upsample(Act.from_ref(x_ref), s).to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        g = _to_nhwc(g_ref)
        return F_.sumpool_slice(g, g.shape[4], 0, ctx.s).permute(0, 1, 4, 2, 3), None


class ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        acts = [Act.from_ref(t) for t in xs]
        ctx.cs = [a.C for a in acts]
        return F_.concat_channels(acts).to_ref()

    @staticmethod
    def backward(ctx, g_ref):
        g = _to_nhwc(g_ref)
        outs, off = [], 0
        for c in ctx.cs:
            outs.append(F_.sumpool_slice(g, c, off, 1).permute(0, 1, 4, 2, 3))
            off += c
        return tuple(outs)


def wants_grad(mod, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    if not mod.training:
        # the backward chains are built for training-mode tdBN (batch statistics); an eval-mode forward whose INPUT asks
        # for a gradient would silently come back without a grad_fn -- fail where the cause is
        if any(torch.is_tensor(t) and t.requires_grad for t in tensors):
            raise RuntimeError(f"{type(mod).__name__}: autograd through an eval-mode (running-statistics) forward is not "
                               "implemented; call .train(), or run under torch.no_grad() / detach the input")
        return False
    return any(t.requires_grad for t in tensors if torch.is_tensor(t)) or any(p.requires_grad for p in mod.parameters())
