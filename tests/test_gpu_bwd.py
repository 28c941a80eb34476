"""GPU tests of the backward building blocks (surrogate-gradient BPTT)."""
import pytest
import torch

import ecs_oracle as O
import seeded as S
from util import ecsy, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _planes(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("rows,Ca,Cb", [(1000, 64, 64), (4096, 128, 256), (777, 64, 192), (5000, 512, 512), (300, 1024, 1024)])
@pytest.mark.parametrize("split", [1, 2])
def test_xty(rows, Ca, Cb, split):
    E = ecsy()
    L = E._cabi.lib()
    g = torch.Generator(device="cuda").manual_seed(rows + Ca)
    P = torch.randn(rows, Ca, device="cuda", generator=g)
    Q = torch.randn(rows, Cb, device="cuda", generator=g)
    ph, pl = _planes(P)
    qh, ql = _planes(Q)
    out = torch.zeros(Ca, Cb, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    E._cabi.check(L.ecsy_xty_bf16(ph.data_ptr(), pl.data_ptr() if split == 2 else None, qh.data_ptr(),
                                  ql.data_ptr() if split == 2 else None, rows, Ca, Cb, 0.75, out.data_ptr(), st), "xty")
    want = 0.75 * (P.double().t() @ Q.double())
    err = rel_l2(out, want)
    assert err < (3e-5 if split == 2 else 6e-3), err


@pytest.mark.parametrize("mode,tol", [("parity", 2e-3), ("fast", 4e-2)])
@pytest.mark.parametrize("name", list(S.LIF_CASES))
def test_lif_ecs_bwd(name, mode, tol):
    """Surrogate-gradient BPTT vs the reference's autograd (golden gx and spread-parameter gradients), in both
    precisions (fast: one bf16 plane for the point-wise spread weights and the gradient operands of its GEMMs -- measured on
    B200: 0.9-2.4e-2 rel-L2 on these fixtures, 2e-7 / 4e-6 in parity precision)."""
    E = ecsy()
    E.set_precision(mode)
    try:
        _lif_ecs_bwd(E, name, tol)
    finally:
        E.set_precision("parity")


def _lif_ecs_bwd(E, name, tol):
    F = E.functional
    spec, gold = S.LIF_CASES[name], load_golden(name)
    inp = S.lif_inputs(spec)
    w = F.make_lif_w(inp["dw_w"].cuda(), inp["dw_b"].cuda(), inp["pw_w"].cuda(), inp["pw_b"].cuda())
    x = F.Act.from_ref(inp["x"].cuda())
    gout = inp["gout"].cuda().permute(0, 1, 3, 4, 2).contiguous()
    gx, g_dw_w, g_dw_b, g_pw_w, g_pw_b = F.lif_ecs_bwd(gout, x, w, inp["pw_w"].cuda())
    got_gx = gx.permute(0, 1, 4, 2, 3).cpu()
    # a flipped near-threshold spike changes the surrogate window of a few elements: compare in rel-L2
    e = rel_l2(got_gx, gold["gx"])
    errs = {"gx": e}
    if spec["T"] > 1:
        for got, k in [(g_dw_w, "g_dw_w"), (g_dw_b, "g_dw_b"), (g_pw_w, "g_pw_w"), (g_pw_b, "g_pw_b")]:
            errs[k] = rel_l2(got.cpu(), gold[k])
    print(f"\n{name}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < tol, f"{name}: {k} rel-L2 {v:.3e}"


def test_colsum2():
    E = ecsy()
    F = E.functional
    g = torch.randn(3000, 128, device="cuda")
    x = torch.randn(3000, 128, device="cuda")
    sg, sgx = F.colsum2(g, x, 128)
    assert torch.allclose(sg.cpu(), g.double().sum(0).float().cpu(), rtol=1e-5, atol=1e-4)
    assert torch.allclose(sgx.cpu(), (g.double() * x.double()).sum(0).float().cpu(), rtol=1e-5, atol=1e-4)


CONV_BWD = [  # ci, co, k, s, H, W, N, T
    (64, 128, 3, 1, 10, 12, 2, 2), (64, 64, 3, 2, 12, 10, 2, 2), (128, 64, 1, 1, 6, 8, 2, 2),
    (192, 256, 3, 1, 9, 9, 1, 2), (128, 256, 3, 2, 16, 16, 3, 1), (512, 128, 3, 1, 8, 8, 2, 1)]


@pytest.mark.parametrize("ci,co,k,s,H,W,N,T", CONV_BWD)
@pytest.mark.parametrize("mode,tol", [("parity", 3e-5), ("fast", 8e-3)])
def test_spike_conv_backward(ci, co, k, s, H, W, N, T, mode, tol):
    """dgrad + wgrad of the spike conv vs torch autograd of F.conv2d (fp32 reference)."""
    import torch.nn.functional as tF
    E = ecsy()
    F = E.functional
    F.set_precision(mode)
    try:
        g = S.gen(ci * 7 + co + k + s)
        x = (torch.rand(T, N, ci, H, W, generator=g) < 0.2).float()
        w = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5)
        p = k // 2
        xr = x.reshape(T * N, ci, H, W).clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        y = tF.conv2d(xr, wr, None, s, p)
        gy = torch.randn(*y.shape, generator=g)
        y.backward(gy)
        Ho, Wo = y.shape[2], y.shape[3]
        gy_nhwc = gy.reshape(T, N, co, Ho, Wo).permute(0, 1, 3, 4, 2).contiguous().cuda()
        sp = F.Spikes.from_act(F.Act.from_ref(x.cuda()))
        dw = F.spike_conv_wgrad(gy_nhwc, sp, k, s, p).cpu()
        e_w = rel_l2(dw, wr.grad)
        wT = F.pack_dgrad_weight(w.cuda(), F.get_splits())
        gx = F.conv_dgrad(gy_nhwc, wT, F.get_splits(), H, W, ci, k, s, p).cpu()
        want_gx = xr.grad.reshape(T, N, ci, H, W).permute(0, 1, 3, 4, 2)
        e_x = rel_l2(gx, want_gx)
        assert e_w < tol and e_x < tol, f"wgrad {e_w:.3e} dgrad {e_x:.3e}"
    finally:
        F.set_precision("parity")


def test_real_conv_wgrad():
    import torch.nn.functional as tF
    E = ecsy()
    F = E.functional
    g = S.gen(99)
    for (ci, co, k, s, H, W, N) in [(3, 64, 7, 2, 32, 32, 2), (128, 64, 3, 1, 6, 6, 2)]:
        x = torch.rand(N, ci, H, W, generator=g)
        w = torch.randn(co, ci, k, k, generator=g).requires_grad_(True)
        y = tF.conv2d(x, w, None, s, k // 2)
        gy = torch.randn(*y.shape, generator=g)
        y.backward(gy)
        a = F.Act.from_ref(x.unsqueeze(0).cuda())
        gyn = gy.permute(0, 2, 3, 1).contiguous().unsqueeze(0).cuda()
        dw = F.real_conv_wgrad(gyn, a, k, s, k // 2).cpu()
        e = rel_l2(dw, w.grad)
        assert e < 3e-5, e
