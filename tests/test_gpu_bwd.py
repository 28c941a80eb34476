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


@pytest.mark.parametrize("name", list(S.LIF_CASES))
def test_lif_ecs_bwd(name):
    """Surrogate-gradient BPTT vs the reference's autograd (golden gx and spread-parameter gradients)."""
    E = ecsy()
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
    assert e < 2e-3, f"{name}: gx rel-L2 {e:.3e}"
    if spec["T"] > 1:
        for got, k in [(g_dw_w, "g_dw_w"), (g_dw_b, "g_dw_b"), (g_pw_w, "g_pw_w"), (g_pw_b, "g_pw_b")]:
            e = rel_l2(got.cpu(), gold[k])
            assert e < 2e-3, f"{name}: {k} rel-L2 {e:.3e}"


def test_colsum2():
    E = ecsy()
    F = E.functional
    g = torch.randn(3000, 128, device="cuda")
    x = torch.randn(3000, 128, device="cuda")
    sg, sgx = F.colsum2(g, x, 128)
    assert torch.allclose(sg.cpu(), g.double().sum(0).float().cpu(), rtol=1e-5, atol=1e-4)
    assert torch.allclose(sgx.cpu(), (g.double() * x.double()).sum(0).float().cpu(), rtol=1e-5, atol=1e-4)
