"""torch.ops.ecsy.* on the GPU: the operators run the CUDA path (no torch fallback) and the autograd formula of
`lif_spike_conv` matches torch autograd through the CPU oracle (surrogate gradient, models/common.py:56-82)."""
import pytest
import torch

import ecs_oracle as O
import seeded as S
from util import agree, ecsy, rel_l2

pytestmark = pytest.mark.gpu


def _nhwc(t):   # [T,N,C,H,W] cpu -> [T,N,H,W,C] cuda contiguous
    return t.cuda().permute(0, 1, 3, 4, 2).contiguous()


def test_ops_forward_match_functional():
    E = ecsy()
    F = E.functional
    ns = torch.ops.ecsy
    inp = S.lif_inputs(dict(T=4, N=2, C=64, H=10, W=12, seed=11))
    x = _nhwc(inp["x"])
    p = [inp[k].cuda() for k in ("dw_w", "dw_b", "pw_w", "pw_b")]
    n0 = F.launches["n"]
    bits = ns.lif_ecs(x, 4, *p, None, None, 5.0, 0.75, 0.25)
    assert F.launches["n"] > n0, "the operator must launch the CUDA kernels"
    want = O.ecs_lif(inp["x"], inp["dw_w"], inp["dw_b"], inp["pw_w"], inp["pw_b"])
    got = F.Spikes(bits, 64).to_act().to_ref().cpu()
    assert agree(got, want) >= 0.9995
    g = S.gen(5)
    w = torch.randn(128, 64, 3, 3, generator=g) * 0.05
    sc, sh = torch.rand(128, generator=g) + 0.5, torch.rand(128, generator=g)
    y = ns.spike_conv(bits, 64, w.cuda(), sc.cuda(), sh.cuda(), None, 2, 1)
    ref = O.snn_conv2d(got, w, None, 2, 1) * sc.view(1, 1, -1, 1, 1) + sh.view(1, 1, -1, 1, 1)
    assert rel_l2(y.permute(0, 1, 4, 2, 3).cpu(), ref) < 2e-5
    mean, var = ns.tdbn_stats(y)
    yc = y.reshape(-1, 128).double().cpu()
    assert torch.allclose(mean.cpu().double(), yc.mean(0), atol=1e-5)
    assert torch.allclose(var.cpu().double(), yc.var(0, unbiased=False), rtol=1e-4, atol=1e-6)


def test_lif_spike_conv_autograd():
    E = ecsy()
    ns = torch.ops.ecsy
    T, N, C, H, W, Co = 4, 2, 64, 9, 11, 64
    inp = S.lif_inputs(dict(T=T, N=N, C=C, H=H, W=W, seed=21))
    g = S.gen(22)
    w = torch.randn(Co, C, 3, 3, generator=g) * 0.05
    gy = torch.randn(T, N, Co, H, W, generator=g)
    # CPU oracle with torch autograd
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("x", "dw_w", "dw_b", "pw_w", "pw_b")}
    wl = w.clone().requires_grad_(True)
    s = O.ecs_lif(leaves["x"], leaves["dw_w"], leaves["dw_b"], leaves["pw_w"], leaves["pw_b"])
    y_ref = O.snn_conv2d(s, wl, None, 1, 1)
    (y_ref * gy).sum().backward()
    # ours
    x = _nhwc(inp["x"]).requires_grad_(True)
    p = [inp[k].cuda().requires_grad_(True) for k in ("dw_w", "dw_b", "pw_w", "pw_b")]
    wc = w.cuda().requires_grad_(True)
    y, bits = ns.lif_spike_conv(x, T, *p, wc, 1, 1, 5.0, 0.75, 0.25)
    assert rel_l2(y.permute(0, 1, 4, 2, 3).detach().cpu(), y_ref.detach()) < 1e-3
    (y * _nhwc(gy)).sum().backward()
    assert rel_l2(x.grad.permute(0, 1, 4, 2, 3).cpu(), leaves["x"].grad) < 5e-3
    assert rel_l2(wc.grad.cpu(), wl.grad) < 5e-3
    for t, k in zip(p, ("dw_w", "dw_b", "pw_w", "pw_b")):
        assert rel_l2(t.grad.cpu(), leaves[k].grad) < 5e-3, k


def test_loss_ops_match_direct_calls():
    """torch.ops.ecsy.yolo_loss / tal_loss return exactly what loss.yolo_loss / loss_tal.tal_loss return (same C-ABI call)."""
    E = ecsy()
    ns = torch.ops.ecsy
    spec = S.LOSS_CASES["loss_basic"]
    li = S.loss_inputs(spec)
    h = spec["hyp"]
    p = [x.cuda() for x in li["p"]]
    tg, an = li["targets"].cuda(), li["anchors"].cuda()
    out, grads = ns.yolo_loss(p, tg, an, [4.0, 1.0], h["box"], h["obj"], h["cls"], 1.0, 1.0, 1.0, 0.0, h["anchor_t"], 1.0)
    out2, grads2 = E.loss.yolo_loss(p, tg, an, [4.0, 1.0], h["box"], h["obj"], h["cls"], anchor_t=h["anchor_t"])
    assert torch.equal(out, out2)
    # cells hit by several matches accumulate with float atomics: equal up to summation order
    assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-8) for a, b in zip(grads, grads2))
    ts = S.TAL_CASES["tal_basic"]
    ti = S.tal_inputs(ts)
    f = [x.cuda() for x in ti["feats"]]
    out, grads = ns.tal_loss(f, ti["targets"].cuda(), ts["strides"], 1.0, 7.5, 0.5, 1.5)
    out2, grads2 = E.loss_tal.tal_loss(f, ti["targets"].cuda(), ts["strides"])
    assert torch.equal(out, out2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))
