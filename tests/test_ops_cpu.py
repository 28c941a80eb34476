"""torch.library custom operators: schemas are registered and the fake (meta) implementations infer the right
shapes / dtypes without a GPU (this is what torch.export / torch.compile tracing uses)."""
import torch

from util import ecsy


def test_custom_ops_registered_and_fake_shapes():
    ecsy()
    ns = torch.ops.ecsy
    for name in ("lif_ecs", "spike_conv", "tdbn_stats", "lif_spike_conv"):
        assert hasattr(ns, name), name
    T, N, H, W, C, Co = 4, 2, 10, 12, 64, 128
    m = "meta"
    x = torch.empty(T, N, H, W, C, device=m)
    dw_w, dw_b = torch.empty(C, 1, 3, 3, device=m), torch.empty(C, device=m)
    pw_w, pw_b = torch.empty(C, C, 1, 1, device=m), torch.empty(C, device=m)
    bits = ns.lif_ecs(x, T, dw_w, dw_b, pw_w, pw_b, None, None, 5.0, 0.75, 0.25)
    assert bits.shape == (T, N, H, W, C // 32) and bits.dtype == torch.int32
    # a T-broadcast input (one stored frame) still yields T spike frames
    assert ns.lif_ecs(x[:1], T, dw_w, dw_b, pw_w, pw_b, None, None, 5.0, 0.75, 0.25).shape[0] == T
    w = torch.empty(Co, C, 3, 3, device=m)
    y = ns.spike_conv(bits, C, w, None, None, None, 2, 1)
    assert y.shape == (T, N, 5, 6, Co) and y.dtype == torch.float32
    mean, var = ns.tdbn_stats(y)
    assert mean.shape == (Co,) and var.shape == (Co,)
    y2, b2 = ns.lif_spike_conv(x, T, dw_w, dw_b, pw_w, pw_b, w, 1, 1, 5.0, 0.75, 0.25)
    assert y2.shape == (T, N, H, W, Co) and b2.shape == bits.shape


def test_loss_ops_fake_shapes():
    ecsy()
    ns = torch.ops.ecsy
    m = "meta"
    p = [torch.empty(4, 3, 40, 40, 18, device=m), torch.empty(4, 3, 20, 20, 18, device=m)]
    out, grads = ns.yolo_loss(p, torch.empty(9, 6, device=m), torch.empty(2, 3, 2, device=m), [4.0, 1.0], 0.05, 1.0, 0.5,
                              1.0, 1.0, 1.0, 0.0, 4.0, 1.0)
    assert out.shape == (6,) and [g.shape for g in grads] == [x.shape for x in p]
    f = [torch.empty(4, 144, 40, 40, device=m), torch.empty(4, 144, 20, 20, device=m)]
    out, grads = ns.tal_loss(f, torch.empty(9, 6, device=m), [16.0, 32.0], 1.0, 7.5, 0.5, 1.5)
    assert out.shape == (6,) and [g.shape for g in grads] == [x.shape for x in f]


def test_custom_ops_have_no_cpu_kernel():
    ecsy()
    import pytest
    x = torch.zeros(1, 1, 2, 2, 64)
    with pytest.raises(RuntimeError):
        torch.ops.ecsy.tdbn_stats(x)
    with pytest.raises(RuntimeError):
        torch.ops.ecsy.yolo_loss([torch.zeros(1, 3, 4, 4, 8)], torch.zeros(1, 6), torch.ones(1, 3, 2), [4.0], 0.05, 1.0, 0.5,
                                 1.0, 1.0, 1.0, 0.0, 4.0, 1.0)
