"""CUDA-graph replay of the eval forward (ecs-yolo_b200/graph.py): bit-identical to the eager forward, re-captured after a
parameter update.  The reference runs eager PyTorch (models/yolo.py:247-312); this is the small-batch serving path."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def ecsy():
    return importlib.import_module("ecs-yolo_b200")


def _model(E, name="resnet10"):
    torch.manual_seed(0)
    m = E.yolo.Model(E.cfg_path(name)).cuda().eval()
    return m


@pytest.mark.parametrize("mode", ["fast", "parity"])
def test_graph_replay_equals_eager(mode):
    E = ecsy()
    E.set_precision(mode)
    try:
        m = _model(E)
        g = torch.Generator(device="cuda").manual_seed(1)
        xs = [torch.rand(1, 3, 128, 128, device="cuda", generator=g) for _ in range(3)]
        gf = E.graph.GraphedForward(m, xs[0])
        assert gf.launches_per_replay > 50
        for x in xs:
            with torch.no_grad():
                z_e, f_e = m(x)
            z_g, f_g = gf(x, clone=True)
            assert torch.equal(z_e, z_g)
            for a, b in zip(f_e, f_g):
                assert torch.equal(a, b)
        assert gf.captures == 1
    finally:
        E.set_precision("parity")


def test_graph_recaptures_after_weight_update():
    E = ecsy()
    m = _model(E)
    x = torch.rand(1, 3, 128, 128, device="cuda")
    gf = E.graph.GraphedForward(m, x)
    z0 = gf(x, clone=True)[0]
    with torch.no_grad():
        for p in m.parameters():      # includes the head biases: the decoded output must move
            p.add_(0.01)
    z1 = gf(x, clone=True)[0]
    assert gf.captures == 2
    with torch.no_grad():
        z_e = m(x)[0]
    assert torch.equal(z1, z_e)
    assert not torch.equal(z0, z1)


def test_graph_rejects_other_shapes_and_training_mode():
    E = ecsy()
    m = _model(E)
    x = torch.rand(1, 3, 128, 128, device="cuda")
    gf = E.graph.GraphedForward(m, x)
    with pytest.raises(RuntimeError):
        gf(torch.rand(2, 3, 128, 128, device="cuda"))
    m.train()
    with pytest.raises(RuntimeError):
        gf(x)
    with pytest.raises(RuntimeError):
        E.graph.GraphedForward(m, x)
