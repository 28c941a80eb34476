"""Host logic of the two ComputeLoss drop-ins that needs no GPU: constructor contracts (what they read off the model,
utils/loss.py:131-160 / utils/loss_tal.py:107-140), argument validation, wrapper selection.  The arithmetic itself has
no CPU path (tests/test_abi.py::test_no_cpu_fallback); its parity tests are tests/test_gpu_post.py."""
import types

import pytest
import torch

import seeded as S
from util import ecsy


class _Holder(torch.nn.Module):
    def __init__(self, det, hyp):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.hyp = hyp
        self.model = [det]


def _det_a(nl=2):
    an = torch.tensor((S._ANCH3 if nl == 3 else S._ANCH2))
    return types.SimpleNamespace(na=3, nc=13, nl=nl, anchors=an, stride=torch.tensor([8.0, 16.0, 32.0][-nl:]))


def test_compute_loss_constructor_mirrors_reference():
    E = ecsy()
    c = E.loss.ComputeLoss(_Holder(_det_a(2), dict(S._HYP, label_smoothing=0.1)))
    assert (c.na, c.nc, c.nl) == (3, 13, 2) and c.anchors.shape == (2, 3, 2)
    assert c.balance[:2] == [4.0, 1.0] and len(c.balance) == 5           # nl != 3: the P3-P7 list (utils/loss.py:156)
    assert (c.cp, c.cn) == (0.95, 0.05) and c.gr == 1.0 and c.ssi == 0 and not c.autobalance
    c3 = E.loss.ComputeLoss(_Holder(_det_a(3), dict(S._HYP)), autobalance=True)
    assert c3.balance == [4.0, 1.0, 0.4] and c3.ssi == 1                 # stride 16 is level 1 of (8, 16, 32)
    # DDP / EMA style wrapper: `.module` is unwrapped like is_parallel / de_parallel do
    wrap = types.SimpleNamespace(module=_Holder(_det_a(2), dict(S._HYP)), hyp=dict(S._HYP))
    assert E.loss.ComputeLoss(wrap).nl == 2
    assert E.loss.ComputeLoss(_Holder(_det_a(2), dict(S._HYP, fl_gamma=1.5))).fl_gamma == 1.5
    assert E.loss.ComputeLoss(_Holder(_det_a(2), dict(S._HYP, slide_ratio=1.0)))._slide_state is None
    with pytest.raises(TypeError):
        E.loss.ComputeLoss(_Holder(_det_a(2), dict(S._HYP, fl_gamma=1.5, slide_ratio=1.0)))


def test_tal_compute_loss_constructor():
    E = ecsy()
    det = types.SimpleNamespace(nl=2, nc=80, no=144, reg_max=16, stride=torch.tensor([16.0, 32.0]))
    c = E.loss_tal.ComputeLoss(_Holder(det, dict(cls_pw=1.0, fl_gamma=0.0, label_smoothing=0.0)))
    assert (c.nc, c.nl, c.no, c.reg_max) == (80, 2, 144, 16) and c._strides == [16.0, 32.0] and c.use_dfl
    with pytest.raises(NotImplementedError):
        E.loss_tal.ComputeLoss(_Holder(det, dict(cls_pw=1.0, fl_gamma=0.0)), use_dfl=False)
    det8 = types.SimpleNamespace(nl=2, nc=80, no=112, reg_max=8, stride=torch.tensor([16.0, 32.0]))
    with pytest.raises(NotImplementedError):
        E.loss_tal.ComputeLoss(_Holder(det8, dict(cls_pw=1.0, fl_gamma=0.0)))


def test_assigner_env_overrides(monkeypatch):
    """utils/loss_tal.py:134-137: topk / alpha / beta from YOLOM / YOLOA / YOLOB at construction."""
    E = ecsy()
    det = types.SimpleNamespace(nl=2, nc=3, no=67, reg_max=16, stride=torch.tensor([16.0, 32.0]))
    assert E.loss_tal.ComputeLoss(_Holder(det, dict(cls_pw=1.0, fl_gamma=0.0))).assigner == (10, 0.5, 6.0)
    monkeypatch.setenv("YOLOM", "13")
    monkeypatch.setenv("YOLOA", "1.0")
    monkeypatch.setenv("YOLOB", "4.0")
    assert E.loss_tal.ComputeLoss(_Holder(det, dict(cls_pw=1.0, fl_gamma=0.0))).assigner == (13, 1.0, 4.0)
    monkeypatch.setenv("YOLOM", "0")
    with pytest.raises(ValueError):
        E.loss_tal.ComputeLoss(_Holder(det, dict(cls_pw=1.0, fl_gamma=0.0)))