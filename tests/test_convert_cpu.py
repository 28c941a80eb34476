"""convert(model): a reference model built by the UNMODIFIED reference code is rebuilt with the
drop-in classes -- same module tree, identical state_dict, attributes carried over.  Needs /root/reference
(container only; skipped on the GPU box).  No kernel is launched: the forward has no CPU path."""
import os

import pytest
import torch

from util import ecsy

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.mark.parametrize("stack,cfg", [("A", "resnet10.yaml"), ("B", "resnet18.yaml"), ("A", "res10-ee.yaml"),
                                       ("A", "res18-ee.yaml")])
def test_convert_reference_model(stack, cfg):
    import ref_shim
    C, Y, S = ref_shim.load(4)
    torch.manual_seed(0)
    ref = (Y.Model if stack == "A" else S.DetectionModel)(os.path.join(REF, "models", cfg))
    ref.names = [f"cls{i}" for i in range(ref.yaml["nc"])]
    ref.hyp = {"box": 0.05}
    ref.eval()
    E = ecsy()
    ours = E.convert(ref, device="cpu")
    assert type(ours).__name__ == type(ref).__name__
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys())
    for k in sd_r:
        assert sd_r[k].shape == sd_o[k].shape and torch.equal(sd_r[k].float(), sd_o[k].float()), k
    assert ours.names == ref.names and ours.hyp == ref.hyp and not ours.training
    assert [type(m).__name__ for m in ours.model] == [type(m).__name__ for m in ref.model]
    assert [(m.i, m.f) for m in ours.model] == [(m.i, m.f) for m in ref.model]
    assert torch.equal(ours.stride.float().cpu(), ref.stride.float().cpu())
    # the reference object is untouched and shares no storage with the result
    p_r = next(ref.parameters()); p_o = next(ours.parameters())
    assert p_r.data_ptr() != p_o.data_ptr()
    # wrappers are unwrapped
    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__(); self.module = m
    assert type(E.convert(Wrap(ref), device="cpu")).__name__ == type(ref).__name__


def test_convert_rejects_unknown_plan():
    E = ecsy()
    with pytest.raises(TypeError):
        E.convert(torch.nn.Linear(2, 2))
    cfg = {"nc": 2, "depth_multiple": 1.0, "width_multiple": 1.0, "anchors": [[1, 2, 3, 4, 5, 6]],
           "backbone": [[-1, 1, "Focus", [64, 3]]], "head": [[-1, 1, "Detect", ["nc", "anchors"]]]}
    m = torch.nn.Module(); m.yaml = cfg
    with pytest.raises(NotImplementedError):
        E.convert(m)
