"""Checkpoint compatibility (models/experimental.py:87-127): checkpoints PICKLED by the unmodified
reference (oracle/gen_golden_ckpt.py -> tests/golden/ckpt_micro_*.pt: whole `models.yolo.Model` /
`models.yolo_snn.DetectionModel` objects, half precision, `{'model', 'ema', ...}` like train.py:659-669) load through
`experimental.attempt_load` in a process that has NO reference tree on its path, and come back as this package's models
with the checkpoint's weights, plan and attributes.  No kernel is launched (the forward has no CPU path)."""
import os
import subprocess
import sys

import pytest
import torch

import seeded as S
from util import ROOT, ecsy


@pytest.mark.parametrize("name,cls", [("micro_a", "Model"), ("micro_b", "DetectionModel")])
def test_attempt_load_reference_pickle(name, cls):
    E = ecsy()
    meta = torch.load(os.path.join(S.GOLDEN_DIR, f"ckpt_{name}_meta.pt"), weights_only=False)
    m = E.experimental.attempt_load(os.path.join(S.GOLDEN_DIR, f"ckpt_{name}.pt"), map_location="cpu")
    assert type(m).__name__ == cls and type(m).__module__.startswith("ecs-yolo_b200")
    assert not m.training
    sd = m.state_dict()
    assert list(sd.keys()) == meta["keys"]
    assert all(v.dtype == torch.float32 for v in sd.values() if v.is_floating_point())     # .float() (experimental.py:96)
    chk = S.sd_checksum({k: v for k, v in sd.items() if v.is_floating_point()})
    assert abs(chk - meta["chk"]) <= 1e-9 * abs(meta["chk"])
    assert [type(x).__name__ for x in m.model] == meta["types"]
    assert torch.equal(m.stride.float().cpu(), meta["stride"].float())
    assert m.names == {i: f"cls{i}" for i in range(3)} and m.hyp == {"box": 0.05, "cls": 0.5}
    ens = E.experimental.attempt_load([os.path.join(S.GOLDEN_DIR, f"ckpt_{name}.pt")] * 2, map_location="cpu")
    assert len(ens) == 2 and torch.equal(ens.stride, m.stride)


def test_attempt_load_needs_no_reference_tree():
    """A fresh interpreter whose sys.path holds only this repo: the pickles resolve `models.*` to the drop-in classes."""
    code = (
        "import sys, os\n"
        "assert not any('reference' in p for p in sys.path)\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import ecs_yolo_b200 as E\n"
        f"m = E.experimental.attempt_load(os.path.join({S.GOLDEN_DIR!r}, 'ckpt_micro_a.pt'), map_location='cpu')\n"
        "assert 'models' not in sys.modules and 'models.yolo' not in sys.modules\n"
        "print(type(m).__module__, type(m).__name__, len(m.state_dict()))\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ecs-yolo_b200.yolo Model" in out.stdout


def test_unknown_class_is_reported():
    import io
    import pickle

    class Focus(torch.nn.Module):
        pass
    Focus.__module__ = "models.common"
    Focus.__qualname__ = "Focus"
    import types
    mod = types.ModuleType("models.common")
    mod.Focus = Focus
    sys.modules.setdefault("models", types.ModuleType("models"))
    sys.modules["models.common"] = mod
    try:
        buf = io.BytesIO()
        torch.save({"model": Focus(), "ema": None}, buf)
    finally:
        sys.modules.pop("models.common", None)
        sys.modules.pop("models", None)
    buf.seek(0)
    E = ecsy()
    with pytest.raises(pickle.UnpicklingError, match="Focus"):
        E.experimental.load_checkpoint(buf)
