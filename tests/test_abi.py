"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/ecsy.h declares, the ctypes table matches the header, and the product never imports the oracle."""
import ctypes
import os
import re

import pytest

from util import ROOT, ecsy


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ecsy.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecsy_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    E = ecsy()
    path = E._cabi.LIB_PATH
    if not os.path.exists(path):
        E.build_library()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ecsy.h but not exported"
    assert lib.ecsy_abi_version() == 2


def test_ctypes_table_matches_header():
    E = ecsy()
    assert sorted(E._cabi.SIGNATURES) == _header_symbols()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ecs-yolo_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "ecs_oracle" not in txt and "oracle/" not in txt, f


def test_no_cpu_fallback():
    import torch
    E = ecsy()
    with pytest.raises(RuntimeError):
        E.functional.Act.from_ref(torch.zeros(1, 1, 64, 2, 2))
    with pytest.raises(RuntimeError):      # the training losses have no CPU path either
        E.loss.yolo_loss([torch.zeros(1, 3, 4, 4, 8)], torch.zeros(1, 6), torch.ones(1, 3, 2), [4.0], 0.05, 1.0, 0.5)
    with pytest.raises(RuntimeError):
        E.loss_tal.tal_loss([torch.zeros(1, 67, 4, 4)], torch.zeros(1, 6), [16.0])


def test_state_dict_keys_match_reference_layout():
    import yaml
    import ecs_oracle as O
    E = ecsy()
    for name in ("resnet10", "resnet34", "tiny"):
        m = E.yolo.Model(E.cfg_path(name))
        cfg = yaml.safe_load(open(E.cfg_path(name)))
        want = O.init_state_dict(cfg, 4)
        got = m.state_dict()
        assert set(got) == set(want)
        assert all(got[k].shape == want[k].shape for k in got)


def test_dispatch_rules_are_pure_host_functions():
    """The measured dispatch rules are part of the C ABI and need no device: tensor-memory operand for every layer it
    supports except the wide ones (Cout % 256 == 0 and Cin >= 256 with single-plane weights), fused LIF only for C = 64."""
    L = ecsy()._cabi.lib()
    assert L.ecsy_spike_conv_ts_supported(64, 64) == 1 and L.ecsy_spike_conv_ts_supported(64, 192) == 1
    assert L.ecsy_spike_conv_ts_supported(32, 64) == 0 and L.ecsy_spike_conv_ts_supported(64, 96) == 0
    assert L.ecsy_spike_conv_prefers_ts(64, 64, 1) == 1 and L.ecsy_spike_conv_prefers_ts(128, 256, 1) == 1
    assert L.ecsy_spike_conv_prefers_ts(512, 512, 1) == 0 and L.ecsy_spike_conv_prefers_ts(384, 256, 1) == 0
    assert L.ecsy_spike_conv_prefers_ts(512, 512, 2) == 1      # split weights: 128-column tiles either way
    assert L.ecsy_lif_ecs_fused_supported(4, 64) == 1 and L.ecsy_lif_ecs_fused_supported(1, 64) == 0
    assert L.ecsy_lif_ecs_fused_supported(4, 128) == 0 and L.ecsy_lif_ecs_fused_supported(9, 64) == 0
    # the stem's operand tiles are gathered on the fly in fast mode: no im2col workspace
    assert L.ecsy_real_conv_ws_bytes(64, 640, 640, 3, 64, 7, 2, 3, 1, 1) == 0
    assert L.ecsy_real_conv_ws_bytes(64, 640, 640, 3, 64, 7, 2, 3, 1, 2) > 0
