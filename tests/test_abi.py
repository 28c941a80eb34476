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
    assert lib.ecsy_abi_version() == 1


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
