"""Checkpoint I/O compatibility: `attempt_load` with the reference's signature
(models/experimental.py:87-127) for the reference's PICKLED checkpoints -- `{'model': Model, 'ema': Model, ...}`
written by train.py:659-669 / train2.py -- WITHOUT the reference tree on `sys.path`.

A pickled `nn.Module` stores class references by module path (`models.yolo.Model`, `models.common.BasicBlock_2`,
`models.yolo_snn.DDetect`, ...).  The un-pickler below resolves every `models.*` name to this package's drop-in
class of the same name (same attribute layout: parameters, buffers, sub-module names), so the checkpoint's own object
graph comes back as instances of OUR classes; `convert()` then rebuilds a clean model from the plan the checkpoint
carries (`model.yaml`) and copies weights and attributes, exactly as for a live reference model.  `utils.*` names
(e.g. a pickled `utils.loss` object inside an optimizer state) resolve to inert placeholders.
"""
from __future__ import annotations

import pickle
import types
from typing import List, Union

import torch
import torch.nn as nn

from . import common, yolo, yolo_snn
from .convert import convert

_NAMESPACES = {"models.yolo": (yolo, common), "models.yolo_snn": (yolo_snn, common), "models.common": (common,),
               "models.experimental": (common,)}


class _Placeholder:
    """Stands in for reference-side helper objects that carry no weights (loggers, loss objects, paths)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


# Reference-side helper layers that only ever appear NESTED inside a drop-in layer (never as a YAML row) and whose
# drop-in parent builds its own equivalent: they un-pickle as plain parameter containers (state_dict keys intact).
# models/common.py:627-645 `Conv2d` is mem_update's `spread` member (a stock nn.Conv2d here).
_HELPER_MODULES = {"Conv2d", "Conv3d"}                  # Conv3d: Conv_7.conv (models/common.py:549-562)
_helper_cache = {}


def _helper_class(module, name):
    key = (module, name)
    if key not in _helper_cache:
        _helper_cache[key] = type(name, (nn.Module,), {"__module__": module})
    return _helper_cache[key]


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in _NAMESPACES:
            for ns in _NAMESPACES[module]:
                if hasattr(ns, name):
                    return getattr(ns, name)
            if name in _HELPER_MODULES:
                return _helper_class(module, name)
            raise pickle.UnpicklingError(f"checkpoint uses {module}.{name}: no B200 drop-in class of that name")
        if module.split(".")[0] in ("models", "utils"):
            return type(name, (_Placeholder,), {"__module__": module})
        return super().find_class(module, name)


_pickle_module = types.SimpleNamespace(Unpickler=_RefUnpickler, load=lambda f, **k: _RefUnpickler(f, **k).load(),
                                       __name__="pickle")


def load_checkpoint(path, map_location="cpu") -> dict:
    """torch.load of a reference checkpoint with `models.*` classes mapped onto this package (no reference tree needed).
    The file is a pickle: load only checkpoints you trust, as with the reference's own `torch.load`."""
    return torch.load(path, map_location=map_location, pickle_module=_pickle_module, weights_only=False)


class Ensemble(nn.ModuleList):
    """models/experimental.py:68-84: NMS ensemble (concatenated predictions)."""

    def forward(self, x, augment=False, profile=False, visualize=False):
        y = [module(x)[0] for module in self]
        return torch.cat(y, 1), None


def attempt_load(weights: Union[str, List[str]], map_location=None, inplace=True, fuse=False):
    """Drop-in for models/experimental.py:87 `attempt_load`: a single model for one path, an `Ensemble` for a list.
    `fuse` is accepted for signature compatibility: eval-mode tdBN is always folded into the conv epilogue here.
    map_location: where the result lives (default cuda:0 -- the forward has no CPU path)."""
    model = Ensemble()
    for w in weights if isinstance(weights, (list, tuple)) else [weights]:
        ckpt = load_checkpoint(w, "cpu")
        ref = ckpt.get("ema") or ckpt["model"] if isinstance(ckpt, dict) else ckpt      # experimental.py:96
        ref = ref.float()
        if not hasattr(ref, "stride"):
            ref.stride = torch.tensor([32.])
        m = convert(ref, device=map_location)
        if hasattr(m, "names") and isinstance(m.names, (list, tuple)):
            m.names = dict(enumerate(m.names))                                          # experimental.py:102-103
        m.inplace = inplace
        model.append(m.eval())
    if len(model) == 1:
        return model[-1]
    model.names = model[-1].names
    model.stride = model[int(torch.argmax(torch.tensor([float(m.stride.max()) for m in model])))].stride
    return model
