"""ecs-yolo_b200: B200-native (sm_100a) implementation of ECS-YOLO's spiking hot path.

    import importlib; ecsy = importlib.import_module("ecs-yolo_b200")      # or: import ecs_yolo_b200 as ecsy
    model = ecsy.yolo.Model(ecsy.cfg_path("resnet34")).cuda().eval()

The package directory carries the repo's hyphenated name, so it is imported through importlib (or the
``ecs_yolo_b200`` alias module at the repo root).  Sub-modules: ``common`` (drop-in layers), ``yolo``
(Stack A model / Detect), ``general`` (batched NMS), ``loss`` / ``loss_tal`` (ComputeLoss of Stack A / Stack B), ``functional`` (tensor-level ops over the C ABI), ``_cabi`` (ctypes binding).
"""
import os as _os

from . import _cabi, functional, common, autograd, yolo, yolo_snn, dist, general, optim, events, loss, loss_tal, experimental, graph  # noqa: F401
from .functional import set_precision  # noqa: F401
from .convert import convert  # noqa: F401
from . import ops  # noqa: F401  (registers torch.ops.ecsy.*)

__all__ = ["common", "yolo", "yolo_snn", "functional", "general", "optim", "events", "loss", "loss_tal", "experimental", "graph", "ops", "set_precision", "convert", "cfg_path", "build_library"]


def cfg_path(name: str) -> str:
    return _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "cfg", name + ".yaml")


def build_library(force: bool = False) -> str:
    from . import build as _b
    return _b.build(force=force)
