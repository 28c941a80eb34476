"""convert(model): move an already-built (or un-pickled) REFERENCE model onto the B200 path.

SURVEY section 8(b)(ii): the reference resolves its layers by bare class name from the model YAML
(models/yolo.py:454, models/yolo_snn.py:841) and ships trained weights as pickled ``nn.Module`` checkpoints
(models/experimental.py:94-96), which carry their YAML dict as ``model.yaml``.  ``convert`` rebuilds the same plan
with this package's drop-in classes (same constructor signatures, same ``state_dict`` keys), copies every
parameter and buffer, and carries over the attributes the reference's loss / validation code reads
(``names``, ``hyp``, ``nc``, ``stride``, ``Detect.anchors``).  The reference object is not modified and nothing of
it is called at run time.
"""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn

from . import yolo, yolo_snn

_CARRY = ("names", "hyp", "gr", "class_weights", "nc", "inplace")


def _is_stack_b(cfg: dict) -> bool:
    return any(row[2] == "DDetect" for row in cfg["head"])


def unsupported_modules(cfg: dict):
    """Module names of a model YAML that this package has no drop-in class for."""
    from . import common
    ns = set(dir(common)) | {"Detect", "DDetect"}
    return sorted({row[2] for row in cfg["backbone"] + cfg["head"] if row[2] not in ns and not row[2].startswith("nn.")})


def convert(model: nn.Module, device=None, strict: bool = True) -> nn.Module:
    """Returns this package's Model / DetectionModel with the reference model's plan, weights and attributes.

    model: a reference ``models.yolo.Model`` / ``models.yolo_snn.DetectionModel`` (or a DDP / EMA wrapper of one:
    ``.module`` / ``.ema`` are unwrapped).  device: where to build the result (default: cuda:0 -- the forward has
    no CPU path; pass 'cpu' to convert on a host and ``.cuda()`` later).
    """
    for attr in ("module", "ema"):
        if hasattr(model, attr) and isinstance(getattr(model, attr), nn.Module) and not hasattr(model, "yaml"):
            model = getattr(model, attr)
    cfg = getattr(model, "yaml", None)
    if not isinstance(cfg, dict):
        raise TypeError("convert(): the model carries no `.yaml` plan (not an ECS-YOLO Model / DetectionModel)")
    missing = unsupported_modules(cfg)
    if missing:
        raise NotImplementedError(f"convert(): no B200 drop-in for module type(s) {missing}")
    cls = yolo_snn.DetectionModel if _is_stack_b(cfg) else yolo.Model
    sd = model.state_dict()
    ours = cls(deepcopy(cfg), ch=cfg.get("ch", 3))
    # Model.__init__ rescales the YAML anchors by the strides; the checkpoint's buffers are already rescaled
    missing_k, unexpected_k = ours.load_state_dict(sd, strict=False)
    if strict and (missing_k or unexpected_k):
        raise RuntimeError(f"convert(): state_dict mismatch: missing {missing_k[:5]}, unexpected {unexpected_k[:5]}")
    for a in _CARRY:
        if hasattr(model, a):
            setattr(ours, a, deepcopy(getattr(model, a)))
    if hasattr(model, "stride") and torch.is_tensor(model.stride):
        head = ours.model[-1]
        if tuple(head.stride.tolist()) != tuple(float(v) for v in model.stride.tolist()):
            raise RuntimeError("convert(): stride mismatch between the reference model and the rebuilt plan")
    for mo, mr in zip(ours.modules(), model.modules()):   # identical module trees: BN momentum / eps follow
        if isinstance(mo, nn.BatchNorm3d) and isinstance(mr, nn.BatchNorm3d):
            mo.momentum, mo.eps = mr.momentum, mr.eps
    ours.train(model.training)
    if device is None:
        device = "cuda:0" if torch.cuda.is_available() else "cpu"
    return ours.to(device)
