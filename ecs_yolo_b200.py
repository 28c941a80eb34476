"""Alias so the hyphen-named package can be imported with a plain statement:
``import ecs_yolo_b200 as ecsy`` (then use attribute access: ``ecsy.yolo.Model``)."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_sys.modules[__name__] = _il.import_module("ecs-yolo_b200")
