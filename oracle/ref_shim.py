"""Import shim for the UNMODIFIED reference (test infrastructure only; container-only).

Used by oracle/gen_golden.py to produce tests/golden/ fixtures and to pin the oracle port.
/root/reference does not exist on the GPU box: nothing at run time may import this module there.
Recipe follows SURVEY.md Appendix A.
"""
import sys
import types
import logging
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return MagicMock(name=f"{self.__name__}.{n}")


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            sys.modules[n] = m


_loaded = None


def load(T=4):
    """Returns (common, yolo, yolo_snn) reference modules with time_window = T."""
    global _loaded
    if _loaded is None:
        for n in ["visualizer", "spikingjelly.activation_based.layer", "spikingjelly.activation_based.neuron",
                  "spikingjelly.activation_based.functional", "spikingjelly.activation_based.surrogate",
                  "matplotlib.pyplot", "matplotlib.font_manager", "seaborn", "IPython", "git",
                  "timm.models.layers", "albumentations", "thop"]:
            try:
                __import__(n)
            except Exception:
                _stub(n)
        sys.modules["visualizer"].get_local = lambda *a, **k: (lambda f: f)
        from PIL import ImageFont as _IF
        _tt = _IF.truetype

        def _safe_tt(font=None, size=10, *a, **k):
            try:
                return _tt(font, size, *a, **k)
            except Exception:
                return _IF.load_default()
        _IF.truetype = _safe_tt
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        logging.disable(logging.INFO)
        import models.common as C
        import models.yolo as Y
        try:
            import models.yolo_snn as S
        except Exception as e:  # pragma: no cover
            S = None
            print("yolo_snn import failed:", e)
        _loaded = (C, Y, S)
    C, Y, S = _loaded
    C.time_window = T
    Y.time_window = T
    if S is not None:
        S.time_window = T
    return C, Y, S
