"""Golden fixtures for the Gen1 event -> frame input path from the UNMODIFIED reference (container only):
    python oracle/gen_golden_events.py
g1-resnet/utils/give_g1_data.py `LoadImagesAndLabels.create_data` (:550-565) paints the T event bins, then the loader
resizes every frame with cv2.resize to the network size (g1-resnet/utils/datasets_g1T.py:518-533) and the training
loop divides by 255 (g1-resnet/train_g1.py:298).  Runs in its own process (its stubs differ from oracle/ref_shim.py)."""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import seeded as S  # noqa: E402

G1 = "/root/reference/g1-resnet"


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            m.__getattr__ = lambda a, n=n: MagicMock(name=f"{n}.{a}")
            sys.modules[n] = m


def main():
    import cv2
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.load(5)                      # stubs + the root reference packages (`utils`, `models`)
    for n in ["turtle", "prophesee_utils.io.psee_loader"]:
        try:
            __import__(n)
        except Exception:
            _stub(n)
    # g1-resnet/utils holds only the files that differ from the root `utils` package: import the root package and load
    # give_g1_data.py from its own path
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("utils.give_g1_data", os.path.join(G1, "utils", "give_g1_data.py"))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    LoadImagesAndLabels = mod.LoadImagesAndLabels
    res = {}
    for name, spec in S.EVENT_CASES.items():
        samples = S.event_inputs(spec)
        frames = []
        for bins in samples:
            ev = []
            for b in bins:
                a = np.zeros(len(b["x"]), dtype=[("x", "<i8"), ("y", "<i8"), ("p", "<i8")])
                a["x"], a["y"], a["p"] = b["x"].numpy(), b["y"].numpy(), b["p"].numpy()
                ev.append(a)
            img = LoadImagesAndLabels.create_data(types.SimpleNamespace(T=spec["T"]), ev)      # [T, 240, 304, 3] uint8
            out = np.zeros([spec["T"], spec["out"], spec["out"], 3])
            for i in range(spec["T"]):
                out[i] = cv2.resize(img[i], (spec["out"], spec["out"]))                        # datasets_g1T.py:527-530
            frames.append((img, np.transpose(out, [0, 3, 1, 2])))
        painted = torch.from_numpy(np.stack([f[0] for f in frames]))                           # [N, T, 240, 304, 3]
        resized = torch.from_numpy(np.stack([f[1] for f in frames])).to(torch.uint8)           # [N, T, 3, S, S]
        res[name] = dict(spec=spec, painted_ch0=S.zpack(painted[..., 0].contiguous()),
                         same_channels=bool((painted[..., 0] == painted[..., 1]).all() and (painted[..., 0] == painted[..., 2]).all()),
                         resized_ch0=S.zpack(resized[:, :, 0].contiguous()),
                         resized_same=bool((resized[:, :, 0] == resized[:, :, 1]).all()))
        print(name, painted.shape, resized.shape, res[name]["same_channels"], res[name]["resized_same"])
    torch.save(res, os.path.join(S.GOLDEN_DIR, "post_events.pt"))


if __name__ == "__main__":
    main()
