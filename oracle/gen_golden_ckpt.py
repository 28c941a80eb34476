"""Pickled reference checkpoints for the checkpoint-compatibility test (container only):
    python oracle/gen_golden_ckpt.py
Builds two micro plans (tests/golden/seeded.py CKPT_PLANS: Stack A with Detect, Stack B with DDetect) with the UNMODIFIED
reference classes, fills them with seeded weights and saves them the way train.py:659-669 does (`{'model': ..., 'ema': ...}`, the
whole nn.Module pickled, half precision as strip_optimizer leaves it).  The files hold tensors and CLASS REFERENCES
(`models.yolo.Model`, `models.common.*`), no reference code.  -> tests/golden/ckpt_micro_{a,b}.pt + meta."""
import os
import sys
from copy import deepcopy

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402


def main():
    C, Y, B = ref_shim.load(4)
    for name, cls in (("micro_a", Y.Model), ("micro_b", B.DetectionModel)):
        torch.manual_seed(0)
        m = cls(deepcopy(S.CKPT_PLANS[name]))
        sd = S.reseed_state_dict(m.state_dict(), 1201)
        for k in sd:                       # anchors / strides stay as built
            if k.endswith("anchors"):
                sd[k] = m.state_dict()[k].clone()
        m.load_state_dict(sd)
        m.names = [f"cls{i}" for i in range(m.yaml["nc"])]
        m.hyp = {"box": 0.05, "cls": 0.5}
        m.nc = m.yaml["nc"]
        ema = deepcopy(m).eval().half() if name == "micro_a" else None       # micro_b: the `ckpt['ema'] or ckpt['model']` fallback
        ckpt = {"epoch": 3, "best_fitness": 0.1, "model": deepcopy(m).half(), "ema": ema, "updates": 7,
                "optimizer": None, "wandb_id": None, "date": "2026-01-01"}
        path = os.path.join(S.GOLDEN_DIR, f"ckpt_{name}.pt")
        torch.save(ckpt, path)
        chk = S.sd_checksum({k: v.half().float() for k, v in m.state_dict().items() if v.is_floating_point()})
        torch.save({"chk": chk, "keys": list(m.state_dict().keys()), "stride": m.stride.clone(),
                    "types": [type(x).__name__ for x in m.model]}, os.path.join(S.GOLDEN_DIR, f"ckpt_{name}_meta.pt"))
        print(name, os.path.getsize(path), chk)


if __name__ == "__main__":
    main()
