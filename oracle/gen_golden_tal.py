"""Golden fixtures for SURVEY section 8f rank 1 (Stack B) from the UNMODIFIED reference (container only):
    python oracle/gen_golden_tal.py
utils.loss_tal.ComputeLoss (TaskAlignedAssigner, SIoU box term, DFL, BCE class term) on seeded DDetect-shaped raw
outputs: loss, loss_items, the gradient w.r.t. every level, the number of foreground anchors.
Outputs go to tests/golden/post_tal.pt together with the inputs' checksum."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402


class _Holder(torch.nn.Module):
    def __init__(self, det, hyp):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.hyp = hyp
        self.model = [det]


def main():
    ref_shim.load(4)
    from utils.loss_tal import ComputeLoss
    res = {}
    for name, spec in S.TAL_CASES.items():
        inp = S.tal_inputs(spec)
        nl = len(spec["grids"])
        det = types.SimpleNamespace(nl=nl, nc=spec["nc"], no=64 + spec["nc"], reg_max=16, stride=inp["strides"])
        hyp = dict(cls_pw=spec.get("cls_pw", 1.0), fl_gamma=spec.get("fl_gamma", 0.0), label_smoothing=spec.get("smooth", 0.0))
        for k, v in zip(("YOLOM", "YOLOA", "YOLOB"), spec.get("assigner", (None, None, None))):
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = str(v)                  # read by ComputeLoss.__init__ (utils/loss_tal.py:134-137)
        crit = ComputeLoss(_Holder(det, hyp))
        fg = {}
        orig = crit.assigner.forward

        def spy(*a, _orig=orig, **k):
            out = _orig(*a, **k)
            fg["n"] = int(out[3].sum())
            fg["score_sum"] = float(out[2].sum())
            return out
        crit.assigner.forward = spy
        feats = [x.clone().requires_grad_(True) for x in inp["feats"]]
        loss, items = crit(feats, inp["targets"].clone())
        (loss * inp["gout"]).sum().backward()
        res[name] = dict(chk=S.checksum(*inp["feats"], inp["targets"]), loss=loss.detach().clone(), items=items.clone(),
                         grads=[x.grad.clone() for x in feats], fg=fg["n"], score_sum=fg["score_sum"])
        print(name, float(loss), items.tolist(), fg)
    torch.save(res, os.path.join(S.GOLDEN_DIR, "post_tal.pt"))


if __name__ == "__main__":
    main()
