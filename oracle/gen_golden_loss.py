"""Golden fixtures for SURVEY section 8f rank 1 from the UNMODIFIED reference (container only):
    python oracle/gen_golden_loss.py
utils.loss.ComputeLoss (SIoU box term, BCE objectness / class terms, build_targets) on seeded Detect-shaped raw
outputs: loss, loss_items, the gradient w.r.t. every level's raw output, per-level match counts.
Outputs go to tests/golden/post_loss.pt together with the inputs' checksum."""
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
    """What ComputeLoss.__init__ reads off a model (utils/loss.py:131-160): a parameter (device), .hyp, .model[-1]."""

    def __init__(self, det, hyp):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.hyp = hyp
        self.model = [det]


def main():
    ref_shim.load(4)
    from utils.loss import ComputeLoss
    res = {}
    for name, spec in S.LOSS_CASES.items():
        inp = S.loss_inputs(spec)
        anchors = inp["anchors"]
        det = types.SimpleNamespace(na=anchors.shape[1], nc=spec["nc"], nl=anchors.shape[0], anchors=anchors,
                                    stride=torch.tensor([16.0, 32.0, 64.0][:anchors.shape[0]]))
        crit = ComputeLoss(_Holder(det, dict(spec["hyp"])))
        for extra in range(spec.get("calls", 1) - 1):          # stateful criteria (SlideLoss EMA): earlier calls first
            pre = S.loss_inputs(dict(spec, seed=spec["seed"] + 50 + extra))
            crit([x.clone().requires_grad_(True) for x in pre["p"]], pre["targets"].clone())
        p = [x.clone().requires_grad_(True) for x in inp["p"]]
        loss, items = crit(p, inp["targets"].clone())
        (loss * inp["gout"]).sum().backward()
        _, _, indices, _ = crit.build_targets(p, inp["targets"].clone())
        res[name] = dict(chk=S.checksum(*inp["p"], inp["targets"]), loss=loss.detach().clone(), items=items.clone(),
                         grads=[x.grad.clone() for x in p], n=[int(ix[0].shape[0]) for ix in indices])
        print(name, float(loss), items.tolist(), res[name]["n"])
    torch.save(res, os.path.join(S.GOLDEN_DIR, "post_loss.pt"))


if __name__ == "__main__":
    main()
