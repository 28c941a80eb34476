"""Golden fixtures for the section-8f rows from the UNMODIFIED reference (container only):
    python oracle/gen_golden_post.py
utils.general.non_max_suppression on seeded Detect-shaped predictions; torch.optim.SGD + utils.torch_utils.ModelEMA
on a seeded parameter set.  Outputs go to tests/golden/post_*.pt together with the inputs' checksum."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402


def main():
    ref_shim.load(4)
    from utils.general import non_max_suppression
    res = {}
    for name, spec in S.NMS_CASES.items():
        pred = S.nms_inputs(spec)
        out = non_max_suppression(pred.clone(), spec["conf"], spec["iou"], classes=spec.get("classes"),
                                  agnostic=spec.get("agnostic", False), multi_label=spec.get("multi_label", False),
                                  max_det=spec.get("max_det", 300))
        res[name] = dict(spec=spec, chk=S.checksum(pred), out=[o.clone() for o in out])
        print(name, [int(o.shape[0]) for o in out])
    torch.save(res, os.path.join(S.GOLDEN_DIR, "post_nms.pt"))

    from utils.torch_utils import ModelEMA
    spec = S.OPT_CASE
    model, grads = S.opt_inputs(spec)
    g0, g1, g2 = S.opt_groups(model)
    opt = torch.optim.SGD(g0, lr=spec["lr"], momentum=spec["momentum"], nesterov=True)        # train.py:282
    opt.add_param_group({"params": g1, "weight_decay": spec["weight_decay"]})                 # train.py:285
    opt.add_param_group({"params": g2})                                                       # train.py:287
    ema = ModelEMA(model)
    states = []
    for step, gs in enumerate(grads):
        for (n_, p), g in zip(model.named_parameters(), gs):
            p.grad = g.clone()
        for j, pg in enumerate(opt.param_groups):                                             # warm-up lr (train.py:520-529)
            pg["lr"] = spec["lr"] * (1.0 + 0.1 * step) * (1.5 if j == 2 else 1.0)
        with torch.no_grad():
            model.bn.running_mean.add_(0.01 * (step + 1))                                     # buffers move too
        opt.step()
        opt.zero_grad()
        ema.update(model)
        states.append(dict(model={k: v.clone() for k, v in model.state_dict().items()},
                           ema={k: v.clone() for k, v in ema.ema.state_dict().items()},
                           updates=ema.updates))
    torch.save(dict(spec=spec, states=states), os.path.join(S.GOLDEN_DIR, "post_opt.pt"))
    print("opt steps", len(states))


if __name__ == "__main__":
    main()
