"""Whole-model training-loss fixtures from the UNMODIFIED reference (container only):
    python oracle/gen_golden_model_loss.py
For the tiny plans (Stack A: tiny, tiny_ee; Stack B: tiny_b) with the seeded weights / images of the MODEL_CASES:
train-mode forward with autograd -> the reference's own ComputeLoss (utils/loss.py / utils/loss_tal.py) on seeded
labels -> backward.  Stored: loss, loss_items, the gradients of the detection head's parameters and the norm of every
parameter gradient.  -> tests/golden/model_loss.pt"""
import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ecs_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402


def main():
    res = {}
    cases = [(n, s, "A") for n, s in S.MODEL_CASES.items()] + [(n, s, "B") for n, s in S.MODEL_B_CASES.items()]
    for name, spec, stack in cases:
        C, Y, SN = ref_shim.load(spec["T"])
        path = os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")
        cfg = yaml.safe_load(open(path))
        inp = S.model_inputs(spec, O, cfg)
        m = (Y.Model if stack == "A" else SN.DetectionModel)(path)
        m.load_state_dict(inp["sd"])
        m.hyp = dict(S.MODEL_LOSS_HYP)
        m.train()
        if stack == "A":
            from utils.loss import ComputeLoss
        else:
            from utils.loss_tal import ComputeLoss
        crit = ComputeLoss(m)
        tg = S.model_targets(spec, cfg["nc"])
        out = m(inp["x"])
        loss, items = crit(out, tg)
        loss.sum().backward()
        head = f"model.{len(m.model) - 1}."
        res[name] = dict(loss=loss.detach().reshape(-1).clone(), items=items.clone(),
                         head_grads={k: p.grad.clone() for k, p in m.named_parameters()
                                     if k.startswith(head) and p.grad is not None
                                     and (stack == "A" or k.endswith((".2.weight", ".2.bias")))},   # B: the last 1x1 convs
                         grad_norms={k: float(p.grad.norm()) for k, p in m.named_parameters() if p.grad is not None})
        print(name, float(loss.sum()), items.tolist(), len(res[name]["head_grads"]))
    torch.save(res, os.path.join(S.GOLDEN_DIR, "model_loss.pt"))


if __name__ == "__main__":
    main()
