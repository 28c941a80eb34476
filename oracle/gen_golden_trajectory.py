"""Training-curve fixtures from the UNMODIFIED reference (container only; SURVEY 8c tier 4):
    python oracle/gen_golden_trajectory.py
For the tiny plans (Stack A: tiny; Stack B: tiny_b) with the seeded weights / images / labels of the whole-model loss
cases: K optimizer steps of the reference's own training recipe on one fixed batch -- train-mode forward, the
reference's ComputeLoss (utils/loss.py / utils/loss_tal.py), autograd, torch.optim.SGD(nesterov) over the three parameter
groups of train.py:262-287 (BatchNorm3d weights / other weights with decay / biases).  Stored: the loss and loss_items of
every step and, after the last step, the head's parameters and the norm of every parameter's total displacement.
-> tests/golden/train_trajectory.pt"""
import os
import sys

import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ecs_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402


def reference_groups(model):
    """train.py:262-277, statement for statement."""
    g0, g1, g2 = [], [], []
    for v in model.modules():
        if hasattr(v, 'bias') and isinstance(v.bias, nn.Parameter):
            g2.append(v.bias)
        if isinstance(v, (nn.BatchNorm3d)):
            g0.append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, nn.Parameter):
            g1.append(v.weight)
    return g0, g1, g2


def main():
    res = {}
    cases = [("tiny_64", S.MODEL_CASES["tiny_64"], "A"), ("tiny_b_64", S.MODEL_B_CASES["tiny_b_64"], "B")]
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
        hp = S.TRAJECTORY_HYP
        g0, g1, g2 = reference_groups(m)
        opt = torch.optim.SGD(g0, lr=hp["lr"], momentum=hp["momentum"], nesterov=True)      # train.py:283
        opt.add_param_group({'params': g1, 'weight_decay': hp["weight_decay"]})              # train.py:286
        opt.add_param_group({'params': g2})                                                  # train.py:288
        start = {k: p.detach().clone() for k, p in m.named_parameters()}
        losses, items_all = [], []
        for _ in range(hp["steps"]):
            opt.zero_grad()
            loss, items = crit(m(inp["x"]), tg)
            loss.sum().backward()
            opt.step()
            losses.append(float(loss.sum()))
            items_all.append(items.detach().clone())
        head = f"model.{len(m.model) - 1}."
        res[name] = dict(losses=torch.tensor(losses, dtype=torch.float64), items=torch.stack(items_all),
                         head_params={k: p.detach().clone() for k, p in m.named_parameters() if k.startswith(head)
                                      and (stack == "A" or k.endswith((".2.weight", ".2.bias")))},   # B: the last 1x1 convs
                         moved={k: float((p.detach() - start[k]).norm()) for k, p in m.named_parameters()},
                         bn_running={k: v.clone() for k, v in m.state_dict().items() if k.endswith("running_mean")})
        print(name, [round(v, 6) for v in losses])
    torch.save(res, os.path.join(S.GOLDEN_DIR, "train_trajectory.pt"))


if __name__ == "__main__":
    main()
