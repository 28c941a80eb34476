"""The section-8f oracles (oracle/post_oracle.py) against fixtures made by the UNMODIFIED reference functions
(oracle/gen_golden_post.py): utils.general.non_max_suppression, torch.optim.SGD + utils.torch_utils.ModelEMA."""
import copy
import os

import pytest
import torch

import post_oracle as P
import seeded as S


def _load(name):
    return torch.load(os.path.join(S.GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", list(S.NMS_CASES))
def test_nms_oracle(name):
    gold = _load("post_nms")[name]
    spec = S.NMS_CASES[name]
    pred = S.nms_inputs(spec)
    assert abs(S.checksum(pred) - gold["chk"]) <= 1e-6 * abs(gold["chk"])
    out = P.non_max_suppression(pred.clone(), spec["conf"], spec["iou"], spec.get("classes"), spec.get("agnostic", False),
                                spec.get("multi_label", False), spec.get("max_det", 300))
    assert len(out) == len(gold["out"])
    for a, b in zip(out, gold["out"]):
        assert a.shape == b.shape
        assert torch.equal(a, b)          # bit-exact: same candidates, same order, same arithmetic


def test_sgd_ema_oracle():
    gold = _load("post_opt")
    spec = S.OPT_CASE
    model, grads = S.opt_inputs(spec)
    groups = S.opt_groups(model)
    gid = {id(p): j for j, g in enumerate(groups) for p in g}
    params = dict(model.named_parameters())
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ema = copy.deepcopy(state)
    bufs = {k: None for k in params}
    for step, gs in enumerate(grads):
        state["bn.running_mean"] += 0.01 * (step + 1)
        for (k, p), g in zip(params.items(), gs):
            j = gid[id(p)]
            lr = spec["lr"] * (1.0 + 0.1 * step) * (1.5 if j == 2 else 1.0)
            wd = spec["weight_decay"] if j == 1 else 0.0
            state[k], bufs[k] = P.sgd_nesterov_step(state[k], g, bufs[k], lr, spec["momentum"], wd)
        P.ema_update(ema, state, P.ema_decay(step + 1))
        want = gold["states"][step]
        for k in state:
            assert torch.allclose(state[k].float(), want["model"][k].float(), rtol=1e-6, atol=1e-8), (step, k)
            assert torch.allclose(ema[k].float(), want["ema"][k].float(), rtol=1e-6, atol=1e-8), (step, k)


@pytest.mark.parametrize("name", list(S.EVENT_CASES))
def test_event_frames_oracle(name):
    """create_data (last event of a pixel wins) and the cv2.resize restatement, bit-exact on the reference's frames."""
    gold = _load("post_events")[name]
    spec = S.EVENT_CASES[name]
    samples = S.event_inputs(spec)
    assert gold["same_channels"] and gold["resized_same"]
    painted = P.events_to_frames(samples, spec["T"])
    assert torch.equal(painted, S.zunpack(gold["painted_ch0"]))
    want = S.zunpack(gold["resized_ch0"])                       # [N, T, S, S]
    for n in range(spec["N"]):
        for t in range(0, spec["T"], 2):
            assert torch.equal(P.resize_linear_u8(painted[n, t], spec["out"], spec["out"]), want[n, t]), (n, t)
    x = P.event_frames(samples[:1], spec["T"], spec["out"])
    assert x.shape == (spec["T"], 1, 3, spec["out"], spec["out"])
    assert torch.equal(x[:, 0, 1], want[0].float() / 255)


@pytest.mark.parametrize("name", list(S.LOSS_CASES))
def test_compute_loss_oracle(name):
    """oracle/loss_oracle.py against utils.loss.ComputeLoss of the unmodified reference: loss, loss_items, match counts and
    the gradient w.r.t. every level's raw head output."""
    import loss_oracle as LO
    gold = _load("post_loss")[name]
    spec = S.LOSS_CASES[name]
    inp = S.loss_inputs(spec)
    assert abs(S.checksum(*inp["p"], inp["targets"]) - gold["chk"]) <= 1e-6 * abs(gold["chk"])
    state = {}
    for extra in range(spec.get("calls", 1) - 1):              # SlideLoss keeps an EMA across calls
        pre = S.loss_inputs(dict(spec, seed=spec["seed"] + 50 + extra))
        LO.compute_loss(pre["p"], pre["targets"], pre["anchors"], spec["hyp"], slide_state=state)
    p = [x.clone().requires_grad_(True) for x in inp["p"]]
    loss, items, counts, _ = LO.compute_loss(p, inp["targets"], inp["anchors"], spec["hyp"], slide_state=state)
    assert counts == gold["n"]
    assert torch.allclose(loss, gold["loss"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(items, gold["items"], rtol=1e-6, atol=1e-7)
    (loss * inp["gout"]).sum().backward()
    for x, g in zip(p, gold["grads"]):
        assert torch.allclose(x.grad, g, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", list(S.TAL_CASES))
def test_tal_loss_oracle(name):
    """oracle/tal_oracle.py against utils.loss_tal.ComputeLoss (TaskAlignedAssigner + SIoU + DFL + BCE) of the unmodified
    reference: loss, loss_items, number of foreground anchors, gradient w.r.t. every level's raw head output."""
    import tal_oracle as TO
    gold = _load("post_tal")[name]
    spec = S.TAL_CASES[name]
    inp = S.tal_inputs(spec)
    assert abs(S.checksum(*inp["feats"], inp["targets"]) - gold["chk"]) <= 1e-6 * abs(gold["chk"])
    feats = [x.clone().requires_grad_(True) for x in inp["feats"]]
    loss, items, n_fg = TO.compute_loss(feats, inp["targets"], spec["strides"], spec.get("cls_pw", 1.0),
                                        spec.get("fl_gamma", 0.0), spec.get("assigner", (10, 0.5, 6.0)))
    assert n_fg == gold["fg"]
    assert torch.allclose(loss, gold["loss"], rtol=2e-6, atol=1e-6), (float(loss), float(gold["loss"]))
    assert torch.allclose(items, gold["items"], rtol=2e-6, atol=1e-6)
    (loss * inp["gout"]).sum().backward()
    for x, g in zip(feats, gold["grads"]):
        assert torch.allclose(x.grad, g, rtol=1e-4, atol=1e-7 * float(g.abs().max()))


@pytest.mark.parametrize("name", ["tiny_64", "tiny_ee_64"])
def test_model_plus_loss_oracle(name):
    """The two oracles chained -- ecs_oracle.forward (train mode) -> loss_oracle.compute_loss -- against the unmodified
    reference's model + utils.loss.ComputeLoss on the same weights, image and labels (oracle/gen_golden_model_loss.py)."""
    import yaml
    import ecs_oracle as O
    import loss_oracle as LO
    from util import ROOT
    spec = S.MODEL_CASES[name]
    gold = _load("model_loss")[name]
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    with torch.no_grad():
        out = O.forward(cfg, inp["sd"], inp["x"], spec["T"], True, stride=inp["stride"])
    anchors = [v for k, v in inp["sd"].items() if k.endswith("anchors")][0]
    loss, items, counts, _ = LO.compute_loss(out, S.model_targets(spec, cfg["nc"]), anchors, S.MODEL_LOSS_HYP)
    assert min(counts) > 0
    assert torch.allclose(loss.reshape(-1), gold["loss"], rtol=1e-5), (float(loss), float(gold["loss"]))
    assert torch.allclose(items, gold["items"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["tiny_64", "tiny_b_64"])
def test_training_trajectory_oracle(name):
    """SURVEY 8c tier 4 on the CPU: ecs_oracle.forward (train mode, autograd = surrogate-gradient BPTT) ->
    loss_oracle.compute_loss (Stack A) / tal_oracle.compute_loss (Stack B) -> torch.optim.SGD over the reference's three
    parameter groups follows the UNMODIFIED reference's six-step training curve on the tiny plans
    (oracle/gen_golden_trajectory.py)."""
    import yaml
    import ecs_oracle as O
    import loss_oracle as LO
    import tal_oracle as TO
    from util import ROOT
    stack_b = name in S.MODEL_B_CASES
    spec, hp = (S.MODEL_B_CASES if stack_b else S.MODEL_CASES)[name], S.TRAJECTORY_HYP
    gold = _load("train_trajectory")[name]
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    sd = {k: v.clone() for k, v in inp["sd"].items()}
    buffers = ("running_mean", "running_var", "num_batches_tracked", "anchors")
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and not k.endswith(buffers)}
    assert set(params) == set(gold["moved"]), set(params) ^ set(gold["moved"])
    g0 = [v for k, v in params.items() if k.endswith(".bn.weight")]
    g2 = [v for k, v in params.items() if k.endswith(".bias")]
    g1 = [v for k, v in params.items() if not k.endswith((".bn.weight", ".bias"))]
    opt = torch.optim.SGD(g0, lr=hp["lr"], momentum=hp["momentum"], nesterov=True)
    opt.add_param_group({'params': g1, 'weight_decay': hp["weight_decay"]})
    opt.add_param_group({'params': g2})
    tg = S.model_targets(spec, cfg["nc"])
    anchors = None if stack_b else [v for k, v in sd.items() if k.endswith("anchors")][0]
    losses = []
    for _ in range(hp["steps"]):
        opt.zero_grad()
        out = O.forward(cfg, sd, inp["x"], spec["T"], True, stride=inp["stride"])
        if stack_b:
            loss = TO.compute_loss(out, tg, inp["stride"])[0]
        else:
            loss = LO.compute_loss(out, tg, anchors, S.MODEL_LOSS_HYP)[0]
        loss.sum().backward()
        opt.step()
        losses.append(float(loss.detach().sum()))
    want = gold["losses"].tolist()
    dev = [abs(a - b) / abs(b) for a, b in zip(losses, want)]
    err = {k: float((sd[k].detach() - v).norm() / v.norm()) for k, v in gold["head_params"].items()}
    print("trajectory rel dev", dev, "head param rel-L2", err)
    # measured here: tiny_64 steps 0-4 bit-identical, step 5 off by 1.5 % -- one near-threshold spike flips between two CPU
    # evaluation orders of the same arithmetic (the reference's own PyTorch-vs-PyTorch noise floor, SURVEY 8c tier 3);
    # tiny_b_64 (Stack B: DDetect + TAL loss) follows all six steps to 1e-7
    assert max(dev[:5]) <= 2e-4 and max(dev) <= 3e-2, (losses, want)
    assert max(err.values()) < 2e-2, err
