"""Backward parity of the drop-in blocks: gradients from the manual C-ABI backward chains vs the
reference's autograd (golden gx / parameter gradients recorded from the unmodified reference)."""
import os

import pytest
import torch
import yaml

import ecs_oracle as O
import seeded as S
from util import ROOT, ecsy, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _build_block(E, spec):
    cls = getattr(E.common, spec["kind"])
    if spec["kind"] == "BasicBlock_1":
        return cls(spec["cin"], spec["cout"], spec["s"])
    return cls(spec["cin"], spec["cout"], spec["k"], spec["s"])


ALL_BLOCKS = {**S.BLOCK_CASES, **S.MS_BLOCK_CASES}   # MS_*: res*-ee.yaml blocks incl. the zero-padded 3 / 32-channel front


@pytest.mark.parametrize("mode,tol_out,tol_gx,tol_p", [("parity", 1e-3, 5e-3, 1e-2), ("fast", 0.2, 0.15, 0.35)])
@pytest.mark.parametrize("name", list(ALL_BLOCKS))
def test_block_backward(name, mode, tol_out, tol_gx, tol_p):
    """Block forward + BPTT against the reference's autograd, in both precisions.  Not teacher-forced inside the block:
    in fast precision (one bf16 plane for weights and gradient operands) a flipped near-threshold spike also moves
    the surrogate window of the elements it feeds, hence the wider fast-mode bounds (measured on B200: output 6-12 %, input
    gradient 5-10 %, worst parameter gradient 11-23 % rel-L2 on these fixtures, whose spread weights are 1.5x the default
    initialisation; parity precision: 3e-6 / 1e-7 / 1e-5); test_block_forward_fast holds the same blocks' forward to the
    north-star tolerances with teacher-forced neurons, test_lif_ecs_bwd the neuron's BPTT alone (1-2.4 %)."""
    E = ecsy()
    E.set_precision(mode)
    try:
        _block_backward(E, name, tol_out, tol_gx, tol_p)
    finally:
        E.set_precision("parity")


def _block_backward(E, name, tol_out, tol_gx, tol_p):
    spec, gold = ALL_BLOCKS[name], load_golden(name)
    inp = S.block_inputs(spec, O)
    m = _build_block(E, spec)
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda().train()
    x = inp["x"].cuda().requires_grad_(True)
    out = m(x)
    e_out = rel_l2(out.detach().cpu(), gold["out_train"])
    gout = S.randn(S.gen(spec["seed"] + 13), *out.shape).cuda()
    out.backward(gout)
    # teacher-forced bound: a near-threshold flip moves the surrogate window of a few elements
    e_gx = rel_l2(x.grad.cpu(), gold["gx"])
    assert e_out < tol_out, e_out
    assert e_gx < tol_gx, f"{name}: gx {e_gx:.3e}"
    named = dict(m.named_parameters())
    worst = ("", 0.0)
    for k, g in gold["grads"].items():
        p = named[k[len("model.0."):]]
        assert p.grad is not None, k
        if isinstance(g, dict):
            got, want = p.grad.cpu().flatten()[::97], g["sample"]
        else:
            got, want = p.grad.cpu(), g
        e = rel_l2(got, want)
        if e > worst[1]:
            worst = (k, e)
    print(f"\n{name}: out {e_out:.2e} gx {e_gx:.2e} worst parameter gradient {worst[0]} {worst[1]:.2e}")
    assert worst[1] < tol_p, f"{name}: worst parameter gradient {worst[0]} rel-L2 {worst[1]:.3e}"


def test_model_training_step():
    """Whole Stack-A model: forward + backward + SGD step runs, every parameter that the reference trains gets
    a finite gradient, and the loss decreases over a few steps on a fixed batch."""
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    tgt = [torch.randn(2, 3, 8, 8, 8, device="cuda"), torch.randn(2, 3, 4, 4, 8, device="cuda")]
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
    losses = []
    for it in range(4):
        opt.zero_grad(set_to_none=True)
        out = m(x)
        loss = sum(((o - t) ** 2).mean() for o, t in zip(out, tgt))
        loss.backward()
        if it == 0:
            missing = [n for n, p in m.named_parameters() if p.grad is None]
            assert not missing, missing[:5]
            assert all(torch.isfinite(p.grad).all() for p in m.parameters())
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


# ---------------------------------------------------------------- Stack B backward
def test_silu_neuron_backward():
    E = ecsy()
    F = E.functional
    name = "silu_c64_t4_inplace"
    spec, gold = S.SILU_CASES[name], load_golden(name)
    inp = S.lif_inputs(spec)
    w = F.make_lif_w(inp["dw_w"].cuda(), inp["dw_b"].cuda(), inp["pw_w"].cuda(), inp["pw_b"].cuda())
    x = F.Act.from_ref(inp["x"].cuda())
    gout = inp["gout"].cuda().permute(0, 1, 3, 4, 2).contiguous()
    gx, g_dw_w, g_dw_b, g_pw_w, g_pw_b = F.lif_silu_bwd(gout, x, w, inp["pw_w"].cuda())
    assert rel_l2(gx.permute(0, 1, 4, 2, 3).cpu(), gold["gx"]) < 2e-4
    for got, k in [(g_dw_w, "g_dw_w"), (g_dw_b, "g_dw_b"), (g_pw_w, "g_pw_w"), (g_pw_b, "g_pw_b")]:
        e = rel_l2(got.cpu(), gold[k])
        assert e < 2e-4, f"{k}: {e:.3e}"


@pytest.mark.parametrize("name", list(S.CONVSILU_CASES))
def test_conv_silu_backward(name):
    E = ecsy()
    spec, gold = S.CONVSILU_CASES[name], load_golden(name)
    inp = S.convsilu_inputs(spec, O)
    m = E.common.Conv(spec["cin"], spec["cout"], spec["k"], spec["s"])
    m.act.actFun.inplace = True
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda().train()
    x = inp["x"].cuda().requires_grad_(True)
    out = m(x)
    assert rel_l2(out.detach().cpu(), gold["out_train"]) < 5e-5
    out.backward(S.randn(S.gen(spec["seed"] + 13), *out.shape).cuda())
    assert rel_l2(x.grad.cpu(), gold["gx"]) < 5e-4
    named = dict(m.named_parameters())
    for k, g in gold["grads"].items():
        e = rel_l2(named[k[len("model.0."):]].grad.cpu(), g)
        assert e < 1e-3, f"{k}: {e:.3e}"


@pytest.mark.parametrize("name", list(S.DDETECT_CASES))
def test_ddetect_backward(name):
    E = ecsy()
    spec, gold = S.DDETECT_CASES[name], load_golden(name)
    inp = S.ddetect_inputs(spec, O)
    m = E.yolo_snn.DDetect(spec["nc"], spec["ch"])
    m.stride = inp["stride"]
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda().train()
    fs = [f.cuda().requires_grad_(True) for f in inp["feats"]]
    out = m(list(fs))
    for a, b in zip(out, gold["out_train"]):
        assert rel_l2(a.detach().cpu(), b) < 1e-3
    gouts = [S.randn(S.gen(spec["seed"] + 13 + i), *o.shape).cuda() for i, o in enumerate(out)]
    sum((o * g).sum() for o, g in zip(out, gouts)).backward()
    for f, g in zip(fs, gold["gfeats"]):
        assert rel_l2(f.grad.cpu(), g) < 1e-2
    named = dict(m.named_parameters())
    worst = ("", 0.0)
    for k, g in gold["grads"].items():
        p = named[k[len("model.0."):]]
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu(), g)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 2e-2, worst


def test_model_b_training_step():
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo_snn.DetectionModel(E.cfg_path("tiny_b")).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    with torch.no_grad():
        probe = m(x)
    tgt = [torch.randn_like(o) for o in probe]
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
    losses = []
    for it in range(4):
        opt.zero_grad(set_to_none=True)
        out = m(x)
        loss = sum(((o - t) ** 2).mean() for o, t in zip(out, tgt))
        loss.backward()
        if it == 0:
            missing = [n for n, p in m.named_parameters() if p.grad is None and p.requires_grad]
            assert not missing, missing[:5]
            assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


def test_stored_lif_state_equals_recompute():
    """set_lif_store(True) keeps the forward's membranes / traces for the BPTT backward; the gradients must be
    the recompute path's (same kernels, same arithmetic on the same state), in both precision modes."""
    E = ecsy()
    F = E.functional
    spec = S.BLOCK_CASES["bb2_64_128_s2"]
    inp = S.block_inputs(spec, O)
    for mode in ("parity", "fast"):
        F.set_precision(mode)
        try:
            res = []
            for store in (True, False):
                F.set_lif_store(store)
                m = _build_block(E, spec)
                m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
                m = m.cuda().train()
                x = inp["x"].cuda().requires_grad_(True)
                out = m(x)
                out.backward(S.randn(S.gen(5), *out.shape).cuda())
                res.append((out.detach().clone(), x.grad.clone(), [p.grad.clone() for p in m.parameters()]))
            assert torch.equal(res[0][0], res[1][0])            # the forward is deterministic
            # the backward reductions use floating-point atomics (run-to-run order), so gradients agree to rounding
            assert rel_l2(res[0][1], res[1][1]) < 1e-5
            for a, b in zip(res[0][2], res[1][2]):
                assert rel_l2(a, b) < 1e-5
        finally:
            F.set_precision("parity")
            F.set_lif_store(True)


def test_model_ee_training_step():
    """res*-ee.yaml topology (Conv_2 on the image, ConcatBlock_ms, BasicBlock_ms): forward + backward + fused optimizer
    step; every parameter gets a finite gradient and the loss goes down on a fixed batch."""
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo.Model(E.cfg_path("tiny_ee")).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    opt = E.optim.SGDNesterovEMA(m, lr=0.002, momentum=0.9, weight_decay=0.0)
    with torch.no_grad():
        tgt = [torch.randn_like(o) for o in m(x)]
    losses = []
    for _ in range(5):
        opt.zero_grad()
        out = m(x)
        loss = sum(((o - t) ** 2).mean() for o, t in zip(out, tgt))
        loss.backward()
        if not losses:
            missing = [k for k, p in m.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
            assert not missing, missing[:5]
        opt.step()
        losses.append(loss.item())
    assert min(losses[1:]) < losses[0], losses


def _coco_like_targets(n_img, nc, seed):
    g = torch.Generator().manual_seed(seed)
    per = torch.randint(2, 5, (n_img,), generator=g)
    img = torch.repeat_interleave(torch.arange(n_img), per).float()
    n = int(img.numel())
    return torch.cat([img[:, None], torch.randint(0, nc, (n, 1), generator=g).float(),
                      torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.4 + 0.2], 1).cuda()


def test_model_trains_against_compute_loss():
    """Stack-A model end to end with the reference's training loss on the device (ecs.loss.ComputeLoss, utils/loss.py
    semantics): finite gradients everywhere and a decreasing loss on a fixed batch."""
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    det = m.model[-1]
    m.hyp = dict(box=0.05 * 3 / det.nl, cls=0.5 * det.nc / 80 * 3 / det.nl, obj=1.0 * 3 / det.nl, cls_pw=1.0, obj_pw=1.0,
                 anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)
    crit = E.loss.ComputeLoss(m)
    x = torch.rand(2, 3, 64, 64, device="cuda")
    tg = _coco_like_targets(2, det.nc, 5)
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
    losses = []
    for it in range(6):
        opt.zero_grad(set_to_none=True)
        loss, items = crit(m(x), tg)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
        opt.step()
        losses.append(float(loss.detach()))
        assert abs(float(items.sum()) * 2 - losses[-1]) < 1e-4 * losses[-1]     # loss = sum(items) * batch size
    assert min(losses[-2:]) < losses[0], losses


def test_model_b_trains_against_tal_loss():
    """Stack-B model end to end with utils/loss_tal.py semantics on the device (ecs.loss_tal.ComputeLoss)."""
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo_snn.DetectionModel(E.cfg_path("tiny_b")).cuda().train()
    m.hyp = dict(cls_pw=1.0, fl_gamma=0.0, label_smoothing=0.0)
    crit = E.loss_tal.ComputeLoss(m)
    x = torch.rand(2, 3, 64, 64, device="cuda")
    tg = _coco_like_targets(2, m.model[-1].nc, 6)
    # the TAL loss is normalised by the (prediction-dependent) sum of assigned scores and re-assigns every step, so it
    # is not monotone under SGD: small steps, and only "some later step is not above the first" (2 % band: float atomics
    # in the weight gradients make the trajectory of this chaotic miniature vary slightly from run to run) is asserted;
    # exactness of the loss and its gradients is established by test_model_loss_matches_reference_other_plans
    opt = torch.optim.SGD(m.parameters(), lr=2e-4, momentum=0.9)
    losses = []
    for it in range(8):
        opt.zero_grad(set_to_none=True)
        loss, items = crit(m(x), tg)
        loss.backward()
        if it == 0:
            missing = [n for n, p in m.named_parameters() if p.grad is None and p.requires_grad]
            assert not missing, missing[:5]
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
        opt.step()
        losses.append(float(loss.detach()))
        assert abs(float(items.sum()) * 2 - losses[-1]) < 1e-4 * losses[-1]     # loss = sum(items) * batch size
    assert min(losses[1:]) < 1.02 * losses[0], losses


# ---------------------------------------------------------------- whole model + the reference's loss
def _model_loss_case(name, mode="parity"):
    E = ecsy()
    E.set_precision(mode)
    try:
        return _model_loss_case_impl(E, name)
    finally:
        E.set_precision("parity")


def _model_loss_case_impl(E, name):
    stack_b = name in S.MODEL_B_CASES
    spec = (S.MODEL_B_CASES if stack_b else S.MODEL_CASES)[name]
    gold = torch.load(os.path.join(S.GOLDEN_DIR, "model_loss.pt"), weights_only=False)[name]
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    m = (E.yolo_snn.DetectionModel if stack_b else E.yolo.Model)(E.cfg_path(spec["cfg"]))
    m.load_state_dict(inp["sd"])
    m = m.cuda().train()
    m.hyp = dict(S.MODEL_LOSS_HYP)
    crit = (E.loss_tal if stack_b else E.loss).ComputeLoss(m)
    loss, items = crit(m(inp["x"].cuda()), S.model_targets(spec, cfg["nc"]).cuda())
    loss.sum().backward()
    return m, gold, loss.detach().reshape(-1).cpu(), items.cpu()


def test_model_loss_matches_reference():
    """north_star: "head outputs and loss agree within 1e-3 relative in fp32-accumulate mode".  Whole Stack-A model
    (train mode, parity precision) -> ComputeLoss, against the unmodified reference's model + utils.loss.ComputeLoss +
    autograd on the same weights, image and labels (oracle/gen_golden_model_loss.py): loss and loss_items within 1e-3,
    the detection head's parameter gradients within 1e-2 rel-L2, every parameter's gradient norm within 25 % (measured on
    B200: loss 2.7e-6 relative, head gradients <= 1.7e-4, worst gradient-norm deviation 0.33 %)."""
    m, gold, loss, items = _model_loss_case("tiny_64")
    params = dict(m.named_parameters())
    print("tiny_64 loss", float(loss), float(gold["loss"]), "head grad errs",
          [round(rel_l2(params[k].grad.cpu(), g), 6) for k, g in gold["head_grads"].items()], "worst norm ratio",
          max(abs(float(params[k].grad.norm()) / n - 1) for k, n in gold["grad_norms"].items() if n > 1e-6))
    assert torch.allclose(loss, gold["loss"], rtol=1e-3), (float(loss), float(gold["loss"]))
    assert torch.allclose(items, gold["items"], rtol=1e-3, atol=1e-6)
    params = dict(m.named_parameters())
    for k, g in gold["head_grads"].items():
        assert rel_l2(params[k].grad.cpu(), g) < 1e-2, (k, rel_l2(params[k].grad.cpu(), g))
    bad = {k: (float(params[k].grad.norm()), n) for k, n in gold["grad_norms"].items()
           if n > 1e-6 and abs(float(params[k].grad.norm()) - n) > 0.25 * n}
    assert not bad, list(bad.items())[:5]


@pytest.mark.parametrize("name", ["tiny_ee_64", "tiny_b_64"])
def test_model_loss_matches_reference_other_plans(name):
    """Same bar for the res*-ee topology (utils/loss.py) and Stack B (DDetect + utils/loss_tal.py: assigner, box, DFL):
    loss and loss_items within 1e-3 of the unmodified reference (measured on B200: 1.5e-6 / 2.3e-6 relative), the last
    head convolutions' gradients within 2e-2 rel-L2."""
    m, gold, loss, items = _model_loss_case(name)
    print(name, float(loss), float(gold["loss"]), items.tolist(), gold["items"].tolist())
    assert torch.allclose(loss, gold["loss"], rtol=1e-3), (float(loss), float(gold["loss"]))
    assert torch.allclose(items, gold["items"], rtol=1e-3, atol=1e-6)
    params = dict(m.named_parameters())
    errs = {k: rel_l2(params[k].grad.cpu(), g) for k, g in gold["head_grads"].items()}
    print(name, "head grad errs", {k: round(v, 6) for k, v in errs.items()})
    assert max(errs.values()) < 2e-2, errs


def test_fused_optimizer_refreshes_derived_weights():
    """The fused optimizer writes parameters through raw pointers; every derived-weight cache (packed bf16 conv / spread
    weights, dgrad weights, folded tdBN affines, keyed on data_ptr + _version) must see the update.
    (1) One step of SGDNesterovEMA (EMA off) and of torch.optim.SGD configured like the reference's three parameter groups
    (train.py:259-287) from identical weights and gradients give the same parameters.  (2) After two steps the model's
    OUTPUT must equal, bit for bit, the output of a fresh model loaded with its state_dict (no caches at all) and differ
    from the output before training.  With stale caches the forward would still run on the initial packed weights."""
    E = ecsy()
    torch.manual_seed(0)
    ma = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    mb = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    mb.load_state_dict(ma.state_dict())
    det = ma.model[-1]
    hyp = dict(box=0.05 * 3 / det.nl, cls=0.5 * det.nc / 80 * 3 / det.nl, obj=1.0 * 3 / det.nl, cls_pw=1.0, obj_pw=1.0,
               anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)
    ma.hyp = mb.hyp = hyp
    ca, cb = E.loss.ComputeLoss(ma), E.loss.ComputeLoss(mb)
    x = torch.rand(2, 3, 64, 64, device="cuda")
    tg = _coco_like_targets(2, det.nc, 5)
    lr, mom, wd = 0.05, 0.9, 5e-4
    oa = E.optim.SGDNesterovEMA(ma, lr=lr, momentum=mom, weight_decay=wd, ema=False)
    g0, g1, g2 = E.optim.param_groups_of(mb)
    ob = torch.optim.SGD([dict(params=g0, weight_decay=0.0), dict(params=g1, weight_decay=wd),
                          dict(params=g2, weight_decay=0.0)], lr=lr, momentum=mom, nesterov=True)
    with torch.no_grad():
        first = [o.clone() for o in ma(x)]
    for it in range(2):
        for m, crit, opt in ((ma, ca, oa), (mb, cb, ob)):
            opt.zero_grad(set_to_none=True)
            loss, _ = crit(m(x), tg)
            loss.backward()
            opt.step()
        if it == 0:     # identical weights in, gradients equal up to the order of float atomics
            pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
            worst = max(rel_l2(pa[k].detach(), pb[k].detach()) for k in pa)
            assert worst < 1e-4, worst
    sd = {k: v.clone() for k, v in ma.state_dict().items()}
    with torch.no_grad():
        ya = ma(x)
    fresh = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    fresh.load_state_dict(sd)
    with torch.no_grad():
        yf = fresh(x)
    for a, f in zip(ya, yf):
        assert torch.equal(a, f), rel_l2(a, f)
    moved = max(rel_l2(a, f) for a, f in zip(ya, first))
    assert moved > 5e-3, f"the outputs did not move ({moved:.2e}): the forward still runs on the initial weights"


def test_fused_optimizer_ema_model_evaluates_current_average():
    """opt.ema.ema (the ModelEMA copy) is updated by the same launch; its packed weights and folded tdBN affines must be
    rebuilt too.  After enough steps at a short EMA horizon its eval output has to differ from its initial eval output
    and equal the output of a fresh model loaded with the EMA state_dict (no caches at all)."""
    E = ecsy()
    torch.manual_seed(0)
    m = E.yolo.Model(E.cfg_path("tiny")).cuda().train()
    det = m.model[-1]
    m.hyp = dict(box=0.05 * 3 / det.nl, cls=0.5 * det.nc / 80 * 3 / det.nl, obj=1.0 * 3 / det.nl, cls_pw=1.0, obj_pw=1.0,
                 anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)
    crit = E.loss.ComputeLoss(m)
    x = torch.rand(2, 3, 64, 64, device="cuda")
    tg = _coco_like_targets(2, det.nc, 5)
    opt = E.optim.SGDNesterovEMA(m, lr=0.05, momentum=0.9, weight_decay=5e-4, ema=True, ema_decay=0.5)
    opt.ema.decay = lambda n: 0.5
    with torch.no_grad():
        z0 = opt.ema.ema(x)[0].clone()
    for it in range(4):
        opt.zero_grad(set_to_none=True)
        loss, _ = crit(m(x), tg)
        loss.backward()
        opt.step()
    with torch.no_grad():
        z1 = opt.ema.ema(x)[0]
    fresh = E.yolo.Model(E.cfg_path("tiny")).cuda().eval()
    fresh.load_state_dict(opt.ema.ema.state_dict())
    with torch.no_grad():
        z2 = fresh(x)[0]
    assert rel_l2(z1, z2) < 1e-5, rel_l2(z1, z2)
    assert rel_l2(z1, z0) > 1e-3, "the EMA model still evaluates its initial weights"


def test_time_window_mismatch_raises():
    """Conv_7 / Detect built for T = 4 must reject a T = 5 input like the reference's Conv3d(T -> 1) does
    (models/common.py:549-562) instead of reading past the weight buffer."""
    E = ecsy()
    c7 = E.common.Conv_7(1, 1).cuda()
    with pytest.raises(RuntimeError):
        c7(torch.rand(5, 1, 64, 4, 4, device="cuda"))
    det = E.yolo.Detect(3, [[10, 14, 23, 27, 37, 58]], (64,)).cuda().eval()
    det.stride = torch.tensor([8.])
    with pytest.raises(RuntimeError):
        det([torch.rand(5, 1, 64, 4, 4, device="cuda")])


@pytest.mark.parametrize("name", ["tiny_64", "tiny_ee_64", "tiny_b_64"])
def test_model_loss_fast_precision(name):
    """The same whole-model + loss cases in the BENCHMARK precision (one bf16 weight plane, fp16 ECS trace), end to end
    and NOT teacher-forced: what moves the loss here is the bf16 rounding of the conv weights and the near-threshold
    spikes it flips (tests/test_gpu_baseline_cfgs.py holds every layer of the real plans to 1e-3 / 99.9 % teacher-forced
    in this precision).  Measured on B200 and printed (1.8 - 7.2 % on these chaotic miniatures); gate: 15 % of the reference loss."""
    m, gold, loss, items = _model_loss_case(name, "fast")
    rel = abs(float(loss.reshape(-1)[0]) - float(gold["loss"].reshape(-1)[0])) / abs(float(gold["loss"].reshape(-1)[0]))
    print(f"\n{name} [fast] loss {float(loss.reshape(-1)[0]):.6f} vs reference {float(gold['loss'].reshape(-1)[0]):.6f} (rel {rel:.2e})")
    assert rel < 0.15, rel
    params = dict(m.named_parameters())
    assert all(torch.isfinite(p.grad).all() for p in params.values() if p.grad is not None)


def test_tdbn_glue_kernels_match_torch():
    """ecsy_tdbn_finish / ecsy_tdbn_bwd_coef (one launch each) against the element-wise torch formulas they replaced --
    nn.BatchNorm3d's running-statistics update inside batch_norm_2d (models/common.py:668-700) and the batch-norm backward
    coefficients: fp32, same operation order, so the results agree to the last bit or two."""
    E = ecsy()
    F_ = E.functional
    L = F_._cabi.lib()
    g = torch.Generator().manual_seed(5)
    C, n, m, eps, ups = 192, 4.0 * 2 * 20 * 24, 0.1, 1e-5, 2
    mean, var = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.1
    w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    nbt = torch.tensor(7, dtype=torch.int64)
    # reference (the torch sequence of autograd._bn_train_fwd_torch)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    for _ in range(ups):
        rm_ref.mul_(1.0 - m).add_(mean, alpha=m)
        rv_ref.mul_(1.0 - m).add_(var, alpha=m * n / (n - 1.0))
    rstd_ref = torch.rsqrt(var + eps)
    scale_ref = w * rstd_ref
    shift_ref = b - mean * scale_ref
    d = lambda t: t.clone().cuda()
    mean_d, var_d, w_d, b_d, rm_d, rv_d, nbt_d = d(mean), d(var), d(w), d(b), d(rm), d(rv), d(nbt)
    scale, shift, rstd = (torch.empty(C, device="cuda") for _ in range(3))
    F_._cabi.check(L.ecsy_tdbn_finish(F_._p(mean_d), F_._p(var_d), F_._p(w_d), F_._p(b_d), F_._p(rm_d), F_._p(rv_d), F_._p(nbt_d),
                                      m, n / (n - 1.0), eps, ups, F_._p(scale), F_._p(shift), F_._p(rstd), C, F_._st()), "tdbn_finish")
    assert int(nbt_d) == 7 + ups
    for got, want in ((rm_d, rm_ref), (rv_d, rv_ref), (rstd, rstd_ref), (scale, scale_ref), (shift, shift_ref)):
        assert torch.allclose(got.cpu(), want, rtol=3e-6, atol=1e-7), float((got.cpu() - want).abs().max())
    # without running statistics (track_running_stats = False): only the affine
    F_._cabi.check(L.ecsy_tdbn_finish(F_._p(mean_d), F_._p(var_d), F_._p(w_d), F_._p(b_d), None, None, None, m, 1.0, eps, 0,
                                      F_._p(scale), F_._p(shift), F_._p(rstd), C, F_._st()), "tdbn_finish")
    assert torch.allclose(scale.cpu(), scale_ref, rtol=3e-6)
    # backward coefficients
    sg, sgy = torch.randn(C, generator=g) * 10, torch.randn(C, generator=g) * 10
    tfac = 4.0
    sgx_ref = rstd_ref * (sgy - mean * sg)
    A_ref = w * rstd_ref
    B_ref = -A_ref * rstd_ref * sgx_ref / n
    C_ref = (-A_ref * sg / n - B_ref * mean) * tfac
    B_ref = B_ref * tfac
    A, B, Cc, gw = (torch.empty(C, device="cuda") for _ in range(4))
    sg_d, sgy_d, rstd_d = d(sg), d(sgy), d(rstd_ref)     # named: a temporary would be freed (and its block reused) before the launch
    F_._cabi.check(L.ecsy_tdbn_bwd_coef(F_._p(sg_d), F_._p(sgy_d), F_._p(mean_d), F_._p(rstd_d), F_._p(w_d), n, tfac,
                                        F_._p(A), F_._p(B), F_._p(Cc), F_._p(gw), C, F_._st()), "tdbn_bwd_coef")
    for got, want in ((A, A_ref), (B, B_ref), (Cc, C_ref), (gw, sgx_ref)):
        assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-7), float((got.cpu() - want).abs().max())


# ---------------------------------------------------------------- tier 4: training curves against the reference
@pytest.mark.parametrize("name", ["tiny_64", "tiny_b_64"])
def test_training_trajectory_matches_reference(name):
    """SURVEY 8c tier 4.  Six optimizer steps on one fixed batch from identical weights: our model (parity precision,
    surrogate-gradient BPTT through the kernels), our device ComputeLoss and the fused SGD-Nesterov step, against the
    UNMODIFIED reference doing the same with its own model / ComputeLoss / autograd / torch.optim.SGD over the three
    parameter groups of train.py:262-287 (oracle/gen_golden_trajectory.py -> tests/golden/train_trajectory.pt).  The loss
    of EVERY step, the head's parameters after the last step and how far every parameter moved must agree."""
    E = ecsy()
    E.set_precision("parity")
    stack_b = name in S.MODEL_B_CASES
    spec = (S.MODEL_B_CASES if stack_b else S.MODEL_CASES)[name]
    gold = torch.load(os.path.join(S.GOLDEN_DIR, "train_trajectory.pt"), weights_only=False)[name]
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    m = (E.yolo_snn.DetectionModel if stack_b else E.yolo.Model)(E.cfg_path(spec["cfg"]))
    m.load_state_dict(inp["sd"])
    m = m.cuda().train()
    m.hyp = dict(S.MODEL_LOSS_HYP)
    crit = (E.loss_tal if stack_b else E.loss).ComputeLoss(m)
    hp = S.TRAJECTORY_HYP
    opt = E.optim.SGDNesterovEMA(m, lr=hp["lr"], momentum=hp["momentum"], weight_decay=hp["weight_decay"], ema=False)
    x, tg = inp["x"].cuda(), S.model_targets(spec, cfg["nc"]).cuda()
    start = {k: p.detach().clone() for k, p in m.named_parameters()}
    losses = []
    for _ in range(hp["steps"]):
        opt.zero_grad(set_to_none=True)
        loss, _items = crit(m(x), tg)
        loss.sum().backward()
        opt.step()
        losses.append(float(loss.sum()))
    want = gold["losses"].tolist()
    dev = [abs(a - b) / abs(b) for a, b in zip(losses, want)]
    params = dict(m.named_parameters())
    head_err = {k: rel_l2(params[k].detach().cpu(), v) for k, v in gold["head_params"].items()}
    moved = {k: float((p.detach() - start[k]).norm()) for k, p in params.items()}
    ratios = sorted(moved[k] / v for k, v in gold["moved"].items() if v > 1e-6)
    med, lo10, hi90 = ratios[len(ratios) // 2], ratios[len(ratios) // 10], ratios[(9 * len(ratios)) // 10]
    print(name, "losses", [round(v, 6) for v in losses], "reference", [round(v, 6) for v in want],
          "rel dev", [f"{d:.1e}" for d in dev], "head param err", f"{max(head_err.values()):.1e}",
          "displacement ratio (ours / reference) median", round(med, 4), "10 % / 90 % quantiles", round(lo10, 4), round(hi90, 4))
    # Step 0 (identical weights) is the 1e-3 bar of test_model_loss_matches_reference.  From step 1 on the bar is the
    # reference's own noise floor: a gradient that differs in the sixth digit moves a handful of near-threshold spikes and
    # the loss by a few per cent (two CPU evaluation orders of the SAME fp32 arithmetic differ by 1.5 % at step 5,
    # tests/test_oracle_post.py::test_training_trajectory_oracle; PyTorch-vs-PyTorch 1.3-2.8 %, SURVEY 8c tier 3), so the
    # curves must OVERLAP, not coincide.  Measured on B200: tiny_64 2.7e-6 then 2.4-6.4 %; tiny_b_64 2.3e-6 then 0.4-9.7 %
    # along a loss that falls 175 -> 30.
    # The bounds leave a factor ~3 over the measured deviations: the gradient reductions use floating-point atomics, so WHICH
    # near-threshold spikes flip differs from run to run; a wrong learning rate, momentum, weight decay or parameter group
    # moves the curve / the displacements by far more.
    assert dev[0] < 1e-4, dev
    assert max(dev) < 0.30 and sum(dev) / len(dev) < 0.15, dev
    assert max(head_err.values()) < 0.30, head_err      # measured: 7.5e-3 (tiny_64), 6.0e-2 (tiny_b_64: a bias of the box branch)
    assert 0.85 < med < 1.15 and lo10 > 0.5 and hi90 < 2.0, (med, lo10, hi90)
