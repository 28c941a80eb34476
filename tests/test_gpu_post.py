"""GPU parity of the section-8f rows (post-decode NMS, fused SGD + EMA step) against the golden outputs of the
unmodified reference functions and the CPU oracle (oracle/post_oracle.py)."""
import copy
import os

import pytest
import torch

import post_oracle as P
import seeded as S
from util import ecsy

pytestmark = pytest.mark.gpu


def _load(name):
    return torch.load(os.path.join(S.GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", list(S.NMS_CASES))
def test_nms_golden(name):
    """Bit-exact: same boxes, same order, same values as utils.general.non_max_suppression."""
    E = ecsy()
    gold = _load("post_nms")[name]
    spec = S.NMS_CASES[name]
    pred = S.nms_inputs(spec)
    out = E.general.non_max_suppression(pred.cuda(), spec["conf"], spec["iou"], classes=spec.get("classes"),
                                        agnostic=spec.get("agnostic", False), multi_label=spec.get("multi_label", False),
                                        max_det=spec.get("max_det", 300))
    assert len(out) == len(gold["out"])
    for i, (a, b) in enumerate(zip(out, gold["out"])):
        assert tuple(a.shape) == tuple(b.shape), (name, i, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), (name, i)


def test_nms_full_size_properties():
    """BASELINE shape (batch 64, 6000 rows, nc = 13): size-independent properties -- descending confidence, every
    kept pair of one class overlaps <= the threshold, idempotence (NMS of the kept boxes keeps all of them), and the
    first images equal the CPU oracle."""
    E = ecsy()
    spec = dict(N=64, R=6000, nc=13, seed=777)
    pred = S.nms_inputs(spec)
    conf, iou = 0.25, 0.45
    out, cnt = E.general.nms_padded(pred.cuda(), conf, iou)
    cnt = cnt.cpu()
    out = out.cpu()
    want = P.non_max_suppression(pred[:3].clone(), conf, iou)
    for i in range(3):
        assert int(cnt[i]) == want[i].shape[0] and torch.equal(out[i, :cnt[i]], want[i])
    for i in range(spec["N"]):
        d = out[i, :cnt[i]]
        assert bool((d[1:, 4] <= d[:-1, 4]).all()) and bool((d[:, 4] > conf).all())
        boxes = d[:, :4] + d[:, 5:6] * 4096
        keep = P.greedy_nms(boxes, d[:, 4], iou)
        assert keep.numel() == d.shape[0]
    # idempotence through the kernel itself: feed the kept boxes back as predictions (obj = conf, one-hot class)
    d = out[0, :cnt[0]]
    back = torch.zeros(1, d.shape[0], 5 + spec["nc"])
    back[0, :, 0] = (d[:, 0] + d[:, 2]) / 2
    back[0, :, 1] = (d[:, 1] + d[:, 3]) / 2
    back[0, :, 2] = d[:, 2] - d[:, 0]
    back[0, :, 3] = d[:, 3] - d[:, 1]
    back[0, :, 4] = d[:, 4]
    back[0, torch.arange(d.shape[0]), 5 + d[:, 5].long()] = 1.0
    again = E.general.non_max_suppression(back.cuda(), conf * 0.5, iou)[0]
    assert again.shape[0] == d.shape[0]


def test_nms_many_batches_equals_oracle():
    """More than 8192 candidates per image are consumed in score-ordered batches (radix select + shared-memory sort per
    batch).  With max_det far above what survives early, the scan has to walk through SEVERAL batches up to the max_nms
    truncation: same boxes, same order, same values as the CPU oracle (itself pinned to utils.general.non_max_suppression
    by the golden fixtures); max_det = 300 on the same input stops inside the first batch and must give the same prefix."""
    E = ecsy()
    spec = dict(N=2, R=6000, nc=13, seed=778)
    pred = S.nms_inputs(spec)
    conf, iou = 0.001, 0.6
    want = P.non_max_suppression(pred.clone(), conf, iou, multi_label=True, max_det=3000)
    got = E.general.non_max_suppression(pred.cuda(), conf, iou, multi_label=True, max_det=3000)
    short = E.general.non_max_suppression(pred.cuda(), conf, iou, multi_label=True, max_det=300)
    for i in range(spec["N"]):
        assert want[i].shape[0] > 300, "the case must need more than the first batch's survivors"
        assert tuple(got[i].shape) == tuple(want[i].shape), (got[i].shape, want[i].shape)
        assert torch.equal(got[i].cpu(), want[i])
        assert torch.equal(short[i].cpu(), want[i][:300])


def test_nms_rejects_bad_arguments():
    E = ecsy()
    with pytest.raises(AssertionError):
        E.general.non_max_suppression(torch.zeros(1, 4, 8).cuda(), conf_thres=1.5)
    with pytest.raises(RuntimeError):
        E.general.non_max_suppression(torch.zeros(1, 4, 8))          # CPU tensor: no fallback
    assert E.general.non_max_suppression(torch.zeros(2, 4, 8).cuda())[1].shape == (0, 6)


def test_sgd_ema_golden():
    """Three optimizer steps (warm-up learning rates per group, weight decay on g1 only, BN buffers moving) + EMA
    updates against torch.optim.SGD + utils.torch_utils.ModelEMA run by the reference (fixture): fp32 tolerance 1e-6
    (the reference's CPU kernels fuse a + alpha*b differently), EMA of buffers included."""
    E = ecsy()
    gold = _load("post_opt")
    spec = S.OPT_CASE
    model, grads = S.opt_inputs(spec)
    model = model.cuda()
    opt = E.optim.SGDNesterovEMA(model, lr=spec["lr"], momentum=spec["momentum"], weight_decay=spec["weight_decay"])
    g0, g1, g2 = S.opt_groups(model)
    assert [len(g["params"]) for g in opt.param_groups] == [len(g0), len(g1), len(g2)]
    for step, gs in enumerate(grads):
        for p, g in zip(model.parameters(), gs):
            p.grad = g.cuda()
        for j, pg in enumerate(opt.param_groups):
            pg["lr"] = spec["lr"] * (1.0 + 0.1 * step) * (1.5 if j == 2 else 1.0)
        with torch.no_grad():
            model.bn.running_mean.add_(0.01 * (step + 1))
        opt.step()
        opt.zero_grad()
        want = gold["states"][step]
        assert opt.ema.updates == want["updates"]
        sd, esd = model.state_dict(), opt.ema.ema.state_dict()
        for k in sd:
            assert torch.allclose(sd[k].cpu().float(), want["model"][k].float(), rtol=1e-6, atol=1e-8), (step, k)
            assert torch.allclose(esd[k].cpu().float(), want["ema"][k].float(), rtol=1e-6, atol=1e-8), (step, k)


def test_sgd_ema_vs_torch_cuda_on_model():
    """On a real model (tiny Stack-A plan, 235 tensors): identical to torch.optim.SGD(nesterov) on the GPU to 1e-6, a
    parameter without a gradient is left alone, and the whole step is one kernel launch."""
    E = ecsy()
    F = E.functional
    torch.manual_seed(3)
    m1 = E.yolo.Model(E.cfg_path("tiny")).cuda()
    m2 = copy.deepcopy(m1)
    opt = E.optim.SGDNesterovEMA(m1, lr=0.02, momentum=0.9, weight_decay=1e-3)
    h0, h1, h2 = E.optim.param_groups_of(m2)
    ref = torch.optim.SGD(h0, lr=0.02, momentum=0.9, nesterov=True)
    ref.add_param_group({"params": h1, "weight_decay": 1e-3})
    ref.add_param_group({"params": h2})
    g = S.gen(5)
    skip = next(iter(m1.parameters()))
    for step in range(2):
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = (torch.randn(p1.shape, generator=g) * 0.05).cuda()
            if p1 is skip:
                p1.grad, p2.grad = None, None
            else:
                p1.grad, p2.grad = gr.clone(), gr.clone()
        n0 = F.launches["n"]
        opt.step()
        assert F.launches["n"] - n0 == 1
        ref.step()
    for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), k
    d1, d2 = opt.ema.decay(1), opt.ema.decay(2)
    assert opt.ema.updates == 2 and 0 < d1 < d2 < 1


def _flat_events(samples, T):
    xs, ys, ps, fs = [], [], [], []
    for n, bins in enumerate(samples):
        for t in range(T):
            b = bins[t]
            xs.append(b["x"]); ys.append(b["y"]); ps.append(b["p"])
            fs.append(torch.full_like(b["x"], n * T + t))
    cat = lambda v: torch.cat(v).to(torch.int32).cuda()
    return cat(xs), cat(ys), cat(ps), cat(fs)


@pytest.mark.parametrize("name", list(S.EVENT_CASES))
def test_event_frames_golden(name):
    """Bit-exact against the reference's create_data + cv2.resize + /255 (fixture), incl. an empty bin, pixels hit by
    many events (the last one wins) and both network sizes."""
    E = ecsy()
    gold = _load("post_events")[name]
    spec = S.EVENT_CASES[name]
    samples = S.event_inputs(spec)
    x, y, p, f = _flat_events(samples, spec["T"])
    out = E.events.event_frames(x, y, p, f, spec["N"], spec["T"], (spec["out"], spec["out"]), check=True)
    assert out.shape == (spec["T"], spec["N"], 3, spec["out"], spec["out"])
    want = (S.zunpack(gold["resized_ch0"]).float() / 255).permute(1, 0, 2, 3)      # [T, N, S, S]
    got = out.cpu()
    for c in range(3):
        assert torch.equal(got[:, :, c], want), (name, c)
    # identity size = the painted frames themselves
    same = E.events.event_frames(x, y, p, f, spec["N"], spec["T"], (240, 304)).cpu()
    painted = S.zunpack(gold["painted_ch0"]).float() / 255                            # [N, T, 240, 304]
    assert torch.equal(same[:, :, 0], painted.permute(1, 0, 2, 3))


def test_event_frames_feed_the_model_and_flag_bad_events():
    E = ecsy()
    spec = S.EVENT_CASES["ev_t4_sparse"]
    samples = S.event_inputs(spec)
    x, y, p, f = _flat_events(samples, spec["T"])
    frames = E.events.event_frames(x, y, p, f, spec["N"], spec["T"], (64, 64))
    a = E.functional.Act.from_ref(frames)
    assert a.data.data_ptr() == frames.data_ptr()            # NHWC memory: no layout conversion in front of the stem
    E.common.time_window = spec["T"]
    try:
        m = E.yolo.Model(E.cfg_path("tiny"), nc=2).cuda().train()
        with torch.no_grad():
            out = m._forward_once(frames)
    finally:
        E.common.time_window = 4
    assert all(torch.isfinite(o).all() for o in out)
    xb = x.clone(); xb[5] = 304
    with pytest.raises(AssertionError):
        E.events.event_frames(xb, y, p, f, spec["N"], spec["T"], (64, 64), check=True)
    empty = torch.zeros(0, dtype=torch.int32).cuda()
    grey = E.events.event_frames(empty, empty, empty, empty, 1, 2, (32, 32))
    assert torch.equal(grey.cpu(), torch.full((2, 1, 3, 32, 32), 127.0) / 255)


# ---- Stack-A training loss ------------------------------------------------------------------
class _LossHolder(torch.nn.Module):
    """What ComputeLoss.__init__ reads off a model (utils/loss.py:131-160)."""

    def __init__(self, anchors, nc, hyp):
        super().__init__()
        import types
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.hyp = dict(hyp)
        nl = anchors.shape[0]
        self.model = [types.SimpleNamespace(na=anchors.shape[1], nc=nc, nl=nl, anchors=anchors.cuda(),
                                            stride=torch.tensor([16.0, 32.0, 64.0][:nl]))]


@pytest.mark.parametrize("name", list(S.LOSS_CASES))
def test_compute_loss_golden(name):
    """ecs.loss.ComputeLoss (ecsy_yolo_loss through autograd) against utils.loss.ComputeLoss of the unmodified
    reference: loss and loss_items within 1e-5 relative, gradients w.r.t. the raw head outputs within 1e-4 relative
    (fp32 transcendental differences; cells hit by several matches accumulate with float atomics)."""
    E = ecsy()
    gold = _load("post_loss")[name]
    spec = S.LOSS_CASES[name]
    inp = S.loss_inputs(spec)
    crit = E.loss.ComputeLoss(_LossHolder(inp["anchors"], spec["nc"], spec["hyp"]))
    for extra in range(spec.get("calls", 1) - 1):              # SlideLoss keeps an EMA across calls (utils/loss.py:50-59)
        pre = S.loss_inputs(dict(spec, seed=spec["seed"] + 50 + extra))
        crit([x.cuda() for x in pre["p"]], pre["targets"].cuda())
    p = [x.cuda().requires_grad_(True) for x in inp["p"]]
    loss, items = crit(p, inp["targets"].cuda())
    assert loss.shape == (1,) and items.shape == (3,) and not items.requires_grad
    assert torch.allclose(loss.detach().cpu(), gold["loss"], rtol=1e-5, atol=1e-6), (float(loss), float(gold["loss"]))
    assert torch.allclose(items.cpu(), gold["items"], rtol=1e-5, atol=1e-6)
    (loss * inp["gout"]).sum().backward()
    for x, g in zip(p, gold["grads"]):
        err = (x.grad.cpu() - g).abs().max() / g.abs().max().clamp_min(1e-12)
        assert float(err) < 1e-4, (name, float(err))
        assert torch.allclose(x.grad.cpu(), g, rtol=1e-3, atol=1e-7 * float(g.abs().max()) + 1e-10)


def test_compute_loss_full_size_vs_oracle():
    """BASELINE training shape (batch 32, 40x40 + 20x20 levels, nc = 13, ~6.5 boxes per image): against the CPU oracle,
    bit-reproducible loss across runs, forward-only call, an empty target list, and an out-of-range image index."""
    import loss_oracle as LO
    E = ecsy()
    spec = dict(N=32, nc=13, grids=[(40, 40), (20, 20)], anchors=S._ANCH2, nt=208, seed=811, hyp=S._HYP)
    inp = S.loss_inputs(spec)
    hyp = dict(S._HYP, box=0.05 * 3 / 2, cls=0.5 * 13 / 80 * 3 / 2, obj=1.0 * 3 / 2)     # train.py:427-433 scaling, nl = 2
    p_ref = [x.clone().requires_grad_(True) for x in inp["p"]]
    want, items, counts, objs = LO.compute_loss(p_ref, inp["targets"], inp["anchors"], hyp)
    want.sum().backward()
    assert min(counts) > 0
    kw = dict(balance=LO.BALANCE_DEFAULT, box=hyp["box"], obj=hyp["obj"], cls=hyp["cls"], anchor_t=hyp["anchor_t"])
    pc = [x.cuda() for x in inp["p"]]
    out, grads = E.loss.yolo_loss(pc, inp["targets"].cuda(), inp["anchors"].cuda(), **kw)
    o = out.cpu()
    assert torch.allclose(o[0:1], want.detach(), rtol=1e-5)
    assert torch.allclose(o[1:4], items, rtol=1e-5)
    assert torch.allclose(o[4:6], torch.tensor(objs), rtol=1e-5)
    for g, x in zip(grads, p_ref):
        err = (g.cpu() - x.grad).abs().max() / x.grad.abs().max()
        assert float(err) < 1e-4, float(err)
    out2, none = E.loss.yolo_loss(pc, inp["targets"].cuda(), inp["anchors"].cuda(), need_grad=False, **kw)
    assert none == [] and torch.equal(out2, out)                        # fixed-order reductions: same bits
    # no targets: objectness against an all-zero target only
    out0, g0 = E.loss.yolo_loss(pc, torch.zeros(0, 6).cuda(), inp["anchors"].cuda(), **kw)
    w0, it0, c0, _ = LO.compute_loss(inp["p"], torch.zeros(0, 6), inp["anchors"], hyp)
    assert c0 == [0, 0] and torch.allclose(out0[0:1].cpu(), w0, rtol=1e-5) and float(out0[1]) == 0.0 == float(out0[3])
    assert float(g0[0][..., :4].abs().max()) == 0.0 and float(g0[0][..., 5:].abs().max()) == 0.0
    # a target whose image index is outside the batch is ignored instead of read out of bounds
    bad = torch.cat([inp["targets"], torch.tensor([[99.0, 1.0, 0.5, 0.5, 0.2, 0.2]])]).cuda()
    out3, _ = E.loss.yolo_loss(pc, bad, inp["anchors"].cuda(), need_grad=False, **kw)
    assert torch.allclose(out3, out, rtol=1e-6)


def test_compute_loss_rejects_unsupported_variants():
    E = ecsy()
    with pytest.raises(TypeError):      # the reference cannot combine the two wrappers either
        E.loss.ComputeLoss(_LossHolder(torch.tensor(S._ANCH2), 3, dict(S._HYP, fl_gamma=1.5, slide_ratio=1.0)))
    with pytest.raises(ValueError):
        E.loss.yolo_loss([torch.zeros(1, 3, 4, 4, 8).cuda()], torch.zeros(2, 5).cuda(), torch.ones(1, 3, 2).cuda(),
                         balance=[4.0], box=0.05, obj=1.0, cls=0.5)


# ---- Stack-B training loss: TAL -------------------------------------------------------------
class _TalHolder(torch.nn.Module):
    def __init__(self, spec, strides):
        super().__init__()
        import types
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.hyp = dict(cls_pw=spec.get("cls_pw", 1.0), fl_gamma=spec.get("fl_gamma", 0.0),
                        label_smoothing=spec.get("smooth", 0.0))
        self.model = [types.SimpleNamespace(nl=len(strides), nc=spec["nc"], no=64 + spec["nc"], reg_max=16, stride=strides)]


@pytest.mark.parametrize("name", list(S.TAL_CASES))
def test_tal_loss_golden(name, monkeypatch):
    """ecs.loss_tal.ComputeLoss (ecsy_tal_loss through autograd) against utils.loss_tal.ComputeLoss of the unmodified
    reference: same number of foreground anchors, loss / loss_items within 1e-5 relative, gradients within 1e-4."""
    E = ecsy()
    gold = _load("post_tal")[name]
    spec = S.TAL_CASES[name]
    inp = S.tal_inputs(spec)
    hp = spec.get("assigner", (10, 0.5, 6.0))
    for k, v in zip(("YOLOM", "YOLOA", "YOLOB"), hp):       # read at construction like the reference (utils/loss_tal.py:134-137)
        monkeypatch.setenv(k, str(v))
    crit = E.loss_tal.ComputeLoss(_TalHolder(spec, inp["strides"]))
    feats = [x.cuda().requires_grad_(True) for x in inp["feats"]]
    loss, items = crit(feats, inp["targets"].cuda())
    out, _ = E.loss_tal.tal_loss([x.detach() for x in feats], inp["targets"].cuda(), spec["strides"],
                                 spec.get("cls_pw", 1.0), need_grad=False, fl_gamma=spec.get("fl_gamma", 0.0), assigner=hp)
    assert int(out[4]) == gold["fg"], (int(out[4]), gold["fg"])
    assert abs(float(out[5]) - gold["score_sum"]) <= 1e-5 * max(gold["score_sum"], 1.0)
    assert loss.dim() == 0 and items.shape == (3,) and not items.requires_grad
    assert torch.allclose(loss.detach().cpu(), gold["loss"], rtol=1e-5), (float(loss), float(gold["loss"]))
    assert torch.allclose(items.cpu(), gold["items"], rtol=1e-5, atol=1e-6)
    (loss * inp["gout"]).backward()
    for x, g in zip(feats, gold["grads"]):
        err = (x.grad.cpu() - g).abs().max() / g.abs().max().clamp_min(1e-12)
        assert float(err) < 1e-4, (name, float(err))


def test_tal_loss_full_size_vs_oracle():
    """BASELINE training shape of resnet18.yaml (batch 16, 40x40 + 20x20 levels, nc = 80, ~6.5 boxes per image) against
    the CPU oracle; bit-reproducible across runs; a label outside the batch is ignored."""
    import tal_oracle as TO
    E = ecsy()
    spec = dict(N=16, nc=80, grids=[(40, 40), (20, 20)], strides=[16.0, 32.0], nt=104, seed=911, wh=(0.05, 0.35))
    inp = S.tal_inputs(spec)
    ref = [x.clone().requires_grad_(True) for x in inp["feats"]]
    want, items, n_fg = TO.compute_loss(ref, inp["targets"], spec["strides"])
    want.backward()
    fc = [x.cuda() for x in inp["feats"]]
    out, grads = E.loss_tal.tal_loss(fc, inp["targets"].cuda(), spec["strides"])
    assert int(out[4]) == n_fg and n_fg > 100
    assert torch.allclose(out[0].cpu(), want.detach(), rtol=1e-5), (float(out[0]), float(want))
    assert torch.allclose(out[1:4].cpu(), items, rtol=1e-5)
    for g, x in zip(grads, ref):
        err = (g.cpu() - x.grad).abs().max() / x.grad.abs().max()
        assert float(err) < 1e-4, float(err)
    out2, _ = E.loss_tal.tal_loss(fc, inp["targets"].cuda(), spec["strides"], need_grad=False)
    assert torch.equal(out2, out)
    bad = torch.cat([inp["targets"][:50], torch.tensor([[99.0, 1.0, 0.5, 0.5, 0.2, 0.2]]), inp["targets"][50:]]).cuda()
    out3, _ = E.loss_tal.tal_loss(fc, bad, spec["strides"], need_grad=False)
    assert torch.equal(out3, out)
