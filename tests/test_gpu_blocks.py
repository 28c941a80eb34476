"""GPU parity of the drop-in modules (blocks, Detect, whole model) against the golden outputs of the
unmodified reference.  Teacher-forced: every module gets the reference's recorded input."""
import os

import pytest
import torch
import yaml

import ecs_oracle as O
import seeded as S
from util import ROOT, agree, ecsy, forced_spikes, load_golden, quantize_weights_bf16, rel_l2

pytestmark = pytest.mark.gpu


def _build_block(E, spec):
    cls = getattr(E.common, spec["kind"])
    if spec["kind"] == "BasicBlock_1":
        return cls(spec["cin"], spec["cout"], spec["s"])
    return cls(spec["cin"], spec["cout"], spec["k"], spec["s"])


ALL_BLOCKS = {**S.BLOCK_CASES, **S.MS_BLOCK_CASES}   # MS_*: the res*-ee.yaml blocks, incl. the 3 / 32-channel front


def _oracle_block(inp, spec, sd, training, rec=None):
    return O.forward(inp["cfg"], sd, inp["x"], spec["T"], training, rec=rec)


class _oracle_forced:
    """ecs_oracle neurons return recorded spikes (evaluation of the oracle on bf16-rounded weights with the spikes of
    its fp32 run: the reference a fast-precision block is compared with when its own neurons are teacher-forced)."""

    def __init__(self, rec, own=None):
        self.rec, self.own = rec, own       # own: what each neuron computes from the input it gets in THIS evaluation

    def __enter__(self):
        self.orig = orig = O.lif_from_sd
        rec, own = self.rec, self.own

        def forced(sd, prefix, x, act=False, silu_inplace=False, record=None):
            key = prefix[:-1]
            if not act and key in rec:
                if own is not None:
                    own[key] = orig(sd, prefix, x, act=act, silu_inplace=silu_inplace, record=record)
                return rec[key]
            return orig(sd, prefix, x, act=act, silu_inplace=silu_inplace, record=record)
        O.lif_from_sd = forced

    def __exit__(self, *exc):
        O.lif_from_sd = self.orig
        return False


@pytest.mark.parametrize("name", list(ALL_BLOCKS))
def test_block_forward(name):
    E = ecsy()
    spec, gold = ALL_BLOCKS[name], load_golden(name)
    inp = S.block_inputs(spec, O)
    m = _build_block(E, spec)
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda()
    x = inp["x"].cuda()
    # --- train: batch statistics, running-stat update (the fixture ran train first, then eval)
    m.train()
    with torch.no_grad():
        out_t = m(x).cpu()
    e_train = rel_l2(out_t, gold["out_train"])
    # --- eval: tdBN (with the just-updated running stats) folded into the conv epilogues
    m.eval()
    with torch.no_grad():
        out = m(x).cpu()
    e_eval = rel_l2(out, gold["out_eval"])
    # A near-threshold flip in the first LIF changes a few conv outputs by O(weight); teacher-forced
    # north-star bound for real tensors is 1e-3 relative.
    assert e_eval < 1e-3 and e_train < 1e-3, f"{name}: eval {e_eval:.3e} train {e_train:.3e}"
    sd = m.state_dict()
    for k, v in gold["bn_after"].items():
        kk = k[len("model.0."):]
        assert torch.allclose(sd[kk].cpu().float(), v.float(), rtol=1e-3, atol=1e-5), k


@pytest.mark.parametrize("name", list(ALL_BLOCKS))
def test_block_forward_fast(name):
    """The benchmark precision (one bf16 weight plane, fp16 ECS trace, tanh.approx) at the north-star tolerances, block by
    block.  Every neuron is teacher-forced (util.forced_spikes): its spikes must equal, at >= 99.9 % of the positions, those
    of the fp32 oracle neuron applied to the same input (the bf16-weight conv output), and the real-valued block output must be within 1e-3 rel-L2 of the oracle evaluated on the same
    bf16-rounded conv / point-wise spread weights and the same spikes (exact products, fp32 accumulation), train and eval
    mode, including the running statistics.  The distance to the fp32-weight reference output is printed and bounded by
    the bf16 rounding of the weights.  Blocks whose shortcut convolves a REAL tensor (BasicBlock_ms) also round that
    operand to bf16: 6e-3."""
    E = ecsy()
    F = E.functional
    spec, gold = ALL_BLOCKS[name], load_golden(name)
    inp = S.block_inputs(spec, O)
    real_operand = spec["kind"] == "BasicBlock_ms" and (spec["s"] != 1 or spec["cin"] != spec["cout"])
    tol = 6e-3 if real_operand else 1e-3
    with torch.no_grad():
        sd32 = {k: v.clone() for k, v in inp["sd"].items()}
        rec_t, rec_e = {}, {}
        _oracle_block(inp, spec, sd32, True, rec_t)
        _oracle_block(inp, spec, sd32, False, rec_e)
        sdq = quantize_weights_bf16(inp["sd"])
        own_t, own_e = {}, {}
        with _oracle_forced(rec_t, own_t):
            want_t = _oracle_block(inp, spec, sdq, True)
        with _oracle_forced(rec_e, own_e):
            want_e = _oracle_block(inp, spec, sdq, False)
    spk = lambda r: {k: v for k, v in r.items() if not k.startswith("layer")}
    F.set_precision("fast")
    try:
        m = _build_block(E, spec)
        m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
        m = m.cuda().train()
        x = inp["x"].cuda()
        with torch.no_grad():
            with forced_spikes(E, m, spk(rec_t), "model.0.", gate_spikes=own_t) as ft:
                out_t = m(x).cpu()
            m.eval()
            with forced_spikes(E, m, spk(rec_e), "model.0.", gate_spikes=own_e) as fe:
                out_e = m(x).cpu()
        e_t, e_e = rel_l2(out_t, want_t), rel_l2(out_e, want_e)
        d_t, d_e = rel_l2(out_t, gold["out_train"]), rel_l2(out_e, gold["out_eval"])
        agree_min = min(list(ft.agree.values()) + list(fe.agree.values()))
        agree_fp32 = min(list(ft.agree_fp32.values()) + list(fe.agree_fp32.values()))
        print(f"\n{name} [fast]: train {e_t:.2e} eval {e_e:.2e} (vs fp32-weight reference {d_t:.2e} / {d_e:.2e}), "
              f"min spike agreement {agree_min:.6f} (vs the fp32-weight run's spikes {agree_fp32:.6f})")
        assert e_t < tol and e_e < tol, f"{name}: train {e_t:.3e} eval {e_e:.3e}"
        assert d_t < 1e-2 and d_e < 1e-2, f"{name}: vs the fp32-weight reference train {d_t:.3e} eval {d_e:.3e}"
        assert len(ft.agree) == len(spk(rec_t)) and len(fe.agree) == len(spk(rec_e))
        assert agree_min >= 0.999, {**ft.agree, **fe.agree}
        sd = m.state_dict()
        for k, v in sdq.items():       # the oracle's train pass updated sdq's running statistics in place
            if "running_" in k:
                kk = k[len("model.0."):]
                assert torch.allclose(sd[kk].cpu().float(), v.float(), rtol=2e-3, atol=1e-5), k
    finally:
        F.set_precision("parity")


def test_detect_head():
    E = ecsy()
    g = S.gen(77)
    T, N = 4, 2
    feats = [torch.randn(T, N, 128, 8, 8, generator=g), torch.randn(T, N, 192, 4, 4, generator=g)]
    anchors = [[10, 14, 23, 27, 37, 58], [81, 82, 135, 169, 344, 319]]
    det = E.yolo.Detect(3, anchors, (128, 192))
    det.stride = torch.tensor([8., 16.])
    det.anchors /= det.stride.view(-1, 1, 1)
    sd = S.reseed_state_dict({k: v for k, v in det.state_dict().items()}, 9)
    det.load_state_dict(sd)
    osd = {"model.0." + k: v for k, v in sd.items()}
    want_train = O.detect_a(osd, "model.0.", feats, 3, sd["anchors"], det.stride, True)
    want_z, want_x = O.detect_a(osd, "model.0.", feats, 3, sd["anchors"], det.stride, False)
    det = det.cuda()
    det.train()
    got = det([f.cuda() for f in feats])
    for a, b in zip(got, want_train):
        assert rel_l2(a.cpu(), b) < 1e-5
    det.eval()
    z, xs = det([f.cuda() for f in feats])
    assert rel_l2(z.cpu(), want_z) < 1e-5
    for a, b in zip(xs, want_x):
        assert rel_l2(a.cpu(), b) < 1e-5


@pytest.mark.parametrize("name", list(S.MODEL_CASES))
def test_model_tiny(name):
    """Whole Stack-A model: train-mode forward, momentum-1 calibration, eval decode.  End to end the
    network is chaotic, so outputs are held to the PyTorch-vs-PyTorch noise floor
    (a few % rel-L2); BN statistics and firing rates must match tightly."""
    E = ecsy()
    spec, gold = S.MODEL_CASES[name], load_golden(name)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    m = E.yolo.Model(E.cfg_path(spec["cfg"]))
    m.load_state_dict(inp["sd"])
    m = m.cuda()
    x = inp["x"].cuda()
    rates = {}
    orig = E.common.mem_update.spikes

    def rec(self, a, affine=None):
        sp = orig(self, a, affine)
        rates[id(self)] = float(sp.to_act().data.mean())
        return sp
    names = {id(mod): n for n, mod in m.named_modules() if isinstance(mod, E.common.mem_update)}
    E.common.mem_update.spikes = rec
    try:
        m.train()
        with torch.no_grad():
            out = m(x)
    finally:
        E.common.mem_update.spikes = orig
    for i, r in rates.items():
        assert abs(r - gold["rates_train"][names[i]]) < 5e-3, (names[i], r, gold["rates_train"][names[i]])
    errs = [rel_l2(a.cpu(), b) for a, b in zip(out, gold["out_train"])]
    assert max(errs) < (5e-2 if name == "tiny_64" else 0.15), errs
    # momentum-1 calibration, then eval.  End to end, eval mode amplifies single near-threshold flips
    # (stem rel error ~6e-6 -> a handful of flips at the first shortcut LIF -> O(1) divergence 3 blocks
    # later; the reference shows the same PyTorch-vs-PyTorch, SURVEY "facts" box item 5), so the eval
    # pass is held to: calibrated statistics equal, per-LIF firing rates equal, early LIFs >= 99.99 %
    # identical, and the head teacher-forced on the reference's own features.
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm3d):
            mod.momentum = 1.0
    got = {}

    def rec2(self, a, affine=None):
        sp = orig(self, a, affine)
        got[names[id(self)]] = float(sp.to_act().data.mean())
        return sp
    with torch.no_grad():
        m(x)
        sd = m.state_dict()
        for k, v in gold["bn_calibrated"].items():
            assert torch.allclose(sd[k].cpu().float(), v.float(), rtol=2e-3, atol=2e-4), k
        m.eval()
        E.common.mem_update.spikes = rec2
        try:
            z, xs = m(x)
        finally:
            E.common.mem_update.spikes = orig
    assert z.shape == gold["z_eval"].shape
    for n_, r in got.items():
        assert abs(r - gold["rates_eval"][n_]) < 1e-2, (n_, r, gold["rates_eval"][n_])
    det = m.model[-1]
    with torch.no_grad():
        zt, _ = det([gold["head_feats"][i].cuda() for i in det.f])
    assert rel_l2(zt.cpu(), gold["z_eval"]) < 1e-5


def test_model_event_frames():
    """BASELINE config 4 path: [T, N, 3, H, W] event frames (T = 5, a different frame every step) straight into
    _forward_once, against the CPU oracle on the same weights (train-mode tdBN: firing-rate-level agreement, and
    the stem -- the only layer that sees the frames directly -- to 1e-4)."""
    E = ecsy()
    T, N, H = 5, 2, 64
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", "tiny.yaml")))
    g = S.gen(77)
    u = torch.rand(N, T, 1, H, H, generator=g)
    f = torch.full_like(u, 127.0 / 255.0)
    f[u < 0.3] = 0.0     # dense events: a nearly constant frame makes train-mode tdBN amplify rounding noise
    f[u > 0.7] = 1.0
    x = f.expand(-1, -1, 3, -1, -1).permute(1, 0, 2, 3, 4).contiguous()      # [T, N, 3, H, W]
    sd = O.init_state_dict(cfg, T, seed=3)
    stride = O.detect_strides(cfg)
    rec = {}
    with torch.no_grad():
        want = O.forward(cfg, {k: v.clone() for k, v in sd.items()}, x, T, True, stride=stride, rec=rec)
    E.common.time_window = T
    try:
        m = E.yolo.Model(E.cfg_path("tiny"))
        m.load_state_dict(sd)
        m = m.cuda().train()
        feats = {}
        h = m.model[0].register_forward_hook(lambda mod, i, o: feats.__setitem__("stem", o.detach().cpu()))
        with torch.no_grad():
            out = m(x.cuda())
        h.remove()
    finally:
        E.common.time_window = 4
    assert feats["stem"].shape[0] == T and rel_l2(feats["stem"], rec["layer0"]) < 1e-4
    assert len(out) == len(want)
    for a_, b_ in zip(out, want):
        assert a_.shape == b_.shape
        assert rel_l2(a_.cpu(), b_) < 8e-2


# ---------------------------------------------------------------- Stack B
@pytest.mark.parametrize("name", list(S.SILU_CASES))
def test_silu_neuron(name):
    E = ecsy()
    F = E.functional
    spec, gold = S.SILU_CASES[name], load_golden(name)
    inp = S.lif_inputs(spec)
    w = F.make_lif_w(inp["dw_w"].cuda(), inp["dw_b"].cuda(), inp["pw_w"].cuda(), inp["pw_b"].cuda())
    out = F.lif_silu(F.Act.from_ref(inp["x"].cuda()), w, inplace=spec["inplace"]).to_ref().cpu()
    e = rel_l2(out, gold["out"])
    assert e < 2e-5, e


@pytest.mark.parametrize("name", list(S.CONVSILU_CASES))
def test_conv_silu(name):
    E = ecsy()
    spec, gold = S.CONVSILU_CASES[name], load_golden(name)
    inp = S.convsilu_inputs(spec, O)
    m = E.common.Conv(spec["cin"], spec["cout"], spec["k"], spec["s"])
    m.act.actFun.inplace = True
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda().train()
    with torch.no_grad():
        e1 = rel_l2(m(inp["x"].cuda()).cpu(), gold["out_train"])
        m.eval()
        e2 = rel_l2(m(inp["x"].cuda()).cpu(), gold["out_eval"])
    assert e1 < 5e-5 and e2 < 5e-5, (e1, e2)


@pytest.mark.parametrize("name", list(S.DDETECT_CASES))
def test_ddetect(name):
    E = ecsy()
    spec, gold = S.DDETECT_CASES[name], load_golden(name)
    inp = S.ddetect_inputs(spec, O)
    m = E.yolo_snn.DDetect(spec["nc"], spec["ch"])
    m.stride = inp["stride"]
    m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
    m = m.cuda().train()
    with torch.no_grad():
        out = m([f.cuda() for f in inp["feats"]])
    # teacher-forced branch outputs: two LIF layers deep, near-threshold flips allowed at the 1e-3 level
    for a, b in zip(out, gold["out_train"]):
        assert rel_l2(a.cpu(), b) < 1e-3
    sd = m.state_dict()
    for k, v in gold["bn_after"].items():   # momentum applied twice per forward (double evaluation)
        assert torch.allclose(sd[k[len("model.0."):]].cpu().float(), v.float(), rtol=1e-3, atol=1e-5), k
    m.eval()
    with torch.no_grad():
        y, xs = m([f.cuda() for f in inp["feats"]])
    assert rel_l2(y.cpu(), gold["y_eval"]) < 2e-3
    for a, b in zip(xs, gold["xs_eval"]):
        assert rel_l2(a.cpu(), b) < 2e-3


@pytest.mark.parametrize("name", list(S.MODEL_B_CASES))
def test_model_b_tiny(name):
    E = ecsy()
    spec, gold = S.MODEL_B_CASES[name], load_golden(name)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")))
    inp = S.model_inputs(spec, O, cfg)
    m = E.yolo_snn.DetectionModel(E.cfg_path(spec["cfg"]))
    m.load_state_dict(inp["sd"])
    m = m.cuda().train()
    rates = {}
    orig_s, orig_a = E.common.mem_update.spikes, E.common.mem_update.analog
    names = {id(mod): n for n, mod in m.named_modules() if isinstance(mod, E.common.mem_update)}

    def rec_s(self, a, affine=None):
        sp = orig_s(self, a, affine)
        rates[names[id(self)]] = float(sp.to_act().data.mean())
        return sp

    def rec_a(self, a, affine=None):
        o = orig_a(self, a, affine)
        rates[names[id(self)]] = float(o.data.mean())
        return o
    E.common.mem_update.spikes, E.common.mem_update.analog = rec_s, rec_a
    try:
        with torch.no_grad():
            out = m(inp["x"].cuda())
    finally:
        E.common.mem_update.spikes, E.common.mem_update.analog = orig_s, orig_a
    # End to end this miniature is chaotic: ~5 near-threshold flips out of 655k spikes in the first block
    # already move its (small) output by 4e-2 rel-L2 and the error saturates three blocks later (measured,
    # tests/diag/diag_model.py tiny_b).  The reference behaves the same PyTorch-vs-PyTorch,
    # so the whole-model check is statistical; exactness is established teacher-forced per block above.
    for n_, r in rates.items():
        assert abs(r - gold["rates_train"][n_]) < 1e-2, (n_, r, gold["rates_train"][n_])
    for a, b in zip(out, gold["out_train"]):
        assert a.shape == b.shape and torch.isfinite(a).all()
    sd = m.state_dict()
    for k, v in gold["bn_after"].items():
        if "tracked" in k:
            assert int(sd[k]) == int(v), k
        elif "model.0." in k or "model.1." in k:   # before the first possible divergence: tight
            assert torch.allclose(sd[k].cpu().float(), v.float(), rtol=5e-3, atol=5e-4), k
