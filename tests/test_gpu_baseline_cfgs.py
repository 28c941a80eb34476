"""Teacher-forced parity of the BASELINE.json plans themselves (SURVEY 8c tier 1 + 2), in BOTH precisions.

configs[0] -- resnet10.yaml, T=4, batch 2, 640x640 -- plus resnet34.yaml (configs[1]'s plan) and resnet18.yaml
(configs[2]'s plan, Stack B) at batch 1, through models/yolo.py:247-312 / models/yolo_snn.py:741-749.

Protocol.  The CPU oracle (oracle/ecs_oracle.py, pinned to the unmodified reference by tests/golden) runs the whole plan
once in fp32 with train-mode tdBN (default-init running statistics collapse the firing rate) and records every
layer's output and every neuron's spikes.  Each layer of OUR model is then fed the oracle's recorded input(s) and, inside
the layer, every mem_update hands the oracle's spikes to its consumer after counting how many of its own spikes agree
(util.forced_spikes): every LIF, conv and tdBN is compared on the reference's input, so a near-threshold flip can neither
masquerade as an arithmetic error nor hide one.

Gates (north_star): spikes >= 99.9 % per neuron in both precisions (fast: against the oracle neuron applied to the input our
neuron actually receives, i.e. the bf16-weight conv output -- the agreement with the fp32-weight run's spikes is printed
next to it and shows what the bf16 rounding of the UPSTREAM conv weights costs); real tensors <= 1e-3 rel-L2 --
  parity (bf16 hi+lo weight planes): against the fp32 oracle;
  fast (one bf16 weight plane, the benchmark mode): against the oracle evaluated on the same bf16-rounded operands (conv /
  point-wise spread weights, real conv inputs) and the same forced spikes -- the products are then exact and only the
  fp32 accumulation order differs; the distance to the fp32-weight oracle is printed and bounded by the bf16 rounding of
  the weights (6e-3).  The SiLU neuron of Stack B's `Conv` layers computes its trace in fp16 with tanh.approx in fast
  mode: 5e-3.
Head and loss: the head gets the oracle's P4 / P5 features, the loss our head output; both against the oracle's head +
loss oracle (utils/loss.py / utils/loss_tal.py restatements) within 1e-3.
"""
import os

import pytest
import torch
import yaml

import ecs_oracle as O
import loss_oracle as LO
import tal_oracle as TO
from util import ROOT, ecsy, forced_spikes, quantize_weights_bf16, rel_l2

pytestmark = pytest.mark.gpu

T = 4
PLANS = [("resnet10", 2, 640), ("resnet34", 1, 640), ("resnet18", 1, 640)]
REAL_INPUT_LAYERS = ("Conv_1", "Conv")       # their conv sees a real tensor: one bf16 plane of it in fast mode


def _targets(n_img, nc, seed):
    g = torch.Generator().manual_seed(seed)
    per = torch.randint(5, 9, (n_img,), generator=g)
    img = torch.repeat_interleave(torch.arange(n_img), per).float()
    n = int(img.numel())
    return torch.cat([img[:, None], torch.randint(0, nc, (n, 1), generator=g).float(),
                      torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1)


def _oracle_layer(sd, L, x, stride, anchors, rec):
    """One entry of the plan, like the loop of ecs_oracle.forward (train-mode tdBN)."""
    p = f"model.{L['i']}."
    if L["type"] == "Detect":
        return O.detect_a(sd, p, list(x), L["args"][0], anchors, stride, True)
    if L["type"] == "DDetect":
        return O.ddetect(sd, p, list(x), L["args"][0], stride, True, True, rec)
    if L["n"] > 1:
        for j in range(L["n"]):
            x = O._run_layer(sd, L, f"{p}{j}.", x, True, False, rec)
        return x
    return O._run_layer(sd, L, p, x, True, False, rec)


class _forced_oracle_spikes:
    """The oracle's neurons return the recorded spikes of the fp32 run (second, bf16-operand evaluation of a layer); what
    each neuron itself computes from the input it receives in THIS evaluation is kept in `own` (the reference our
    neuron's arithmetic is gated against: same input, fp32 oracle arithmetic)."""

    def __init__(self, rec, own=None):
        self.rec, self.own = rec, own

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


@pytest.mark.parametrize("mode", ["parity", "fast"])
@pytest.mark.parametrize("name,N,img", PLANS, ids=[p[0] for p in PLANS])
def test_baseline_plan_teacher_forced(name, N, img, mode):
    E = ecsy()
    F = E.functional
    cfg = yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", name + ".yaml")))
    stack_b = any(row[2] == "DDetect" for row in cfg["head"])
    torch.manual_seed(0)
    sd = O.init_state_dict(cfg, T, seed=0)
    stride = O.detect_strides(cfg)
    for k in sd:
        if k.endswith("anchors"):
            sd[k] = sd[k] / stride.view(-1, 1, 1)
    x = torch.rand(N, 3, img, img, generator=torch.Generator().manual_seed(0))
    layers, save = O.plan_model(cfg, 3)
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- reference run: fp32 oracle, whole plan, everything recorded
    rec = {}
    with torch.no_grad():
        head_ref = O.forward(cfg, {k: v.clone() for k, v in sd.items()}, x, T, True, stride=stride, rec=rec)
    xin = x.unsqueeze(0).expand(T, -1, -1, -1, -1)     # the direct-coded image: one frame, T-broadcast (models/yolo.py:248-251)

    def layer_input(L):
        f = L["f"]
        prev = lambda j: xin if (j < 0 and L["i"] + j < 0) else rec[f"layer{L['i'] + j if j < 0 else j}"]
        return prev(f) if isinstance(f, int) else [prev(j) for j in f]

    # ---- our model
    E.set_precision(mode)
    try:
        E.common.time_window = T
        m = (E.yolo_snn.DetectionModel if stack_b else E.yolo.Model)(E.cfg_path(name))
        missing = m.load_state_dict(sd, strict=False)
        assert not missing.missing_keys, missing.missing_keys[:4]
        m = m.cuda().train()
        sd_q = quantize_weights_bf16(sd) if mode == "fast" else None
        ref_spk = {k: v for k, v in rec.items() if not k.startswith("layer")}
        worst = dict(real=("", 0.0), fp32=("", 0.0), spike=("", 1.0))
        rows = []
        with torch.no_grad():
            for L in layers:
                if L["type"] in ("Detect", "DDetect"):
                    continue
                inp = layer_input(L)
                want32 = rec[f"layer{L['i']}"]
                want = want32
                own = None
                if mode == "fast" and L["type"] not in ("Sample", "Concat"):
                    qin = inp.bfloat16().float().contiguous() if L["type"] in REAL_INPUT_LAYERS else inp
                    own = {}
                    with _forced_oracle_spikes(ref_spk, own):
                        want = _oracle_layer({k: v.clone() for k, v in sd_q.items()}, L, qin, stride, None, None)
                mod = m.model[L["i"]]
                cu = [t.cuda() for t in inp] if isinstance(inp, list) else inp.cuda()
                with forced_spikes(E, m, ref_spk, gate_spikes=own) as fs:
                    got = mod(cu).cpu()
                e, e32 = rel_l2(got, want), rel_l2(got, want32)
                tol = 5e-3 if (mode == "fast" and L["type"] == "Conv") else 1e-3
                rows.append((L["i"], L["type"], e, e32, min(fs.agree.values()) if fs.agree else 1.0,
                             min(fs.agree_fp32.values()) if fs.agree_fp32 else 1.0))
                assert e < tol, f"{name} {mode} layer {L['i']} {L['type']}: rel-L2 {e:.3e} (vs fp32 oracle {e32:.3e})"
                assert e32 < (1e-3 if mode == "parity" else 6e-3), f"{name} {mode} layer {L['i']}: vs fp32 oracle {e32:.3e}"
                for k, a in fs.agree.items():
                    assert a >= 0.999, f"{name} {mode} {k}: spike agreement {a:.6f}"
                    if a < worst["spike"][1]:
                        worst["spike"] = (k, a)
                if e > worst["real"][1]:
                    worst["real"] = (f"layer{L['i']}", e)
                if e32 > worst["fp32"][1]:
                    worst["fp32"] = (f"layer{L['i']}", e32)
                del got, cu
            # ---- head (teacher-forced on the oracle's features) and loss
            L = layers[-1]
            feats = layer_input(L)
            det = m.model[-1]
            want_head = head_ref
            own = None
            if mode == "fast" and stack_b:
                own = {}
                with _forced_oracle_spikes(ref_spk, own):
                    want_head = _oracle_layer({k: v.clone() for k, v in sd_q.items()}, L, feats, stride, None, None)
            with forced_spikes(E, m, ref_spk, gate_spikes=own) as fs:
                out = det([f.cuda() for f in feats])
            for k, a in fs.agree.items():
                assert a >= 0.999, f"{name} {mode} head {k}: spike agreement {a:.6f}"
            e_head = max(rel_l2(a.cpu(), b) for a, b in zip(out, want_head))
            e_head32 = max(rel_l2(a.cpu(), b) for a, b in zip(out, head_ref))
            assert e_head < 1e-3, f"{name} {mode}: head rel-L2 {e_head:.3e}"
        nc = det.nc
        tg = _targets(N, nc, 1)
        if stack_b:
            m.hyp = dict(cls_pw=1.0, fl_gamma=0.0, label_smoothing=0.0)
            loss, items = E.loss_tal.ComputeLoss(m)([o.detach() for o in out], tg.cuda())
            want_loss, want_items, _ = TO.compute_loss([h.clone() for h in head_ref], tg, [float(s) for s in stride])
        else:
            nl = det.nl
            hyp = dict(box=0.05 * 3 / nl, cls=0.5 * nc / 80 * 3 / nl, obj=1.0 * (img / 640) ** 2 * 3 / nl, cls_pw=1.0,
                       obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)
            m.hyp = hyp
            loss, items = E.loss.ComputeLoss(m)([o.detach() for o in out], tg.cuda())
            want_loss, want_items, _, _ = LO.compute_loss([h.clone() for h in head_ref], tg, sd[f"model.{L['i']}.anchors"], hyp)
        # both ComputeLoss classes return loss * batch size (utils/loss.py:234, utils/loss_tal.py:215)
        e_loss = abs(float(loss.reshape(-1)[0]) - float(want_loss.reshape(-1)[0])) / abs(float(want_loss.reshape(-1)[0]))
        print(f"\n{name} [{mode}] teacher-forced: worst real {worst['real'][0]} {worst['real'][1]:.2e}, vs fp32 oracle "
              f"{worst['fp32'][0]} {worst['fp32'][1]:.2e}, min spike agreement {worst['spike'][1]:.6f} ({worst['spike'][0]}), "
              f"head {e_head:.2e} (fp32 {e_head32:.2e}), loss rel {e_loss:.2e}")
        for r in rows:
            print("   layer %2d %-14s rel-L2 %.2e  vs-fp32 %.2e  min-spike-agree %.6f  (vs the fp32-weight run's spikes %.6f)" % r)
        assert e_loss < 1e-3, f"{name} {mode}: loss {float(loss.reshape(-1)[0])} vs {float(want_loss.reshape(-1)[0])}"
        if os.environ.get("ECSY_WRITE_PARITY_SUMMARY"):
            # committed beside the bench numbers (profiles/parity_summary.json, read by bench.py's `parity` record)
            import json
            path = os.path.join(ROOT, os.environ["ECSY_WRITE_PARITY_SUMMARY"])
            summ = json.load(open(path)) if os.path.exists(path) else {}
            summ[f"{name}|{mode}"] = {
                "plan": f"cfg/{name}.yaml, T={T}, batch {N}, {img}x{img}, train-mode tdBN, every layer fed the oracle's input",
                "min_spike_agreement": worst["spike"][1], "min_spike_agreement_neuron": worst["spike"][0],
                "min_spike_agreement_vs_fp32_weight_run": min(r[5] for r in rows),
                "worst_layer_rel_l2": worst["real"][1], "worst_layer_rel_l2_vs_fp32_weight_oracle": worst["fp32"][1],
                "head_rel_l2": e_head, "loss_rel": e_loss,
                "reference": "fp32 oracle" if mode == "parity" else "oracle on the same bf16-rounded conv / spread weights"}
            json.dump(summ, open(path, "w"), indent=1)
    finally:
        E.set_precision("parity")
