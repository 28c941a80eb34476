"""Generate tests/golden/*.pt by running the UNMODIFIED reference (container only).

    python oracle/gen_golden.py

Imports /root/reference through oracle/ref_shim.py, feeds it the seeded tensors of
tests/golden/seeded.py and stores the reference's OUTPUTS (plus an input checksum).  The fixtures
travel to the GPU box; the reference does not.  Test infrastructure, not product code.
"""
import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim  # noqa: E402
import seeded as S  # noqa: E402
import ecs_oracle as O  # noqa: E402  (key/shape bookkeeping only)


ONLY = set(sys.argv[1:])   # optional: regenerate just the named fixtures


def save(name, obj):
    path = os.path.join(S.GOLDEN_DIR, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def ref_lif(C, inp, T):
    m = C.mem_update()
    m.InitEcsSpread(inp["x"][0])
    with torch.no_grad():
        m.spread[0].weight.copy_(inp["dw_w"]); m.spread[0].bias.copy_(inp["dw_b"])
        m.spread[1].weight.copy_(inp["pw_w"]); m.spread[1].bias.copy_(inp["pw_b"])
    return m


def main():
    torch.set_num_threads(4)
    # ---------------- LIF (forward + surrogate-gradient BPTT) ----------------
    for name, spec in S.LIF_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, _ = ref_shim.load(spec["T"])
        inp = S.lif_inputs(spec)
        m = ref_lif(C, inp, spec["T"])
        x = inp["x"].clone().requires_grad_(True)
        out = m(x)
        out.backward(inp["gout"])
        g = {"gx": x.grad.clone()}
        for k, p in [("g_dw_w", m.spread[0].weight), ("g_dw_b", m.spread[0].bias),
                     ("g_pw_w", m.spread[1].weight), ("g_pw_b", m.spread[1].bias)]:
            g[k] = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
        save(name, dict(spec=spec, chk=S.checksum(*[inp[k] for k in sorted(inp)]),
                        spikes=S.pack_spikes(out.detach()), rate=float(out.mean()), **g))
    # ---------------- Snn_Conv2d ----------------
    for name, spec in S.CONV_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, _ = ref_shim.load(spec["T"])
        inp = S.conv_inputs(spec)
        m = C.Snn_Conv2d(spec["ci"], spec["co"], spec["k"], spec["s"], spec["p"], bias=inp["b"] is not None)
        with torch.no_grad():
            m.weight.copy_(inp["w"])
            if inp["b"] is not None:
                m.bias.copy_(inp["b"])
            out = m(inp["x"])
        save(name, dict(spec=spec, chk=S.checksum(inp["x"], inp["w"]), out=out))
    # ---------------- tdBN ----------------
    for name, spec in S.BN_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, _ = ref_shim.load(spec["T"])
        res = {}
        for cls, tag in [(C.batch_norm_2d, "bn1"), (C.batch_norm_2d1, "bn2")]:
            inp = S.bn_inputs(spec)
            m = cls(spec["C"])
            m.load_state_dict(inp["sd"])
            m.train()
            with torch.no_grad():
                res[tag + "_train"] = m(inp["x"]).contiguous()
            res[tag + "_sd_after"] = {k: v.clone() for k, v in m.state_dict().items()}
            m.eval()
            with torch.no_grad():
                res[tag + "_eval"] = m(inp["x"]).contiguous()
        inp = S.bn_inputs(spec)
        save(name, dict(spec=spec, chk=S.checksum(inp["x"]), **res))
    # ---------------- blocks ----------------
    for name, spec in {**S.BLOCK_CASES, **S.MS_BLOCK_CASES}.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, _ = ref_shim.load(spec["T"])
        inp = S.block_inputs(spec, O)
        cls = getattr(C, spec["kind"])
        m = cls(spec["cin"], spec["cout"], spec["s"]) if spec["kind"] == "BasicBlock_1" else \
            cls(spec["cin"], spec["cout"], spec["k"], spec["s"])
        m.train()
        with torch.no_grad():
            m(torch.zeros_like(inp["x"]))  # creates the lazy spread convs
        m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
        spikes = {}
        hooks = []
        for n_, mod in m.named_modules():
            if isinstance(mod, C.mem_update):
                hooks.append(mod.register_forward_hook(
                    lambda mod_, i_, o_, n_=n_: spikes.__setitem__(n_, S.pack_spikes(o_.detach())) or None))
        x = inp["x"].clone().requires_grad_(True)
        out_train = m(x)
        spikes_train = dict(spikes)
        gout = S.randn(S.gen(spec["seed"] + 13), *out_train.shape)
        out_train.backward(gout)
        # big weight grads are stored as a strided sample (every 97th element) + L2 norm
        grads = {}
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            gk = p.grad.detach()
            grads["model.0." + k] = gk.clone() if gk.numel() <= 65536 else \
                dict(sample=gk.flatten()[::97].clone(), norm=float(gk.norm()))
        sd_after = {"model.0." + k: v.clone() for k, v in m.state_dict().items()}
        m.eval()
        with torch.no_grad():
            out_eval = m(inp["x"])
        save(name, dict(spec=spec, chk=S.sd_checksum(inp["sd"]) + S.checksum(inp["x"]),
                        out_train=out_train.detach().contiguous(), spikes_train=spikes_train,
                        gx=x.grad.clone(), grads=grads,
                        bn_after={k: v for k, v in sd_after.items() if "running" in k or "tracked" in k},
                        out_eval=out_eval.contiguous(), spikes_eval=dict(spikes)))
        for h in hooks:
            h.remove()
    # ---------------- whole model (Stack A) ----------------
    for name, spec in S.MODEL_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, _ = ref_shim.load(spec["T"])
        path = os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")
        cfg = yaml.safe_load(open(path))
        inp = S.model_inputs(spec, O, cfg)
        m = Y.Model(path)
        m.load_state_dict(inp["sd"])
        assert torch.equal(m.stride, inp["stride"])
        rates = {}
        hooks = [mod.register_forward_hook(lambda mod_, i_, o_, n_=n_: rates.__setitem__(n_, float(o_.mean())) or None)
                 for n_, mod in m.named_modules() if isinstance(mod, C.mem_update)]
        m.train()
        with torch.no_grad():
            out_train = [o.clone() for o in m(inp["x"])]
        rates_train = dict(rates)
        # calibrated eval: momentum 1.0, one train-mode pass, then eval
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm3d):
                mod.momentum = 1.0
        with torch.no_grad():
            m(inp["x"])
        m.eval()
        feats = {}
        hk = [m.model[i].register_forward_hook(lambda mod_, i_, o_, i=i: feats.__setitem__(i, o_.detach().clone()) or None)
              for i in m.model[-1].f]
        with torch.no_grad():
            z, xs = m(inp["x"])
        save(name, dict(spec=spec, chk=S.sd_checksum(inp["sd"]) + S.checksum(inp["x"]),
                        out_train=out_train, rates_train=rates_train, z_eval=z, xs_eval=[o.clone() for o in xs],
                        rates_eval=dict(rates), head_feats={k: v for k, v in feats.items()},
                        bn_calibrated={k: v.clone() for k, v in m.state_dict().items()
                                       if "running" in k or "tracked" in k}))

    # ---------------- Stack B: SiLU neuron, Conv, DDetect, whole model ----------------
    for name, spec in S.SILU_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, SN = ref_shim.load(spec["T"])
        inp = S.lif_inputs(spec)
        m = C.mem_update(act=True)
        m.InitEcsSpread(inp["x"][0])
        m.actFun.inplace = spec["inplace"]
        with torch.no_grad():
            m.spread[0].weight.copy_(inp["dw_w"]); m.spread[0].bias.copy_(inp["dw_b"])
            m.spread[1].weight.copy_(inp["pw_w"]); m.spread[1].bias.copy_(inp["pw_b"])
            out = m(inp["x"].clone())
        x = inp["x"].clone().requires_grad_(True)
        og = m(x * 1.0)          # in-place SiLU must not hit a leaf
        og.backward(inp["gout"])
        g = {"gx": x.grad.clone(), "g_dw_w": m.spread[0].weight.grad.clone(), "g_dw_b": m.spread[0].bias.grad.clone(),
             "g_pw_w": m.spread[1].weight.grad.clone(), "g_pw_b": m.spread[1].bias.grad.clone()}
        save(name, dict(spec=spec, chk=S.checksum(*[inp[k] for k in sorted(inp)]), out=out, **g))
    for name, spec in S.CONVSILU_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, SN = ref_shim.load(spec["T"])
        inp = S.convsilu_inputs(spec, O)
        m = C.Conv(spec["cin"], spec["cout"], spec["k"], spec["s"])
        m.act.actFun.inplace = True   # as inside a built model (initialize_weights)
        m.train()
        with torch.no_grad():
            m(torch.zeros_like(inp["x"]))
        m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
        sd_before = {k: v.clone() for k, v in m.state_dict().items()}
        x = inp["x"].clone().requires_grad_(True)
        og = m(x)
        gout = S.randn(S.gen(spec["seed"] + 13), *og.shape)
        og.backward(gout)
        grads = {"model.0." + k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        gx = x.grad.clone()
        m.load_state_dict(sd_before)
        with torch.no_grad():
            out_train = m(inp["x"].clone())
            m.eval()
            out_eval = m(inp["x"].clone())
        save(name, dict(spec=spec, chk=S.sd_checksum(inp["sd"]) + S.checksum(inp["x"]), out_train=out_train,
                        out_eval=out_eval, gx=gx, grads=grads))
    for name, spec in S.DDETECT_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, SN = ref_shim.load(spec["T"])
        inp = S.ddetect_inputs(spec, O)
        m = SN.DDetect(spec["nc"], spec["ch"])
        m.stride = inp["stride"]
        m.train()
        with torch.no_grad():
            m([torch.zeros_like(f) for f in inp["feats"]])
        m.load_state_dict({k[len("model.0."):]: v for k, v in inp["sd"].items()})
        sd_before = {k: v.clone() for k, v in m.state_dict().items()}
        fs = [f.clone().requires_grad_(True) for f in inp["feats"]]
        og = m(list(fs))
        gouts = [S.randn(S.gen(spec["seed"] + 13 + i), *o.shape) for i, o in enumerate(og)]
        sum((o * g).sum() for o, g in zip(og, gouts)).backward()
        grads = {"model.0." + k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        gfeats = [f.grad.clone() for f in fs]
        m.load_state_dict(sd_before)
        with torch.no_grad():
            out_train = m([f.clone() for f in inp["feats"]])
            bn_after = {"model.0." + k: v.clone() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}
            m.eval()
            y, xs = m([f.clone() for f in inp["feats"]])
        save(name, dict(spec=spec, chk=S.sd_checksum(inp["sd"]) + S.checksum(*inp["feats"]),
                        out_train=[o.clone() for o in out_train], bn_after=bn_after, y_eval=y, xs_eval=[o.clone() for o in xs],
                        gfeats=gfeats, grads=grads))
    for name, spec in S.MODEL_B_CASES.items():
        if ONLY and name not in ONLY:
            continue
        C, Y, SN = ref_shim.load(spec["T"])
        path = os.path.join(ROOT, "ecs-yolo_b200", "cfg", spec["cfg"] + ".yaml")
        cfg = yaml.safe_load(open(path))
        inp = S.model_inputs(spec, O, cfg)
        m = SN.DetectionModel(path)
        m.load_state_dict(inp["sd"])
        assert torch.equal(m.stride, inp["stride"])
        rates = {}
        hooks = [mod.register_forward_hook(lambda mod_, i_, o_, n_=n_: rates.__setitem__(n_, float(o_.mean())) or None)
                 for n_, mod in m.named_modules() if isinstance(mod, C.mem_update)]
        m.train()
        with torch.no_grad():
            out_train = [o.clone() for o in m(inp["x"])]
        rates_train = dict(rates)
        bn_after = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}
        save(name, dict(spec=spec, chk=S.sd_checksum(inp["sd"]) + S.checksum(inp["x"]), out_train=out_train,
                        rates_train=rates_train, bn_after=bn_after))


if __name__ == "__main__":
    main()
