"""Per-layer divergence of the tiny model, CUDA path vs CPU oracle (end-to-end, not teacher-forced)."""
import importlib, os, sys, yaml, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ecs_oracle as O, seeded as S
from util import rel_l2, load_golden
E = importlib.import_module("ecs-yolo_b200")
spec = S.MODEL_CASES["tiny_64"]; gold = load_golden("tiny_64")
cfg = yaml.safe_load(open(E.cfg_path("tiny")))
inp = S.model_inputs(spec, O, cfg)
m = E.yolo.Model(E.cfg_path("tiny")); m.load_state_dict(inp["sd"]); m = m.cuda()
x = inp["x"].cuda()
sd = {k: v.clone() for k, v in inp["sd"].items()}

def run(training):
    outs = {}
    hooks = [mod.register_forward_hook(lambda mod_, i_, o_, i=i: outs.__setitem__(i, o_.detach().cpu()) if torch.is_tensor(o_) else None)
             for i, mod in enumerate(m.model)]
    m.train(training)
    with torch.no_grad():
        out = m(x)
    for h in hooks: h.remove()
    rec = {}
    with torch.no_grad():
        want = O.forward(cfg, sd, inp["x"], 4, training, stride=inp["stride"], rec=rec)
    for i in sorted(outs):
        print(f"  layer {i}: rel-L2 {rel_l2(outs[i], rec[f'layer{i}']):.3e}")
    return out, want

print("train"); out, want = run(True)
print("  head", [f"{rel_l2(a.cpu(), b):.3e}" for a, b in zip(out, want)], "vs golden", [f"{rel_l2(a.cpu(), b):.3e}" for a, b in zip(out, gold['out_train'])])
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm3d): mod.momentum = 1.0
O.BN_MOMENTUM = 1.0
print("calibrate"); run(True)
bad = [(k, float((m.state_dict()[k].cpu().float()-sd[k].float()).abs().max())) for k in sd if 'running' in k]
print("  max running-stat diff", max(b for _, b in bad))
print("eval"); (z, xs), (wz, wxs) = run(False)
print("  z", f"{rel_l2(z.cpu(), wz):.3e}", "raw", [f"{rel_l2(a.cpu(), b):.3e}" for a, b in zip(xs, wxs)], "z vs golden", f"{rel_l2(z.cpu(), gold['z_eval']):.3e}")
# teacher-forced head: reference features -> our Detect
det = m.model[-1]
feats = [gold["head_feats"][i].cuda() for i in det.f]
zz, _ = det(feats)
print("  teacher-forced head z vs golden", f"{rel_l2(zz.cpu(), gold['z_eval']):.3e}")

# ---- per-LIF spike agreement in eval mode
print("per-LIF agreement (eval)")
got = {}
orig = E.common.mem_update.spikes
names = {id(mod): n for n, mod in m.named_modules() if isinstance(mod, E.common.mem_update)}
def rec_sp(self, a, affine=None):
    sp = orig(self, a, affine)
    got[names[id(self)]] = sp.to_act().to_ref().cpu()
    return sp
E.common.mem_update.spikes = rec_sp
m.eval()
with torch.no_grad():
    m(x)
E.common.mem_update.spikes = orig
rec = {}
with torch.no_grad():
    O.forward(cfg, sd, inp["x"], 4, False, stride=inp["stride"], rec=rec)
for k in list(got)[:8]:
    r = rec[k]
    print(f"  {k}: agree {float((got[k]==r).float().mean()):.6f} rate {float(got[k].mean()):.4f} / {float(r.mean()):.4f}")
