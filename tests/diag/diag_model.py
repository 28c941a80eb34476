"""Per-layer divergence CUDA path vs CPU oracle (train mode, end to end).  usage: diag_model.py tiny|tiny_b"""
import importlib, os, sys, yaml, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ecs_oracle as O, seeded as S
from util import rel_l2
E = importlib.import_module("ecs-yolo_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
stack_b = name.endswith("_b")
spec = (S.MODEL_B_CASES if stack_b else S.MODEL_CASES)[name + "_64"]
cfg = yaml.safe_load(open(E.cfg_path(name)))
inp = S.model_inputs(spec, O, cfg)
m = (E.yolo_snn.DetectionModel if stack_b else E.yolo.Model)(E.cfg_path(name)); m.load_state_dict(inp["sd"]); m = m.cuda()
x = inp["x"].cuda()
sd = {k: v.clone() for k, v in inp["sd"].items()}
outs = {}
hooks = [mod.register_forward_hook(lambda mod_, i_, o_, i=i: outs.__setitem__(i, o_.detach().cpu()) if torch.is_tensor(o_) else None)
         for i, mod in enumerate(m.model)]
m.train()
with torch.no_grad():
    out = m(x)
rec = {}
with torch.no_grad():
    want = O.forward(cfg, sd, inp["x"], 4, True, stride=inp["stride"], rec=rec)
for i in sorted(outs):
    print(f"  layer {i} {m.model[i].type}: rel-L2 {rel_l2(outs[i], rec[f'layer{i}']):.3e}  shape {tuple(outs[i].shape)} strideT {outs[i].stride(0)}")
print("  head", [f"{rel_l2(a.cpu(), b):.3e}" for a, b in zip(out, want)])
