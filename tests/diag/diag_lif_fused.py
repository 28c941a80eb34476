"""Diagnose the fused ECS-LIF kernel with structured spread weights (which taps / channel mixes are right?)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import ecs_oracle as O
E = importlib.import_module("ecs-yolo_b200")
F = E.functional
F.set_precision("fast")
T, N, C, H, W = 4, 1, 64, 16, 16
g = torch.Generator().manual_seed(1)
x = torch.randn(T, N, C, H, W, generator=g) * 0.5 + 0.1
def run(name, dw_w, dw_b, pw_w, pw_b):
    w = F.make_lif_w(dw_w.cuda(), dw_b.cuda(), pw_w.cuda(), pw_b.cuda())
    a = F.Act.from_ref(x.cuda())
    F.set_lif_fused(True); got = F.lif_ecs(a, w).to_act().to_ref().cpu()
    F.set_lif_fused(False); unf = F.lif_ecs(a, w).to_act().to_ref().cpu()
    want = O.ecs_lif(x, dw_w, dw_b, pw_w, pw_b)
    per_t = [float((got[t] == want[t]).float().mean()) for t in range(T)]
    per_tu = [float((unf[t] == want[t]).float().mean()) for t in range(T)]
    # agreement on interior pixels only (2-pixel border removed)
    gi, wi = got[:, :, :, 3:-3, 3:-3], want[:, :, :, 3:-3, 3:-3]
    print(f"{name:28s} fused {['%.4f' % v for v in per_t]}  unfused {['%.4f' % v for v in per_tu]}  interior {float((gi == wi).float().mean()):.4f}", flush=True)
z = torch.zeros
big = 3.0   # large weights so that a wrong ECS term flips many spikes
def dw_tap(ky, kx, v=1.0):
    d = z(C, 1, 3, 3); d[:, 0, ky, kx] = v; return d
eye = torch.eye(C).reshape(C, C, 1, 1)
run("pw=I*3 dw=center", dw_tap(1, 1), z(C), eye * big, z(C))
run("pw=I*3 dw=left(1,0)", dw_tap(1, 0), z(C), eye * big, z(C))
run("pw=I*3 dw=right(1,2)", dw_tap(1, 2), z(C), eye * big, z(C))
run("pw=I*3 dw=up(0,1)", dw_tap(0, 1), z(C), eye * big, z(C))
run("pw=I*3 dw=down(2,1)", dw_tap(2, 1), z(C), eye * big, z(C))
run("pw=I*3 dw=corner(0,0)", dw_tap(0, 0), z(C), eye * big, z(C))
run("pw=0 bias=1", dw_tap(1, 1, 0.0), z(C), eye * 0, torch.ones(C))
perm = torch.roll(torch.eye(C), 1, 0).reshape(C, C, 1, 1)
run("pw=shift1*3 dw=center", dw_tap(1, 1), z(C), perm * big, z(C))
perm17 = torch.roll(torch.eye(C), 17, 0).reshape(C, C, 1, 1)
run("pw=shift17*3 dw=center", dw_tap(1, 1), z(C), perm17 * big, z(C))
gw = torch.Generator().manual_seed(2)
run("random all", (torch.rand(C, 1, 3, 3, generator=gw) - .5) * 2 / 3, (torch.rand(C, generator=gw) - .5) * 2 / 3,
    (torch.rand(C, C, 1, 1, generator=gw) - .5) / 4, (torch.rand(C, generator=gw) - .5) / 4)
