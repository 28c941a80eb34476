"""Micro-benchmark: wavefront ECS-LIF kernel (csrc/lif_wave.cu) vs the per-timestep pipeline on the 64-channel layer
shapes of resnet34 at batch N, T = 4 (fast precision).  Algorithmic bytes = elems * (4 B of x + 1/8 B of spikes)."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
F = E.functional
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--N", type=int, default=64)
ap.add_argument("--T", type=int, default=4)
ap.add_argument("--shapes", default="160,80,320,320b")
args = ap.parse_args()
F.set_precision("fast")
T, N, C = args.T, args.N, 64
for tag in args.shapes.split(","):
    bcast = tag.endswith("b")
    H = int(tag.rstrip("b"))
    x = torch.randn(1 if bcast else T, N, H, H, C, device="cuda") * 0.5
    a = F.Act(x, T)
    dw = torch.randn(C, 1, 3, 3, device="cuda") * 0.3
    pw = torch.randn(C, C, 1, 1, device="cuda") / C ** 0.5
    w = F.make_lif_w(dw, torch.zeros(C, device="cuda"), pw, torch.zeros(C, device="cuda"))
    res = {}
    for name, on in (("pipeline", "off"), ("wave", "all")):
        F.set_lif_wave(on)
        for _ in range(2):
            sp = F.lif_ecs(a, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            sp = F.lif_ecs(a, w)
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / args.reps, sp.bits.clone())
    elems = T * N * H * H * C
    gb = elems * (4 + 0.125) / 1e9 if not bcast else (N * H * H * C * 4 + elems * 0.125) / 1e9
    agree = float((res["pipeline"][1] == res["wave"][1]).float().mean())
    print(dict(shape=f"C64@{H}{' T-broadcast' if bcast else ''}", pipeline_ms=round(res["pipeline"][0], 3),
               wave_ms=round(res["wave"][0], 3), wave_GBs=round(gb / res["wave"][0] * 1e3, 1),
               words_equal=round(agree, 5)), flush=True)
    del x, a, sp, res
