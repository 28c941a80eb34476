"""Micro-benchmark of the ECS-LIF pipeline on single layers (batch 64, T=4): lif_ecs forward."""
import argparse, importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
F = E.functional
ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fast")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--N", type=int, default=64)
args = ap.parse_args()
F.set_precision(args.mode)
T, N = 4, args.N
SHAPES = [(64, 160), (128, 80), (256, 40), (512, 20), (64, 320)]
for idx, (C, H) in enumerate(SHAPES):
    if args.only >= 0 and idx != args.only:
        continue
    x = torch.randn(T, N, H, H, C, device="cuda") * 0.5
    a = F.Act(x, T)
    dw = torch.randn(C, 1, 3, 3, device="cuda") * 0.3
    pw = torch.randn(C, C, 1, 1, device="cuda") / C ** 0.5
    w = F.make_lif_w(dw, torch.zeros(C, device="cuda"), pw, torch.zeros(C, device="cuda"))
    for _ in range(2):
        sp = F.lif_ecs(a, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        sp = F.lif_ecs(a, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    elems = T * N * H * H * C
    print(dict(shape=f"C{C}@{H}", ms=round(ms, 3), gelem_s=round(elems / ms / 1e6, 1), rate=round(float(sp.to_act().data.mean()), 3)), flush=True)
    del x, a, sp
