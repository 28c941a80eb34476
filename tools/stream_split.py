"""Experiment: eval forward of one batch as `n` concurrent sub-batches on `n` CUDA streams (issue-bound and HBM-bound
kernels of different sub-batches overlap).  python tools/stream_split.py [--streams 2] [--batch 64]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--model", default="resnet34")
args = ap.parse_args()
E.set_precision("fast")
torch.manual_seed(0)
model = E.yolo.Model(E.cfg_path(args.model)).cuda()
x = torch.rand(args.batch, 3, 640, 640, device="cuda")
for m in model.modules():
    if isinstance(m, torch.nn.BatchNorm3d):
        m.momentum = 1.0
model.train()
with torch.no_grad():
    model(x)
model.eval()

POOL = [torch.cuda.Stream() for _ in range(8)]


def run(n):
    parts = x.chunk(n)
    streams = POOL[:n]
    cur = torch.cuda.current_stream()
    outs = [None] * n
    for s in streams:
        s.wait_stream(cur)
    for i, (p, s) in enumerate(zip(parts, streams)):
        with torch.cuda.stream(s):
            outs[i] = model(p)[0]
    for s in streams:
        cur.wait_stream(s)
    return torch.cat(outs)

with torch.no_grad():
    ref = model(x)[0]
    for n in (1, args.streams, 4):
        for _ in range(4):
            z = run(n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            z = run(n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"streams={n}: {ms:.2f} ms/step  {args.batch / ms * 1e3:.1f} img/s  max|z - ref| = {float((z - ref).abs().max()):.3e}", flush=True)
