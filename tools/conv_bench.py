"""Micro-benchmark of the spike implicit-GEMM conv on the resnet34 layer shapes (batch 64, T=4).
CUDA events, 3 warm-ups, inputs ~ GB-scale activations (>> L2 for the big layers).
    python tools/conv_bench.py [--mode fast|parity] [--reps 5] [--only IDX]"""
import argparse, importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
F = E.functional
ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fast")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--N", type=int, default=64)
ap.add_argument("--res", action="store_true", help="with a residual tensor added in the epilogue (block shortcut)")
ap.add_argument("--ts", default="auto", help="auto | all (spike operand in tensor memory) | off (shared memory)")
args = ap.parse_args()
F.set_precision(args.mode)
F.set_conv_ts(args.ts)
T, N = 4, args.N
# (cin, cout, k, stride, H)
SHAPES = [(64, 64, 3, 1, 160), (128, 128, 3, 1, 80), (256, 256, 3, 1, 40), (512, 512, 3, 1, 20),
          (64, 64, 3, 2, 320), (128, 256, 3, 2, 80), (512, 1024, 3, 1, 20), (1024, 256, 3, 1, 20),
          (64, 128, 1, 1, 80), (384, 256, 3, 1, 40)]
res = []
for idx, (ci, co, k, s, H) in enumerate(SHAPES):
    if args.only >= 0 and idx != args.only:
        continue
    g = torch.Generator(device="cuda").manual_seed(idx)
    bits = torch.randint(-2**31, 2**31 - 1, (T, N, H, H, ci // 32), device="cuda", dtype=torch.int32, generator=g)
    bits &= torch.randint(-2**31, 2**31 - 1, bits.shape, device="cuda", dtype=torch.int32, generator=g)
    bits &= torch.randint(-2**31, 2**31 - 1, bits.shape, device="cuda", dtype=torch.int32, generator=g)  # ~12.5 % ones
    w = torch.randn(co, ci, k, k, device="cuda") * 0.05
    cw = F.make_conv_w(w, None, s, k // 2, 1, True, False)
    sp = F.Spikes(bits, ci)
    sc = torch.rand(co, device="cuda") + 0.5
    sh = torch.rand(co, device="cuda")
    resid = None
    if args.res and s == 1:
        resid = F.Act(torch.randn(T, N, H, H, co, device="cuda"), T)
    for _ in range(3):
        out = F.spike_conv(sp, cw, sc, sh, resid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = F.spike_conv(sp, cw, sc, sh, resid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    Ho = out.H
    fl = 2.0 * T * N * Ho * Ho * co * ci * k * k
    r = dict(shape=f"{ci}->{co} k{k} s{s} @{H}", ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))
    res.append(r)
    print(r, flush=True)
    del bits, out
print(json.dumps({"mode": args.mode, "ts": args.ts, "results": res}))
