"""Micro-benchmark of the depth-wise spread kernel (version 1 = global byte loads, 2 = shared-memory tiles) on the
layer shapes of resnet34 at batch 64.  Output bytes (bf16 rows, 2 B/elem) over the measured HBM peak is the roofline."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
F = E.functional
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
F.set_precision("fast")
for C, H in [(64, 160), (128, 80), (256, 40), (512, 20), (1024, 20), (384, 40), (64, 320)]:
    N = args.N
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (1, N, H, H, C // 32), device="cuda", dtype=torch.int32)
    bits &= torch.randint(-2 ** 31, 2 ** 31 - 1, bits.shape, device="cuda", dtype=torch.int32)
    bits &= torch.randint(-2 ** 31, 2 ** 31 - 1, bits.shape, device="cuda", dtype=torch.int32)   # ~12.5 % firing
    sp = F.Spikes(bits, C)
    w = F.make_lif_w(torch.randn(C, 1, 3, 3, device="cuda") * 0.3, torch.zeros(C, device="cuda"),
                     torch.randn(C, C, 1, 1, device="cuda") / C ** 0.5, torch.zeros(C, device="cuda"))
    out = {}
    for v in (1, 2):
        for _ in range(3):
            F.spread_dw(sp, 0, w, version=v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            F.spread_dw(sp, 0, w, version=v)
        e1.record()
        torch.cuda.synchronize()
        out[v] = e0.elapsed_time(e1) / args.reps
    elems = N * H * H * C
    print(f"C{C}@{H} N{N}: v1 {out[1]:.3f} ms  v2 {out[2]:.3f} ms  ({elems * 2 / out[2] / 1e6:.0f} GB/s written, "
          f"{elems / out[2] / 1e6:.1f} Gelem/s)", flush=True)
