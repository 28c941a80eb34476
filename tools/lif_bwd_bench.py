"""Microbenchmark of ecsy_lif_ecs_bwd (stored state) on the resnet18 / batch-32 layer shapes; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split, or plain for CUDA-event totals."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
E = importlib.import_module("ecs-yolo_b200")
F_ = importlib.import_module("ecs-yolo_b200.functional")
torch.manual_seed(0)
F_.set_precision(os.environ.get("PRECISION", "fast"))
dev = "cuda"
T, N = 4, int(os.environ.get("N", 32))
iters = int(os.environ.get("ITERS", 5))
SHAPES = [(64, 160), (128, 80), (256, 40), (512, 20)]
if os.environ.get("SHAPE"):
    SHAPES = [SHAPES[int(os.environ["SHAPE"])]]
for C, H in SHAPES:
    x = F_.Act(torch.randn(T, N, H, H, C, device=dev) * 0.6 + 0.2, T)
    dw_w = torch.randn(C, 1, 3, 3, device=dev) * 0.2
    dw_b = torch.randn(C, device=dev) * 0.05
    pw_w = torch.randn(C, C, 1, 1, device=dev) * (1.0 / C ** 0.5)
    pw_b = torch.randn(C, device=dev) * 0.05
    w = F_.make_lif_w(dw_w, dw_b, pw_w, pw_b, C)
    saved = F_.lif_ecs(x, w, None, save_mem=True)
    g = torch.randn(T, N, H, H, C, device=dev)
    F_.lif_ecs_bwd(g, x, w, pw_w, saved=saved)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        F_.lif_ecs_bwd(g, x, w, pw_w, saved=saved)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    el = T * N * H * H * C
    print(f"C={C:4d} {H}x{H} N={N}: {ms:7.3f} ms  ({el / ms / 1e6:.1f} G elem-steps/s)")
