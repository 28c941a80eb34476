"""Microbenchmark of the image stem (3 -> 64, 7x7 / 2, batch N at 640x640, fast precision): ecsy_stem_conv vs the generic
im2col-on-the-fly tcgen05 path (ECSY_STEM_KERNEL=0)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
E = importlib.import_module("ecs-yolo_b200")
F_ = E.functional
F_.set_precision("fast")
N = int(os.environ.get("N", 64))
x = F_.Act.from_ref(torch.rand(1, N, 3, 640, 640, device="cuda"))
w = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
cw = F_.make_conv_w(w, None, 2, 3, 1, True, True)
sc, sh = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
for _ in range(3):
    y = F_.real_conv(x, cw, sc, sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = F_.real_conv(x, cw, sc, sh)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"stem conv N={N}: {ms:.3f} ms  ({y.data.numel() * 4 / ms / 1e6:.0f} GB/s of output)")
