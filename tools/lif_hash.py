"""Prints a digest of the fast-precision per-timestep LIF pipeline's outputs (spikes, membranes, traces) on fixed seeded
inputs; tests/test_gpu_ops.py runs it with ECSY_ECS_GEMM=0 and =1 and compares (fused GEMM-epilogue step == two kernels)."""
import hashlib, importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
E = importlib.import_module("ecs-yolo_b200")
F_ = E.functional
F_.set_precision("fast")
F_.set_lif_wave("off")
h = hashlib.sha256()
for (C, H, W, N, T, aff, save) in [(64, 24, 40, 2, 4, False, False), (128, 17, 23, 3, 4, True, False), (256, 20, 20, 2, 5, False, True),
                                   (512, 10, 13, 2, 4, True, True), (192, 16, 16, 1, 4, False, False)]:
    g = torch.Generator().manual_seed(C + H)
    x = F_.Act((torch.randn(T, N, H, W, C, generator=g) * 0.6 + 0.2).cuda(), T)
    dw_w, dw_b = (torch.randn(C, 1, 3, 3, generator=g) * 0.2).cuda(), (torch.randn(C, generator=g) * 0.05).cuda()
    pw_w, pw_b = (torch.randn(C, C, 1, 1, generator=g) / C ** 0.5).cuda(), (torch.randn(C, generator=g) * 0.05).cuda()
    w = F_.make_lif_w(dw_w, dw_b, pw_w, pw_b, C)
    a = ((torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()) if aff else None
    out = F_.lif_ecs(x, w, a, save_mem=save, allow_wave=False)
    ts = [out[0].bits, out[1], out[2]] if save else [out.bits]
    for t in ts:
        h.update(t.cpu().numpy().tobytes())
    print(C, H, W, float(ts[0].float().abs().mean()), file=sys.stderr)
print(h.hexdigest())
