"""Timing of the section-8f rows on the BASELINE shapes (batch 64): batched NMS, fused SGD + EMA step, event frames.
GPU legs: CUDA events after warm-up.  CPU legs: the oracle port with torchvision.ops.nms / torch.optim.SGD + a
ModelEMA-style loop -- the reference's own algorithm and libraries -- on this box's host cores."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import post_oracle as P
import seeded as S
E = importlib.import_module("ecs-yolo_b200")
res = {"cores": os.cpu_count()}


def gpu_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def cpu_ms(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) * 1e3 / reps


# ---- NMS: decoded Detect output of resnet34 @640 (6000 rows, nc = 13), batch 64
pred = S.nms_inputs(dict(N=64, R=6000, nc=13, seed=777))
pd = pred.cuda()
for tag, kw in [("detect(conf .25, iou .45)", dict(conf_thres=0.25, iou_thres=0.45)),
                ("val(conf .001, iou .6, multi_label)", dict(conf_thres=0.001, iou_thres=0.6, multi_label=True))]:
    g = gpu_ms(lambda: E.general.nms_padded(pd, **kw))
    glist = gpu_ms(lambda: E.general.non_max_suppression(pd, **kw))
    eager = gpu_ms(lambda: P.non_max_suppression(pd, use_torchvision=True, **kw), reps=2, warm=1)
    c = cpu_ms(lambda: P.non_max_suppression(pred, use_torchvision=True, **kw), reps=1)
    cnt = E.general.nms_padded(pd, **kw)[1]
    res["nms " + tag] = dict(ours_ms=round(g, 3), ours_list_api_ms=round(glist, 3), reference_algorithm_on_gpu_ms=round(eager, 2),
                             reference_algorithm_on_cpu_ms=round(c, 1), images=64, kept_mean=float(cnt.float().mean()))
    print(tag, res["nms " + tag], flush=True)

# ---- optimizer + EMA step on the resnet34 model
m = E.yolo.Model(E.cfg_path("resnet34")).cuda()
for p_ in m.parameters():
    p_.grad = torch.randn_like(p_) * 0.01
opt = E.optim.SGDNesterovEMA(m, lr=1e-3, momentum=0.937, weight_decay=5e-4)
fused = gpu_ms(lambda: opt.step())
t0 = time.perf_counter(); [opt.step() for _ in range(10)]; torch.cuda.synchronize(); fused_wall = (time.perf_counter() - t0) * 100
g0, g1, g2 = E.optim.param_groups_of(m)
ref = torch.optim.SGD(g0, lr=1e-3, momentum=0.937, nesterov=True)
ref.add_param_group({"params": g1, "weight_decay": 5e-4}); ref.add_param_group({"params": g2})
import copy
ema = copy.deepcopy(m).eval()


def ref_step():
    ref.step()
    with torch.no_grad():
        msd = m.state_dict()
        for k, v in ema.state_dict().items():
            if v.dtype.is_floating_point:
                v *= 0.999
                v += (1 - 0.999) * msd[k].detach()


eager = gpu_ms(ref_step)
t0 = time.perf_counter(); [ref_step() for _ in range(10)]; torch.cuda.synchronize(); eager_wall = (time.perf_counter() - t0) * 100
n_par = sum(p_.numel() for p_ in m.parameters())
res["sgd_ema_step resnet34"] = dict(ours_gpu_ms=round(fused, 3), ours_wall_ms=round(fused_wall, 3), tensors=len(opt._vals),
                                    torch_sgd_plus_ema_loop_gpu_ms=round(eager, 3), torch_wall_ms=round(eager_wall, 3),
                                    params=n_par, bytes_moved_GB=round(n_par * 4 * 7 / 1e9, 3),
                                    hbm_GBps=round(n_par * 4 * 7 / fused / 1e6, 1))
print(res["sgd_ema_step resnet34"], flush=True)

# ---- event frames: batch 64, T = 5, ~100k events per bin, 304x240 -> 320x320
N, T, per = 64, 5, 100000
g = torch.Generator(device="cuda").manual_seed(1)
n_ev = N * T * per
x = torch.randint(0, 304, (n_ev,), device="cuda", dtype=torch.int32, generator=g)
y = torch.randint(0, 240, (n_ev,), device="cuda", dtype=torch.int32, generator=g)
p = torch.randint(0, 2, (n_ev,), device="cuda", dtype=torch.int32, generator=g)
f = torch.arange(N * T, device="cuda", dtype=torch.int32).repeat_interleave(per)
ev = gpu_ms(lambda: E.events.event_frames(x, y, p, f, N, T, (320, 320)))
import numpy as np
try:
    import cv2
    xs, ys, ps = x[:per].cpu().numpy(), y[:per].cpu().numpy(), p[:per].cpu().numpy()

    def cpu_sample():     # one sample = T bins, as the reference's loader does per item
        img = 127 * np.ones((T, 240, 304, 3), dtype=np.uint8)
        for i in range(T):
            img[i, ys, xs, :] = 255 * ps[:, None]
        out = np.zeros([T, 320, 320, 3])
        for i in range(T):
            out[i] = cv2.resize(img[i], (320, 320))
        return np.transpose(out, [0, 3, 1, 2])
    c = cpu_ms(cpu_sample, reps=5)
except Exception as e:
    c = None
res["event_frames b64 T5"] = dict(ours_ms=round(ev, 3), events=n_ev, Gevents_per_s=round(n_ev / ev / 1e6, 2),
                                  reference_algorithm_cpu_ms_per_sample=None if c is None else round(c, 2),
                                  reference_algorithm_cpu_ms_per_batch_1core=None if c is None else round(c * N, 1))
print(res["event_frames b64 T5"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "post_bench.json"), "w"), indent=1)
