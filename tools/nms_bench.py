"""Timing of ecsy_nms at the BASELINE shape (batch 64, 6000 rows, nc = 13) in the detect and val settings."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import seeded as S
E = importlib.import_module("ecs-yolo_b200")
pd = S.nms_inputs(dict(N=64, R=6000, nc=13, seed=777)).cuda()
for tag, kw in [("detect(conf .25, iou .45)", dict(conf_thres=0.25, iou_thres=0.45)),
                ("val(conf .001, iou .6, multi_label)", dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
                ("val, max_det 3000", dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=3000))]:
    for _ in range(3):
        E.general.nms_padded(pd, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out, cnt = E.general.nms_padded(pd, **kw)
    b.record()
    torch.cuda.synchronize()
    print(f"nms {tag}: {a.elapsed_time(b) / 10:.3f} ms per batch of 64, kept mean {float(cnt.float().mean()):.1f}", flush=True)
