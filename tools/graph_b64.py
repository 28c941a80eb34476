"""Experiment: the batch-64 eval forward eager vs replayed from a CUDA graph (ecs-yolo_b200/graph.py).
python tools/graph_b64.py [--batch 64] [--model resnet34] [--steps 10]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
E = importlib.import_module("ecs-yolo_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--model", default="resnet34")
ap.add_argument("--steps", type=int, default=10)
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


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


with torch.no_grad():
    t_eager = timed(lambda: model(x))
    print(f"eager  {t_eager:.2f} ms/step  {args.batch / t_eager * 1e3:.1f} img/s  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    gf = E.graph.GraphedForward(model, x)
    t_graph = timed(lambda: gf(x))
    print(f"graph  {t_graph:.2f} ms/step  {args.batch / t_graph * 1e3:.1f} img/s  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB  launches {gf.launches_per_replay}", flush=True)
    t_eager2 = timed(lambda: model(x))
    print(f"eager  {t_eager2:.2f} ms/step (again)")
