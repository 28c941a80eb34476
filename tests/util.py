import importlib
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ecsy():
    return importlib.import_module("ecs-yolo_b200")


def load_golden(name):
    import seeded as S
    return torch.load(os.path.join(S.GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def agree(a, b):
    return float((a.cpu() == b.cpu()).float().mean())


def nhwc(x):
    """[T,N,C,H,W] cpu -> cuda NHWC Act-ready view (reference-shaped, NHWC memory)."""
    return x.cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
