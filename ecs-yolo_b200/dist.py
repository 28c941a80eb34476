"""Multi-GPU plumbing: one process per GPU (torch.distributed over NCCL/NVLink; gloo on CPU for tests).

The hot path shards by batch (images are independent, reference train.py:368,419): inference needs no
data-path collective; training has ONE exchange step per optimizer step, the gradient all-reduce
(DDP in the reference, train.py:419).  ``allreduce_grads`` is the bucketed all-reduce used when the model
is not wrapped in DistributedDataParallel; bench.py wraps with DDP so the reduction overlaps the backward
kernels.  BN statistics stay per-GPU (reference default, train.py:814) unless the model went through
``torch.nn.SyncBatchNorm.convert_sync_batchnorm`` like the reference's ``--sync-bn`` (train.py:359-360): then every tdBN
exchanges one [2, C] all-reduce per direction (``sync_bn_stats`` / ``sync_bn_sums``).
"""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str = None) -> None:
    if dist.is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend)


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` independent units (images) for `rank`; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values: List[float], device="cpu") -> List[float]:
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values: List[float], device="cpu") -> List[float]:
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, average: bool = True) -> int:
    """Bucketed all-reduce of .grad (fp32).  Parameters without a gradient on this rank (the spread convs at
    T = 1 receive none) contribute zeros so every rank issues the same collectives.  Returns the
    number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    plist = [p for p in params if p.requires_grad]
    n_coll, i = 0, 0
    while i < len(plist):
        bucket, size = [], 0
        while i < len(plist) and (not bucket or size + plist[i].numel() * 4 <= bucket_bytes):
            bucket.append(plist[i])
            size += plist[i].numel() * 4
            i += 1
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= world
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        n_coll += 1
    return n_coll


def sync_bn_stats(mean: torch.Tensor, var: torch.Tensor):
    """SyncBatchNorm forward (train.py:359-360): per-rank (mean, biased variance) over equal-sized shards -> the statistics
    of the union, through ONE all-reduce of [E x, E x^2].  -> (mean, var, world)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return mean, var, 1
    world = dist.get_world_size()
    st = torch.stack([mean, var + mean * mean])
    dist.all_reduce(st, op=dist.ReduceOp.SUM)
    st /= world
    m = st[0].contiguous()
    return m, (st[1] - m * m).clamp_min_(0).contiguous(), world


def sync_bn_sums(sg: torch.Tensor, sgy: torch.Tensor):
    """SyncBatchNorm backward: the batch sums of the input gradient (sum g, sum g*y) over ALL ranks (the weight / bias
    gradients keep the local sums: DDP averages those like every other parameter gradient).  -> (sg, sgy, world)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sg, sgy, 1
    st = torch.stack([sg, sgy])
    dist.all_reduce(st, op=dist.ReduceOp.SUM)
    return st[0].contiguous(), st[1].contiguous(), dist.get_world_size()
