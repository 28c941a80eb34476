"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, max-over-ranks timing and the
bucketed gradient all-reduce (equal to the single-process sum, missing gradients handled)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from util import ecsy


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    D = importlib.import_module("ecs-yolo_b200").dist
    D.init("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    net[2].weight.requires_grad_(True)
    lo, hi = D.shard_range(10, rank, world)
    x = torch.arange(10 * 8, dtype=torch.float32).reshape(10, 8)[lo:hi] / 80.0
    y = net[1](net[0](x)).sum()          # net[2] gets no gradient on any rank
    y.backward()
    ncoll = D.allreduce_grads(net.parameters(), bucket_bytes=256, average=False)
    tmax = D.max_over_ranks([float(rank + 1), 5.0 - rank])
    out[rank] = dict(g0=net[0].weight.grad.clone(), g2=net[2].weight.grad.clone(), ncoll=ncoll, tmax=tmax, lo=lo, hi=hi)
    torch.distributed.destroy_process_group()


def test_gloo_world2():
    ecsy()
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 5, 5, 10)
    # single-process reference: full batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    x = torch.arange(10 * 8, dtype=torch.float32).reshape(10, 8) / 80.0
    net[1](net[0](x)).sum().backward()
    assert torch.allclose(r0["g0"], net[0].weight.grad, atol=1e-5) and torch.equal(r0["g0"], r1["g0"])
    assert torch.equal(r0["g2"], torch.zeros(4, 4))           # missing gradients all-reduced as zeros
    assert r0["ncoll"] == r1["ncoll"] and r0["ncoll"] >= 2   # several buckets, same count on every rank
    assert r0["tmax"] == [2.0, 5.0] and r1["tmax"] == [2.0, 5.0]


def test_shard_range_covers_everything():
    D = ecsy().dist
    for total in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _sync_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    D = importlib.import_module("ecs-yolo_b200").dist
    D.init("gloo")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2 * 50, 6, generator=g) * 2 + 1          # rows of both ranks
    gg = torch.randn(2 * 50, 6, generator=g)
    xs, gs = x[rank * 50:(rank + 1) * 50], gg[rank * 50:(rank + 1) * 50]
    mean, var, w = D.sync_bn_stats(xs.mean(0), xs.var(0, unbiased=False))
    sg, sgy, w2 = D.sync_bn_sums(gs.sum(0), (gs * xs).sum(0))
    out[rank] = dict(mean=mean, var=var, sg=sg, sgy=sgy, w=(w, w2))
    torch.distributed.destroy_process_group()


def test_sync_bn_statistics_world2():
    """--sync-bn (train.py:359-360): the [2, C] all-reduces of a tdBN give the statistics / gradient sums of the UNION of the
    ranks' shards; a single process gets its inputs back."""
    D = ecsy().dist
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sync_worker, args=(world, port, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(100, 6, generator=g) * 2 + 1
    gg = torch.randn(100, 6, generator=g)
    for r in (0, 1):
        assert out[r]["w"] == (2, 2)
        assert torch.allclose(out[r]["mean"], x.mean(0), atol=1e-6)
        assert torch.allclose(out[r]["var"], x.var(0, unbiased=False), rtol=1e-5, atol=1e-6)
        assert torch.allclose(out[r]["sg"], gg.sum(0), atol=1e-5)
        assert torch.allclose(out[r]["sgy"], (gg * x).sum(0), rtol=1e-5, atol=1e-5)
    m, v = torch.ones(3), torch.full((3,), 2.0)
    assert D.sync_bn_stats(m, v) == (m, v, 1) or D.sync_bn_stats(m, v)[2] == 1


def test_param_groups_follow_reference_with_and_without_sync_bn():
    """train.py:262-277: BatchNorm3d weights -> g0 (no decay), other weights -> g1 (decay), biases -> g2.  The reference
    builds the groups BEFORE --sync-bn converts the model (train.py:283 vs :359), so the tdBN weights stay in g0; the
    package's optimizer gives the same split whether it is built before or after the conversion."""
    E = ecsy()
    blk = E.common.BasicBlock_2(64, 128, 3, 2)
    ref = dict(g0=[], g1=[], g2=[])
    for v in blk.modules():                      # the reference's loop, statement for statement
        if hasattr(v, 'bias') and isinstance(v.bias, torch.nn.Parameter):
            ref["g2"].append(v.bias)
        if isinstance(v, (torch.nn.BatchNorm3d)):
            ref["g0"].append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, torch.nn.Parameter):
            ref["g1"].append(v.weight)
    ids = lambda ps: sorted(id(p) for p in ps)
    g0, g1, g2 = E.optim.param_groups_of(blk)
    assert (ids(g0), ids(g1), ids(g2)) == (ids(ref["g0"]), ids(ref["g1"]), ids(ref["g2"]))
    assert len(g0) == 3 and all(p.dim() == 1 for p in g0)
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(blk)
    h0, h1, h2 = E.optim.param_groups_of(conv)
    assert (ids(h0), ids(h1), ids(h2)) == (ids(g0), ids(g1), ids(g2))     # the conversion keeps the Parameter objects
