"""2-GPU check of the optional SyncBN path (train.py:359-360):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/syncbn_check.py
Every rank runs a BasicBlock_2 converted by SyncBatchNorm.convert_sync_batchnorm on ITS half of a seeded batch (train mode,
parity precision, forward + backward of a fixed linear loss); rank 0 also runs the unconverted block on the WHOLE batch.  With
equal shards the union statistics are the full-batch statistics, so: outputs and input gradients of the shards = the halves of
the full-batch run, running statistics equal, and the SUM over ranks of the local parameter gradients = the full-batch ones."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
E = importlib.import_module("ecs-yolo_b200")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
E.dist.init("nccl")
assert world == 2 and dist.is_initialized(), "run with two ranks"
E.set_precision("parity")
T, N, C, H = 4, 4, 64, 16
torch.manual_seed(5)
blk = E.common.BasicBlock_2(C, C, 3, 1)
for p in blk.parameters():
    if p.dim() == 1:
        p.data.add_(torch.randn_like(p) * 0.05)
full = E.common.BasicBlock_2(C, C, 3, 1)
full.load_state_dict(blk.state_dict())
x = torch.randn(T, N, C, H, H, generator=torch.Generator().manual_seed(11)) * 0.6 + 0.3
gout = torch.randn(T, N, C, H, H, generator=torch.Generator().manual_seed(12))
sync = torch.nn.SyncBatchNorm.convert_sync_batchnorm(blk).cuda().train()
lo, hi = rank * N // 2, (rank + 1) * N // 2
xs = x[:, lo:hi].cuda().requires_grad_(True)
ys = sync(xs)
(ys * gout[:, lo:hi].cuda()).sum().backward()
pg = {k: p.grad.detach().clone() for k, p in sync.named_parameters() if p.grad is not None}
for g in pg.values():
    dist.all_reduce(g)          # sum of the local parameter gradients (DDP would divide by the world size)
ok = True
if rank == 0:
    full = full.cuda().train()
    xf = x.cuda().requires_grad_(True)
    yf = full(xf)
    (yf * gout.cuda()).sum().backward()

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))
    e_out = rel(ys.detach(), yf.detach()[:, lo:hi])
    e_gx = rel(xs.grad, xf.grad[:, lo:hi])
    fp = dict(full.named_parameters())
    e_p = {k: rel(g, fp[k].grad) for k, g in pg.items()}
    fs, ss = full.state_dict(), sync.state_dict()
    e_rs = max(rel(ss[k].float(), fs[k].float()) for k in fs if "running" in k)
    worst = max(e_p, key=e_p.get)
    print(f"syncbn 2-rank vs full batch: out {e_out:.2e}  gx {e_gx:.2e}  running stats {e_rs:.2e}  worst param grad {worst} {e_p[worst]:.2e}", flush=True)
    ok = e_out < 1e-4 and e_gx < 1e-3 and e_rs < 1e-5 and e_p[worst] < 1e-3
    print("SYNCBN_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
