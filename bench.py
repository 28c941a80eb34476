#!/usr/bin/env python
"""bench.py -- SNN-YOLO images/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3                      # our CUDA path
    python bench.py --impl reference --steps 2 --warmup 1              # reference algorithm on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- EMS-ResNet34 SNN-YOLO (cfg/resnet34.yaml), T=4,
inference, per-GPU batch 64, synthetic 640x640 images, default-init weights with momentum-1 calibrated
tdBN, bf16 tensor-core operands with fp32 accumulation.  A "step" is one forward
pass over one batch.  `value`: inputs resident in HBM; `e2e`: Model.forward called with a pinned HOST
batch, H2D copy and D2H read of the decoded detections inside the timed region.  Batch-sharded over
ranks with no data-path collective (weak scaling).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import yaml  # noqa: E402


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 7 and r[3 + j] == "Active" for r in self.rows)]
        mx = max(float(r[1]) for r in self.rows if len(r) >= 7)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def measured_traffic(model, batch, T, precision):
    """{"lif": {...}, "conv": {...}}: DRAM bytes per step of the two dominant operators from the committed ncu capture
    (profiles/traffic.json, raw CSV beside it), or None when no capture exists for this exact configuration."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(f"{model}|{batch}|{T}|{precision}")


def parity_record(args, parity_ips):
    """What the benchmarked precision was VALIDATED at: the teacher-forced per-layer figures of the BASELINE plans measured
    on a B200 by tests/test_gpu_baseline_cfgs.py (committed summary: profiles/parity_summary.json), and the throughput of
    the same step in `parity` precision (bf16 hi + lo weight planes, fp32 state)."""
    rec = {"precision": args.precision,
           "gates": "tests/test_gpu_baseline_cfgs.py (resnet10 / resnet34 / resnet18 plans at 640x640, every layer "
                    "teacher-forced on the oracle's input, both precisions): spikes >= 99.9 % per neuron, real tensors, head "
                    "and loss <= 1e-3"}
    path = os.path.join(ROOT, "profiles", "parity_summary.json")
    if os.path.exists(path):
        summ = json.load(open(path))
        rec["measured"] = summ.get(f"{args.model}|{args.precision}") or summ
    if parity_ips is not None:
        rec["parity_precision_images_per_s"] = parity_ips
    return rec


def load_cfg(name):
    return yaml.safe_load(open(os.path.join(ROOT, "ecs-yolo_b200", "cfg", name + ".yaml")))


def event_frames(n, T, img, gen):
    """Gen1-style event frames (g1-resnet/utils/give_g1_data.py:550-565): ternary {0, 127, 255}/255 frames
    on a grey background, ~3 % event pixels, three identical channels -> [T, n, 3, img, img] for _forward_once."""
    u = torch.rand(n, T, 1, img, img, generator=gen)
    f = torch.full_like(u, 127.0 / 255.0)
    f[u < 0.015] = 0.0
    f[u > 0.985] = 1.0
    return f.expand(-1, -1, 3, -1, -1).permute(1, 0, 2, 3, 4).contiguous()


def cpu_reference(model_name, T, img, sample_imgs, steps, warmup, threads, events=False, device="cpu", autocast=False):
    """Times the oracle port of the reference's CPU path (torch fp32, all host threads) on a bounded
    sample of the same workload: `sample_imgs` images per step.  device="cuda" (`--impl reference --ref-device cuda`)
    runs the same stock-PyTorch eager code on the GPU instead: the same-box GPU comparison of SURVEY 8(d)."""
    import ecs_oracle as O
    torch.set_num_threads(threads)
    cfg = load_cfg(model_name)
    if events:
        cfg["nc"] = 2
    sd = O.init_state_dict(cfg, T, seed=0)
    stride = O.detect_strides(cfg)
    for k in sd:
        if k.endswith("anchors"):
            sd[k] = sd[k] / stride.view(-1, 1, 1)
    g = torch.Generator().manual_seed(0)
    x = event_frames(sample_imgs, T, img, g) if events else torch.rand(sample_imgs, 3, img, img, generator=g)
    on_gpu = device != "cpu"
    if on_gpu:
        sd = {k: v.to(device) for k, v in sd.items()}
        x, stride = x.to(device), stride.to(device)

    def sync():
        if on_gpu:
            torch.cuda.synchronize()
    amp = torch.autocast("cuda", dtype=torch.float16, enabled=bool(autocast and on_gpu))   # the reference's AMP (train.py:546)
    with torch.no_grad(), amp:
        O.BN_MOMENTUM = 1.0
        O.forward(cfg, sd, x, T, True, stride=stride)   # calibration pass (train-mode tdBN) == warm-up 0
        O.BN_MOMENTUM = 0.1
        for _ in range(max(warmup - 1, 0)):
            O.forward(cfg, sd, x, T, False, stride=stride)
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(cfg, sd, x, T, False, stride=stride)
        sync()
        dt = time.perf_counter() - t0
    return sample_imgs * steps / dt, dt / steps


def train_bench(args, E, F, model, x, x_host, rank, world, local, dist, workload, emit_line=True):
    """fwd + loss + bwd + SGD-nesterov step per batch (train.py:555-582 shape of a step).  Returns the JSON record (rank 0;
    None elsewhere) and, with emit_line, prints it as the bench line (`--mode train`).  Stack-A models (Detect head)
    train against the reference's ComputeLoss (utils/loss.py:130-234: SIoU + BCE, build_targets), DDetect models (Stack B)
    against utils/loss_tal.py:105-215 (TaskAlignedAssigner + box + DFL + BCE), both on COCO-shaped synthetic targets and
    computed on the device (ecsy_yolo_loss / ecsy_tal_loss); `--loss quadratic` is a synthetic quadratic
    on the raw head outputs.  The JSON line says which."""
    model.train()
    net = model
    if dist is not None:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=25,
                                                        gradient_as_bucket_view=True)
    if args.optim == "fused":
        # the reference's three parameter groups + ModelEMA in one multi-tensor launch per step
        opt = E.optim.SGDNesterovEMA(model, lr=1e-3, momentum=0.937, weight_decay=5e-4)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.937, nesterov=True)
    with torch.no_grad():
        probe = model(x[:, :1] if args.events else x[:1])
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    det = model.model[-1]
    use_yolo_loss = args.loss == "yolo" and type(det).__name__ in ("Detect", "DDetect")
    if use_yolo_loss:
        # hyp.scratch.yaml gains with train.py:427-433's scaling to nl levels / nc classes / the image size
        nl, nc = det.nl, det.nc
        model.hyp = dict(box=0.05 * 3 / nl, cls=0.5 * nc / 80 * 3 / nl, obj=1.0 * (args.img / 640) ** 2 * 3 / nl,
                         cls_pw=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, slide_ratio=0.0, label_smoothing=0.0)
        stack_a = type(det).__name__ == "Detect"
        compute_loss = E.loss.ComputeLoss(model) if stack_a else E.loss_tal.ComputeLoss(model)
        # COCO-shaped labels: 5-8 boxes per image, class ~ U{0..nc-1}, centre ~ U(.2,.8), size ~ U(.05,.35)
        gc = torch.Generator().manual_seed(1 + rank)
        per = torch.randint(5, 9, (args.batch,), generator=gc)
        img = torch.repeat_interleave(torch.arange(args.batch), per).float()
        n_t = int(img.numel())
        tgt_host = torch.cat([img[:, None], torch.randint(0, nc, (n_t, 1), generator=gc).float(),
                              torch.rand(n_t, 2, generator=gc) * 0.6 + 0.2,
                              torch.rand(n_t, 2, generator=gc) * 0.3 + 0.05], 1).pin_memory()
        tgt = tgt_host.cuda()
        loss_name = ((f"ComputeLoss (utils/loss.py:130-234: SIoU box + BCE obj / cls, build_targets) on the device "
                      f"(ecsy_yolo_loss)") if stack_a else
                     (f"ComputeLoss (utils/loss_tal.py:105-215: TaskAlignedAssigner + box + DFL + BCE) on the device "
                      f"(ecsy_tal_loss)")) + f", {n_t} synthetic COCO-shaped boxes per batch"
    else:
        tgt = [torch.randn(args.batch, *o.shape[1:], device="cuda", generator=g) for o in probe]
        tgt_host = None
        loss_name = "synthetic quadratic on the raw head outputs"

    def step(inp, targets=None):
        opt.zero_grad(set_to_none=True)
        out = net(inp)
        if use_yolo_loss:
            loss, _items = compute_loss(out, tgt if targets is None else targets)
        else:
            loss = sum(((o - t) ** 2).mean() for o, t in zip(out, tgt))
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, args.min_warmup)):
        step(x)
    barrier()
    F.launches["n"] = 0
    F.launches_by_op.clear()
    for k in F.flops:
        F.flops[k] = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            step(x)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = F.launches["n"]
    launch_by_op = dict(F.launches_by_op)
    flops = dict(F.flops)
    # per-operator breakdown from ONE extra step outside the timed region (two CUDA events per operator cost host
    # time, and the training step is close to launch-bound)
    F.profile_begin()
    step(x)
    barrier()
    per_op = {k: v * args.steps for k, v in F.profile_end().items()}
    barrier()
    # the exchange step (train.py:419 DDP, :559-567): the same steps WITHOUT the gradient all-reduce (no_sync) -> what of the
    # communication is not hidden behind the backward kernels
    comm = None
    if dist is not None:
        n_sync = min(args.steps, 5)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(n_sync):
            step(x)
        e1.record()
        with net.no_sync():
            for _ in range(n_sync):
                step(x)
        e2.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / n_sync, e1.elapsed_time(e2) / n_sync], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gbytes = sum(p.numel() * 4 for p in model.parameters() if p.requires_grad)
        try:
            ld = net._get_ddp_logging_data()   # after the first iteration DDP re-buckets in gradient-ready order
            raw = ld.get("rebuilt_bucket_sizes") or ld.get("bucket_sizes", "")
            bucket_sizes = [int(v) for v in str(raw).split(",") if v.strip()]
        except Exception:
            bucket_sizes = []
        comm = {"collective": "NCCL all-reduce (sum) of the fp32 gradients through DistributedDataParallel, overlapped with "
                              "the backward kernels (train.py:419, :559-567)",
                "allreduce_bytes_per_step": gbytes, "bucket_cap_mb": 25, "buckets": len(bucket_sizes) or None,
                "largest_bucket_bytes": max(bucket_sizes) if bucket_sizes else None,
                "ms_per_step_with_allreduce": float(tt[0]), "ms_per_step_no_sync": float(tt[1]),
                "exposed_comm_ms_per_step": float(tt[0] - tt[1]),
                "exposed_note": "step time with the all-reduce minus the same step under no_sync(); what stays exposed is the "
                                "last bucket(s): the gradients of the first layers are ready only when the backward ends"}
        opt.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    loss_host = 0.0
    for _ in range(0 if args.no_e2e else args.steps):
        xd = x_host.to("cuda", non_blocking=True)
        td = tgt_host.to("cuda", non_blocking=True) if tgt_host is not None else None
        loss_host = float(step(xd, td).detach().cpu())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms_total, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None and emit_line:
            dist.destroy_process_group()
        return None
    imgs = args.batch * world * args.steps
    pk = peaks()
    conv_ms = (per_op.get("spike_conv", 0.0) + per_op.get("conv_dgrad", 0.0) + per_op.get("conv_wgrad", 0.0)
               + per_op.get("conv_bwd", 0.0)) / args.steps
    conv_fl = (flops.get("spike_conv", 0.0) + flops.get("conv_bwd", 0.0)) / args.steps
    conv_tf = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    line = {
        "metric": "images/s", "value": imgs / (ms_total * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, args.min_warmup), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload.replace("inference", "training (fwd + loss + bwd + SGD step" + (" + EMA, fused)" if args.optim == "fused" else ")")),
                   "model_cfg": f"cfg/{args.model}.yaml", "T": args.T, "global_batch": args.batch * world,
                   "precision": args.precision, "accumulate": "fp32",
                   "parallelism": f"DDP x{world} (NCCL all-reduce of fp32 gradients, 25 MB buckets)" if world > 1
                                  else "single GPU",
                   "loss": loss_name,
                   "l2": "activations per step (GBs) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": (imgs / (e2e_ms * 1e-3)) if e2e_ms > 0 else None, "unit": "images/s",
                "h2d_bytes_per_step": x_host.numel() * 4 + (tgt_host.numel() * 4 if tgt_host is not None else 0),
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clk.summary(),
        "roofline": {"kernel": "tcgen05 conv kernels (spike conv fwd + dgrad + wgrad)", "bound": "tensor",
                     "achieved": conv_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": conv_tf / pk["tf_sust"],
                     "traffic": None, "peak_source": pk["src"] + " sustained bf16",
                     "algorithmic_gflop_per_step": conv_fl / 1e9, "kernel_ms_per_step": conv_ms},
        "breakdown_ms_per_step": {k: v / args.steps for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])},
        "last_loss": loss_host,
    }
    line["config"]["lif_state"] = ("membranes + ECS traces of the forward kept for the backward (set_lif_store, 8 B per "
                                   "element-step; falls back to recomputation per layer when memory is short)"
                                   if F._state["lif_store"] else "membranes recomputed in the backward")
    line["config"]["optimizer"] = ("fused SGD-Nesterov + EMA, one multi-tensor launch (ecsy_sgd_ema_step)" if args.optim == "fused"
                                   else "torch.optim.SGD (nesterov)")
    # the LIF backward against its algorithmic traffic (gout 4 + membrane 4 + gx 4 bytes per element-step)
    lif_elems = flops.get("lif_elems", 0.0) / args.steps
    lb_ms = per_op.get("lif_ecs_bwd", 0.0) / args.steps
    if lb_ms > 0 and lif_elems > 0:
        gbs = lif_elems * 12.0 / (lb_ms * 1e-3) / 1e9
        line["roofline_lif_bwd"] = {"kernel": "ecsy_lif_ecs_bwd (reverse scan + spread dgrad / wgrad)", "bound": "hbm",
                                    "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                    "traffic": None, "kernel_ms_per_step": lb_ms,
                                    "algorithmic_bytes_per_step": lif_elems * 12.0}
    # tdBN batch statistics (train mode): every conv output read once (4 B per element)
    bn_elems = flops.get("tdbn_elems", 0.0) / args.steps
    bn_ms = per_op.get("tdbn_stats", 0.0) / args.steps
    if bn_ms > 0 and bn_elems > 0:
        gbs = bn_elems * 4.0 / (bn_ms * 1e-3) / 1e9
        line["roofline_tdbn"] = {"kernel": "ecsy_tdbn_stats (k_bn_partial + k_bn_final: per-channel mean / variance)", "bound": "hbm",
                                 "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                                 "kernel_ms_per_step": bn_ms, "algorithmic_bytes_per_step": bn_elems * 4.0}
    if comm is not None:
        line["allreduce"] = comm
    line["launches_by_op_per_step"] = {k: v / args.steps for k, v in sorted(launch_by_op.items(), key=lambda kv: -kv[1])}
    if emit_line:
        emit(line)
        if dist is not None:
            dist.destroy_process_group()
    return line


_REAL_STDOUT = None


def protect_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, torch warnings) goes to stderr; the ONE JSON line is
    written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="resnet34")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--T", type=int, default=4)
    ap.add_argument("--precision", default="fast", choices=["fast", "parity"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-warmup", type=int, default=3, help="profiling runs under ncu may lower this")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--cpu-sample", type=int, default=4, help="images per CPU step")
    ap.add_argument("--ref-autocast", action="store_true", help="--ref-device cuda only: run the eager port under fp16 autocast")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the reference arm proper (host cores); cuda = the same stock-PyTorch "
                         "eager code on the GPU (extra same-box comparison, SURVEY 8d)")
    ap.add_argument("--events", action="store_true",
                    help="BASELINE config 4: Gen1-style event frames [T,N,3,H,W] straight into _forward_once, nc=2 "
                         "(use with --T 5)")
    ap.add_argument("--optim", default="fused", choices=["fused", "torch"],
                    help="training: fused SGD-Nesterov + EMA kernel (default) or torch.optim.SGD without EMA")
    ap.add_argument("--loss", default="yolo", choices=["yolo", "quadratic"],
                    help="training: the reference's ComputeLoss on the device (utils/loss.py for Detect models, utils/loss_tal.py for "
                         "DDetect models) or a synthetic quadratic on the raw head outputs")
    ap.add_argument("--no-train", action="store_true", help="infer mode: skip the `train` sub-record")
    ap.add_argument("--no-parity-leg", action="store_true", help="skip the parity-precision timing of the same step")
    ap.add_argument("--no-small-batch", action="store_true", help="skip the batch-1 eager / CUDA-graph record")
    ap.add_argument("--train-model", default="resnet18", help="model of the `train` sub-record (BASELINE configs[2])")
    ap.add_argument("--train-batch", type=int, default=32, help="images per GPU per training step")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: forward + loss + backward + SGD step (DDP gradient all-reduce when --gpus > 1)")
    args = ap.parse_args()
    protect_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    workload = f"EMS-{args.model} SNN-YOLO T={args.T} inference, batch {args.batch}/GPU, synthetic {args.img}x{args.img}"
    if args.events:
        workload += " Gen1-style event frames (ternary, nc=2)"

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(args.steps, 1)
        ips, spstep = cpu_reference(args.model, args.T, args.img, args.cpu_sample, steps, max(args.warmup, 1), threads,
                                    args.events, args.ref_device, args.ref_autocast)
        line = {"impl": "reference", "metric": "images/s", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": spstep * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "sample": f"{args.cpu_sample} image(s) per step"},
                "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                                 "sample": f"{args.cpu_sample} image(s)/step of the same workload, oracle port "
                                           "(torch fp32 CPU) of the reference forward"},
                "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if args.ref_device != "cpu":      # not the reference arm proper: the same eager code on the GPU, labelled as such
            line["config"]["device"] = ("cuda: stock PyTorch eager (oracle port, " +
                                        ("fp16 autocast" if args.ref_autocast else "fp32, cuDNN TF32 convolutions") +
                                        "), same-box GPU comparison")
            line["cpu_baseline"] = None
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = os.path.join(ROOT, "ecs-yolo_b200", "lib", "libecsy.so")
    if not os.path.exists(lib) and rank == 0:
        import __graft_entry__
        __graft_entry__.build()
    if dist is not None:
        dist.barrier()
    E = importlib.import_module("ecs-yolo_b200")
    F = E.functional
    E.set_precision(args.precision)
    E.common.time_window = args.T

    torch.manual_seed(0)
    is_b = any(row[2] == "DDetect" for row in load_cfg(args.model)["head"])
    model = (E.yolo_snn.DetectionModel if is_b else E.yolo.Model)(E.cfg_path(args.model),
                                                                  nc=2 if args.events else None).cuda()
    g = torch.Generator().manual_seed(1000 + rank)
    if args.events:
        x_host = event_frames(args.batch, args.T, args.img, g).pin_memory()
    else:
        x_host = torch.rand(args.batch, 3, args.img, args.img, generator=g).pin_memory()
    x = x_host.cuda()

    if args.mode == "train":
        return train_bench(args, E, F, model, x, x_host, rank, world, local, dist, workload)

    # momentum-1 calibration of every tdBN on the seed batch, then eval
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model(x)
    model.eval()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, args.min_warmup)):
            z, _ = model(x)
        barrier()
        F.launches["n"] = 0
        F.launches_by_op.clear()
        for k in F.flops:
            F.flops[k] = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            ev0.record()
            for _ in range(args.steps):
                z, _ = model(x)
            ev1.record()
            barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches = F.launches["n"]
        launch_by_op = dict(F.launches_by_op)
        flops = dict(F.flops)
        # per-operator CUDA-event breakdown (and the dominant kernel's duration for `roofline`) from ONE extra step
        # outside the timed region, scaled to `steps`
        F.profile_begin()
        model(x)
        barrier()
        per_op = {k: v * args.steps for k, v in F.profile_end().items()}

        # end to end through the public API: every step uploads ITS batch from pinned host memory and downloads
        # its decoded detections; the upload of batch i+1 runs on a copy stream while batch i computes.
        barrier()
        z_host = torch.empty(z.shape, dtype=z.dtype).pin_memory()
        copy_stream = torch.cuda.Stream()
        d2h_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        bufs = [torch.empty_like(x), torch.empty_like(x)]
        up = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        nsteps = 0 if args.no_e2e else args.steps
        t0 = time.perf_counter()
        if nsteps:
            with torch.cuda.stream(copy_stream):
                bufs[0].copy_(x_host, non_blocking=True)
                up[0].record(copy_stream)
        for i in range(nsteps):
            cur, nxt = i % 2, (i + 1) % 2
            main.wait_event(up[cur])
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(free[nxt])
                    bufs[nxt].copy_(x_host, non_blocking=True)
                    up[nxt].record(copy_stream)
            zz, _ = model(bufs[cur])
            free[cur].record(main)
            # the step's result goes back on its own stream, so the download overlaps the next step's kernels
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                z_host.copy_(zz, non_blocking=True)
                zz.record_stream(d2h_stream)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

    # the same step in `parity` precision (bf16 hi + lo weight planes, fp32 ECS state, tanhf): the precision whose
    # real-valued outputs match the fp32 reference to 1e-5; two timed steps
    parity_ips = None
    if args.precision != "parity" and not args.no_parity_leg:
        E.set_precision("parity")
        with torch.no_grad():
            model(x)
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(2):
                model(x)
            p1.record()
            barrier()
        parity_ips = args.batch * world * 2 / (p0.elapsed_time(p1) * 1e-3)
        E.set_precision(args.precision)

    # small-batch serving: one image per call, eager launches vs CUDA-graph replay of the same launch sequence
    small_rec = None
    if not args.no_small_batch and rank == 0:
        x1 = (x[:, :1] if x.dim() == 5 else x[:1]).contiguous()
        x1_host = (x_host[:, :1] if x_host.dim() == 5 else x_host[:1]).contiguous().pin_memory()

        def per_call_ms(fn, n=30):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        with torch.no_grad():
            eager_ms = per_call_ms(lambda: model(x1))
            gf = E.graph.GraphedForward(model, x1)
            graph_ms = per_call_ms(lambda: gf(x1))
            z1_host = torch.empty(gf.static_out[0].shape, dtype=torch.float32).pin_memory()

            def e2e_call():
                gf(x1_host)                                    # H2D copy of the image into the graph's input buffer + replay
                z1_host.copy_(gf.static_out[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()      # the caller needs the detections before the next request
            e2e1_ms = per_call_ms(e2e_call)
        small_rec = {"workload": f"EMS-{args.model} T={args.T} inference, batch 1, {args.img}x{args.img}, {args.precision}",
                     "eager_ms_per_image": eager_ms, "graph_ms_per_image": graph_ms, "graph_e2e_ms_per_image": e2e1_ms,
                     "images_per_s_eager": 1e3 / eager_ms, "images_per_s_graph": 1e3 / graph_ms,
                     "images_per_s_graph_e2e": 1e3 / e2e1_ms, "launches_per_replay": gf.launches_per_replay,
                     "note": "ecs-yolo_b200/graph.py GraphedForward: the eval forward captured once, replayed per request; "
                             "outputs bit-identical to eager (tests/test_gpu_graph.py); e2e = pinned-host image in, "
                             "decoded predictions out, synchronised per request"}
        del gf

    t = torch.tensor([ms_total, e2e_s * 1e3, parity_ips or 0.0], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t[:2], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[2:], op=dist.ReduceOp.MIN)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    parity_ips = float(t[2]) if parity_ips is not None else None

    # "fwd + train" half of the metric (BASELINE configs[2]): surrogate-gradient BPTT training of the train model with the
    # reference's loss on the device, fused optimizer, DDP gradient all-reduce when launched on several GPUs
    train_rec = None
    if not args.no_train:
        del model, x, bufs, z, zz
        torch.cuda.empty_cache()
        targs = argparse.Namespace(**vars(args))
        targs.model, targs.batch, targs.mode = args.train_model, args.train_batch, "train"
        targs.steps, targs.warmup, targs.no_e2e = min(args.steps, 10), 3, False
        torch.manual_seed(0)
        is_bt = any(row[2] == "DDetect" for row in load_cfg(targs.model)["head"])
        tmodel = (E.yolo_snn.DetectionModel if is_bt else E.yolo.Model)(E.cfg_path(targs.model)).cuda()
        gt = torch.Generator().manual_seed(2000 + rank)
        xt_host = torch.rand(targs.batch, 3, args.img, args.img, generator=gt).pin_memory()
        twork = f"EMS-{targs.model} SNN-YOLO T={args.T} inference, batch {targs.batch}/GPU, synthetic {args.img}x{args.img}"
        train_rec = train_bench(targs, E, F, tmodel, xt_host.cuda(), xt_host, rank, world, local, dist, twork, emit_line=False)
        if train_rec is not None:
            train_rec["images_per_s"] = train_rec.pop("value")
            for k in ("metric", "unit", "higher_is_better", "scaling", "vs_baseline", "data", "clocks"):
                train_rec.pop(k, None)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    imgs = args.batch * world * args.steps
    pk = peaks()
    conv_ms = per_op.get("spike_conv", 0.0) / args.steps
    conv_tf = flops["spike_conv"] / args.steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # ECS-LIF: the dominant operator of the step.  Algorithmic bytes: every input element once (4 B; a
    # T-broadcast input is ONE frame) + 1 bit of spike per element-step.
    lif_ms = per_op.get("lif_ecs", 0.0) / args.steps
    lif_bytes = (flops["lif_fwd_in_elems"] * 4.0 + flops["lif_fwd_elems"] / 8.0) / args.steps
    lif_bytes_full = flops["lif_fwd_elems"] * 4.125 / args.steps       # counting the T-broadcast frame T times
    lif_gbs = lif_bytes_full / (lif_ms * 1e-3) / 1e9 if lif_ms > 0 else 0.0
    tr = measured_traffic(args.model, args.batch, args.T, args.precision) or {}
    line = {
        "metric": "images/s", "value": imgs / (ms_total * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, args.min_warmup), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "model_cfg": f"cfg/{args.model}.yaml", "T": args.T,
                   "global_batch": args.batch * world, "precision": args.precision,
                   "accumulate": "fp32", "parallelism": f"batch-sharded x{world}, no data-path collective",
                   "l2": "activations per step (GBs) exceed the 126 MB L2; no explicit flush",
                   "weights": "default init (seed 0), tdBN calibrated with momentum 1.0 on the seed batch"},
        "e2e": {"value": imgs / (e2e_ms * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": z_host.numel() * 4},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        # the DOMINANT operator of the step: the ECS-LIF (wavefront kernel on the 64-channel layers, per-timestep pipeline
        # elsewhere), HBM-bound by design; the convolutions' tensor-pipe roofline follows in `roofline_conv`
        "roofline": {"kernel": "ecsy_lif_ecs_*: k_lif_ecs_wave64 (C = 64: all T steps on chip) + k_lif_first / k_spread_dw / "
                               "k_dense_tma_h / k_ecs_step (other widths, per timestep)", "bound": "hbm",
                     "achieved": lif_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": lif_gbs / pk["hbm"],
                     "traffic": tr.get("lif", {}).get("dram_bytes_per_step"),
                     "traffic_note": "DRAM bytes per step summed over the operator's launches (ncu, profiles/traffic.json)",
                     "peak_source": pk["src"] + " HBM copy bandwidth",
                     "algorithmic_bytes_per_step": lif_bytes_full,
                     "algorithmic_bytes_note": "BASELINE.md section 2 / SURVEY 8d: elements x (4 B of input current + 1/8 B of "
                                               "spike), 137.4 M elements per image for resnet34; the kernels read the "
                                               "T-broadcast stem output as ONE frame, which makes the bytes they must move "
                                               f"{lif_bytes:.4g}",
                     "kernel_ms_per_step": lif_ms, "launches_per_step": launch_by_op.get("lif_ecs", 0) / args.steps},
        "roofline_conv": {"kernel": "k_spike_conv_ts + k_umma_gemm<spikes> (spike implicit-GEMM convs, tcgen05; operand in "
                                    "tensor memory / shared memory)", "bound": "tensor",
                          "achieved": conv_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                          "frac": conv_tf / pk["tf_sust"],
                          "traffic": tr.get("conv", {}).get("dram_bytes_per_step"),
                          "peak_source": pk["src"] + " sustained bf16",
                          "algorithmic_gflop_per_step": flops["spike_conv"] / args.steps / 1e9,
                          "kernel_ms_per_step": conv_ms, "launches_per_step": launch_by_op.get("spike_conv", 0) / args.steps},
        "breakdown_ms_per_step": {k: v / args.steps for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])},
        "launches_by_op_per_step": {k: v / args.steps for k, v in sorted(launch_by_op.items(), key=lambda kv: -kv[1])},
        "dense_tflops_whole_step": (flops["spike_conv"] + flops["ecs_pw"] + flops["real_conv"]) / args.steps
                                   / (ms_total / args.steps * 1e-3) / 1e12,
    }
    line["parity"] = parity_record(args, parity_ips)
    if small_rec is not None:
        line["small_batch"] = small_rec
    if train_rec is not None:
        line["train"] = train_rec
    if world == 1 and not args.no_cpu_baseline:
        ips, _ = cpu_reference(args.model, args.T, args.img, args.cpu_sample, 3, 1, threads, args.events)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_sample} image(s)/step x 3 steps (+ 1 calibration step) of the same "
                                          "workload, oracle port (torch fp32 CPU) of the reference forward"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
