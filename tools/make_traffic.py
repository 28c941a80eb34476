"""gpurun_out/kernel_metrics.csv (tools/ncu_metrics.sh) -> profiles/traffic.json + a per-kernel table.
The capture holds three forwards (momentum-1 calibration in train mode, the timed eval forward, the per-operator profile
forward of bench.py); the last third of every kernel's launches is one eval step."""
import collections, csv, json, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kernel_metrics.csv"
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.OrderedDict()   # launch id -> dict
for row in csv.DictReader(lines):
    k = row["ID"]
    d = per.setdefault(k, {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    m = row["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6}.get(unit, 1)   # -> ms
    d[m] = v
by = collections.OrderedDict()
for d in per.values():
    by.setdefault(d["name"], []).append(d)
rows = []
conv = {"ms": 0.0, "rd": 0.0, "wr": 0.0, "n": 0, "tw": 0.0}
lif = {"ms": 0.0, "rd": 0.0, "wr": 0.0, "n": 0, "tw": 0.0}
LIF_KERNELS = ("k_lif_ecs_wave64", "k_lif_first", "k_spread_dw", "k_dense_tma_h", "k_ecs_step")
for name, ls in by.items():
    half = ls[len(ls) - len(ls) // 3:]            # one eval forward
    ms = sum(x["gpu__time_duration.sum"] for x in half)
    rd = sum(x["dram__bytes_read.sum"] for x in half)
    wr = sum(x["dram__bytes_write.sum"] for x in half)
    tp = sum(x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * x["gpu__time_duration.sum"] for x in half) / max(ms, 1e-9)
    rows.append((ms, name, len(half), rd, wr, tp))
    if name.startswith("k_spike_conv_ts") or re.match(r"k_umma_gemm<\d+, 1,", name):
        conv["ms"] += ms; conv["rd"] += rd; conv["wr"] += wr; conv["n"] += len(half); conv["tw"] += tp * ms
    if name.startswith(LIF_KERNELS):
        lif["ms"] += ms; lif["rd"] += rd; lif["wr"] += wr; lif["n"] += len(half); lif["tw"] += tp * ms
print("# per kernel over the timed eval forward (resnet34, T=4, batch 64, fast): ncu serialised, cold-cache durations")
print(f"# {'kernel':58s} {'n':>4s} {'ms':>8s} {'DRAM GB':>8s} {'GB/s':>7s} {'tensor %':>8s}")
for ms, name, n, rd, wr, tp in sorted(rows, reverse=True):
    print(f"  {name[:58]:58s} {n:4d} {ms:8.3f} {(rd + wr) / 1e9:8.2f} {(rd + wr) / 1e9 / (ms * 1e-3):7.0f} {tp:8.1f}")
SRC = ("profiles/r02_kernel_metrics_b64.csv (tools/ncu_metrics.sh: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
       "dram__bytes_write.sum,sm__pipe_tensor_cycles_active... on `python bench.py --batch 64 --steps 1 --warmup 0 "
       "--min-warmup 0 --no-e2e --no-cpu-baseline --no-train --no-parity-leg`)")
def group(d, what):
    return {"kernel": what, "dram_bytes_per_step": d["rd"] + d["wr"], "dram_read_bytes_per_step": d["rd"],
            "dram_write_bytes_per_step": d["wr"], "ncu_time_ms_per_step": d["ms"],
            "tensor_pipe_pct_time_weighted": d["tw"] / max(d["ms"], 1e-9), "launches": d["n"], "source": SRC}
out = {"resnet34|64|4|fast": {
    "conv": group(conv, "k_spike_conv_ts<*> + k_umma_gemm<*,spikes,*> (the spike-conv launches of the timed eval forward)"),
    "lif": group(lif, "k_lif_ecs_wave64 + k_lif_first + k_spread_dw + k_dense_tma_h + k_ecs_step (the ECS-LIF launches of the "
                      "timed eval forward)")}}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print("# conv totals:", conv)
print("# lif totals:", lif)
