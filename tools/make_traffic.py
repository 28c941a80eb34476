"""gpurun_out/kernel_metrics.csv (tools/ncu_metrics.sh) -> profiles/traffic.json + a per-kernel table.
The capture holds two forwards (momentum-1 calibration in train mode, then the timed eval forward); the second half of
every kernel's launches is the timed step."""
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
for name, ls in by.items():
    half = ls[len(ls) // 2:]            # timed forward
    ms = sum(x["gpu__time_duration.sum"] for x in half)
    rd = sum(x["dram__bytes_read.sum"] for x in half)
    wr = sum(x["dram__bytes_write.sum"] for x in half)
    tp = sum(x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * x["gpu__time_duration.sum"] for x in half) / max(ms, 1e-9)
    rows.append((ms, name, len(half), rd, wr, tp))
    if name.startswith("k_spike_conv_ts") or re.match(r"k_umma_gemm<\d+, 1,", name):
        conv["ms"] += ms; conv["rd"] += rd; conv["wr"] += wr; conv["n"] += len(half); conv["tw"] += tp * ms
print("# per kernel over the timed eval forward (resnet34, T=4, batch 64, fast): ncu serialised, cold-cache durations")
print(f"# {'kernel':58s} {'n':>4s} {'ms':>8s} {'DRAM GB':>8s} {'GB/s':>7s} {'tensor %':>8s}")
for ms, name, n, rd, wr, tp in sorted(rows, reverse=True):
    print(f"  {name[:58]:58s} {n:4d} {ms:8.3f} {(rd + wr) / 1e9:8.2f} {(rd + wr) / 1e9 / (ms * 1e-3):7.0f} {tp:8.1f}")
out = {"resnet34|64|4|fast": {
    "kernel": "k_spike_conv_ts<*> + k_umma_gemm<*,spikes,*> (the spike-conv launches of the timed eval forward)",
    "dram_bytes_per_step": conv["rd"] + conv["wr"], "dram_read_bytes_per_step": conv["rd"],
    "dram_write_bytes_per_step": conv["wr"], "ncu_time_ms_per_step": conv["ms"],
    "tensor_pipe_pct_time_weighted": conv["tw"] / max(conv["ms"], 1e-9), "launches": conv["n"],
    "source": "profiles/r01_kernel_metrics_b64.csv (tools/ncu_metrics.sh: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
              "dram__bytes_write.sum,sm__pipe_tensor_cycles_active... on `python bench.py --batch 64 --steps 1 --warmup 0 "
              "--min-warmup 0 --no-e2e --no-cpu-baseline`)"}}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print("# conv totals:", conv)
