"""Per-launch table from an `ncu --csv --metrics ...` log: one row per launch, one column per metric."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows, order = {}, []
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in rows:
        rows[i] = {"k": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""), "grid": r["Grid Size"]}
        order.append(i)
    rows[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
last = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sel = [rows[i] for i in order if pat in rows[i]["k"] and not rows[i]["k"].startswith("at::")][-last:]
for r in sel:
    t = r.get("gpu__time_duration.sum", 0) / 1e3
    rd, wr = r.get("dram__bytes_read.sum", 0) / 1e6, r.get("dram__bytes_write.sum", 0) / 1e6
    l2 = r.get("lts__t_bytes.sum", 0) / 1e6
    occ = r.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
    print(f"{r['k'][:34]:34s} {r['grid']:14s} {t:8.1f} us  rd {rd:7.1f} wr {wr:7.1f} MB  L2 {l2:8.1f} MB  {(rd + wr) / t / 1e3 if t else 0:6.2f} TB/s  occ {occ:4.1f}%")
