"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
rows = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")) / 1e6
    short = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    agg[short][0] += 1
    agg[short][1] += v
    rows.append((short, v, row["Grid Size"]))
if "--last-step" in sys.argv:
    # one training step = the launches between the last two optimizer launches (k_sgd_ema_step closes a step)
    ends = [i for i, r in enumerate(rows) if "k_sgd_ema" in r[0]]
    if len(ends) >= 2:
        rows = rows[ends[-2] + 1:ends[-1] + 1]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for k, v, _ in rows:
            agg[k][0] += 1
            agg[k][1] += v
        print("# one training step (between the last two k_sgd_ema launches)")
    sys.argv.remove("--last-step")
tot = sum(v for _, v in agg.values())
print(f"# total {tot:.2f} ms over {len(rows)} launches")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.3f} ms {100 * v / tot:5.1f}% n={n:5d}  {k[:100]}")
if len(sys.argv) > 2:
    pat = sys.argv[2]
    print("# individual launches matching", pat)
    for k, v, g in rows:
        if pat in k:
            print(f"   {v:8.4f} ms grid={g} {k[:80]}")
