#!/bin/bash
# Full GPU pass: all parity tests, inference bench, training bench.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -8
echo "== bench infer b64"
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "rc=$?"; tail -2 gpurun_out/bench_infer.err
echo "== bench train b32"
timeout -k 10 600 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
echo "rc=$?"; tail -2 gpurun_out/bench_train.err
python - <<'PY'
import json
for f in ("bench_infer", "bench_train"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", d["roofline"]["achieved"], {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
