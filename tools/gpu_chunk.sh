#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -x -k "lif or block or model" 2>&1 | tail -4
for mb in 48 0 24 96; do
ECSY_LIF_CHUNK_MB=$mb timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_chunk$mb.json 2> gpurun_out/bench_chunk$mb.err; echo "chunk $mb rc=$?"
done
python - <<'PY'
import json
for mb in (48, 0, 24, 96):
    try:
        d = json.load(open(f"gpurun_out/bench_chunk{mb}.json")); print(mb, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d["breakdown_ms_per_step"].items() if v > 1}, d["gpu_launches"])
    except Exception as e:
        print(mb, "failed", e)
PY
