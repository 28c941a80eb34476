#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -x -k "${K:-conv or model or block}" 2>&1 | tail -6
echo "== bench"
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "rc=$?"; tail -2 gpurun_out/bench_infer.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_infer.json"))
print(round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), d["roofline"]["achieved"], {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
PY
