#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd.py tests/test_gpu_customops.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -6
timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/train_v5.json 2> gpurun_out/train_v5.err; echo rc=$?; tail -2 gpurun_out/train_v5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/train_v5.json")); print(round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms", {k: round(v, 1) for k, v in d["breakdown_ms_per_step"].items() if v > 3}, d["roofline"]["achieved"])
PY
