#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -6
for st in 1 0; do
ECSY_LIF_STORE=$st timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/train_store$st.json 2> gpurun_out/train_store$st.err; echo rc=$?; tail -2 gpurun_out/train_store$st.err
done
python - <<'PY'
import json
for f in ("train_store1", "train_store0"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms", {k: round(v, 1) for k, v in d["breakdown_ms_per_step"].items() if v > 5})
    except Exception as e:
        print(f, "failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
