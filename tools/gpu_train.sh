#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_train.py tests/test_gpu_customops.py tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -6
for cfg in "resnet34 r34_train" "resnet18 r18_train"; do set -- $cfg
  timeout -k 10 600 python bench.py --model $1 --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/cfg_$2.json 2> gpurun_out/cfg_$2.err
  echo "$2 rc=$?"; tail -1 gpurun_out/cfg_$2.err | cut -c1-200
  python - <<PY
import json
d = json.load(open("gpurun_out/cfg_$2.json"))
print(round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 1) for k, v in d.get("breakdown_ms_per_step", {}).items()})
PY
done
