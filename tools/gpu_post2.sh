#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 500 python tests/diag/post_bench.py 2>&1 | tail -12
echo "== train bench fused vs torch (batch 32)"
timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/train_fused.json 2> gpurun_out/train_fused.err; echo rc=$?; tail -2 gpurun_out/train_fused.err
timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --optim torch --no-e2e > gpurun_out/train_torch.json 2> gpurun_out/train_torch.err; echo rc=$?
python - <<'PY'
import json
for f in ("train_fused", "train_torch"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms", d.get("gpu_launches"))
    except Exception as e:
        print(f, "failed", e)
PY
