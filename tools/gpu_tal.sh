#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_post.py -m gpu -q --no-header -p no:cacheprovider -k "tal or loss" 2>&1 | tail -30
echo "== train bench resnet18 (TAL loss)"
timeout -k 10 400 python bench.py --mode train --model resnet18 --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/train_r18_tal.json 2> gpurun_out/train_r18_tal.err
echo "rc=$?"; tail -3 gpurun_out/train_r18_tal.err
timeout -k 10 400 python bench.py --mode train --model resnet18 --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --loss quadratic > gpurun_out/train_r18_quad.json 2> gpurun_out/train_r18_quad.err
echo "rc=$?"; tail -3 gpurun_out/train_r18_quad.err
python - <<'PY'
import json
for f in ("train_r18_tal", "train_r18_quad"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "loss", d["last_loss"], {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
