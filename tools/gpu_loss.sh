#!/bin/bash
# loss kernel parity + smoke + training bench with the real loss (and the quadratic one for comparison)
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_post.py -m gpu -q --no-header -p no:cacheprovider -k "loss" 2>&1 | tail -25
echo "== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
echo "== train bench (yolo loss)"
timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/train_yolo_loss.json 2> gpurun_out/train_yolo_loss.err
echo "rc=$?"; tail -3 gpurun_out/train_yolo_loss.err
timeout -k 10 400 python bench.py --mode train --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --loss quadratic > gpurun_out/train_quad_loss.json 2> gpurun_out/train_quad_loss.err
echo "rc=$?"; tail -3 gpurun_out/train_quad_loss.err
python - <<'PY'
import json
for f in ("train_yolo_loss", "train_quad_loss"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "loss", d["last_loss"], {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
