#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "spike_conv" 2>&1 | tail -8
echo "== conv bench auto"
timeout -k 10 300 python tools/conv_bench.py --ts auto 2>&1 | tee gpurun_out/conv_bench_auto.txt | grep -v '^{"mode'
echo "== bench"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "rc=$?"; tail -3 gpurun_out/bench_infer.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_infer.json"))
print(round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), d["roofline"]["achieved"], {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
PY
