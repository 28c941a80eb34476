#!/bin/bash
# Session 3, call 1: full GPU parity suite (incl. res*-ee blocks, spread_dw versions), dw microbench, bench v1 vs v2,
# one ncu full capture of each depth-wise kernel.
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
echo "== dw microbench"
timeout -k 10 200 python tools/dw_bench.py 2>&1 | tee gpurun_out/dw_bench.txt | tail -8
echo "== bench (dw v2 default)"
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dw2.json 2> gpurun_out/bench_dw2.err
echo "rc=$?"; tail -2 gpurun_out/bench_dw2.err
echo "== bench (dw v1)"
ECSY_DW_V=1 timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dw1.json 2> gpurun_out/bench_dw1.err
echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_dw2", "bench_dw1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d.get("breakdown_ms_per_step", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
echo "== ncu dw kernels"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_spread_dw -c 2 -o gpurun_out/ncu_dw -f python tools/dw_bench.py --reps 1 > gpurun_out/ncu_dw.log 2>&1
echo "ncu rc=$?"
echo "== ee model bench"
timeout -k 10 300 python bench.py --model res18-ee --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_res18ee.json 2> gpurun_out/bench_res18ee.err
echo "rc=$?"; tail -3 gpurun_out/bench_res18ee.err; cat gpurun_out/bench_res18ee.json | cut -c1-400
