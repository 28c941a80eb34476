#!/bin/bash
# Round-end verification: full GPU parity suite, smoke(), the default bench (with cpu_baseline), the reference arm,
# and the training bench (real loss).
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_final.txt
echo "== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/smoke_final.txt
echo "== bench default"
timeout -k 10 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "rc=$?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "rc=$?"; cat gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
