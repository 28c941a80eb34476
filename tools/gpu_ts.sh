#!/bin/bash
# TS-mode spike conv: parity tests, micro-benchmarks, then the model bench.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "spike_conv" 2>&1 | tail -15
echo "== conv bench TS"
timeout -k 10 300 python tools/conv_bench.py --ts all 2>&1 | tee gpurun_out/conv_bench_ts1.txt | grep -v '^{"mode'
echo "== bench"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ts.json 2> gpurun_out/bench_ts.err
echo "rc=$?"; cat gpurun_out/bench_ts.json; tail -3 gpurun_out/bench_ts.err
