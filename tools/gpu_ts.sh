#!/bin/bash
# TS-mode spike conv + L2-chunked LIF: parity tests, micro-benchmarks, then the model bench.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "spike_conv or lif" 2>&1 | tail -15
echo "== conv bench TS"
timeout -k 10 300 python tools/conv_bench.py --ts 1 2>&1 | tee gpurun_out/conv_bench_ts1.txt | grep -v '^{"mode'
echo "== conv bench SS"
timeout -k 10 300 python tools/conv_bench.py --ts 0 2>&1 | tee gpurun_out/conv_bench_ts0.txt | grep -v '^{"mode'
for mb in 0 16 32 48 80; do
  echo "== ecs bench chunk ${mb} MB"
  ECSY_LIF_CHUNK_MB=$mb timeout -k 10 300 python tools/ecs_bench.py 2>&1 | tail -6
done
echo "== bench"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ts.json 2> gpurun_out/bench_ts.err
echo "rc=$?"; cat gpurun_out/bench_ts.json; tail -3 gpurun_out/bench_ts.err
