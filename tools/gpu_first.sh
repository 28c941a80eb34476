#!/bin/bash
# First-contact GPU run: every test file in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_ops.py tests/test_gpu_blocks.py; do
  b=$(basename $f .py)
  timeout -k 10 600 python -m pytest $f -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -40 gpurun_out/$b.log
done
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/smoke.log
