#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "lif" 2>&1 | tail -30
echo "== ecs bench fused"
timeout -k 10 300 python tools/ecs_bench.py 2>&1 | tail -6
echo "== ecs bench unfused"
ECSY_LIF_FUSED=0 timeout -k 10 300 python tools/ecs_bench.py 2>&1 | tail -6
