#!/bin/bash
mkdir -p gpurun_out
PROF="python bench.py --batch 64 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_spread_dw|k_ecs_step|k_lif_first|k_umma_gemm" -s 526 -c 560 --csv --log-file gpurun_out/lif_metrics.csv $PROF > gpurun_out/ncu_lif.log 2>&1
echo "rc=$?"
