#!/bin/bash
# Round-end evidence run: tests, bench (infer + train), ncu DRAM/tensor metrics of the tcgen05 kernels over one
# timed forward, one full-set capture of a representative spike-conv launch.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4
echo "== bench infer"
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench.json
echo "== bench train"
timeout -k 10 900 python bench.py --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_train.json; tail -2 gpurun_out/bench_train.err
echo "== conv microbench"
timeout -k 10 300 python tools/conv_bench.py --mode fast > gpurun_out/conv_bench_fast.txt 2>&1; tail -1 gpurun_out/conv_bench_fast.txt | cut -c1-600
PROF="python bench.py --batch 64 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:k_umma_gemm -c 400 --csv --log-file gpurun_out/umma_metrics.csv $PROF > gpurun_out/ncu_umma.log 2>&1
echo "ncu umma metrics rc=$?"
CB="python tools/conv_bench.py --mode fast --only 2 --reps 2"
timeout -k 10 300 $CB > gpurun_out/cb_plain.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -s 3 -c 1 -o gpurun_out/prof_conv256 -f $CB > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
