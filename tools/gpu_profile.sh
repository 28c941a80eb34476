#!/bin/bash
# Round-end evidence run: tests, bench (infer with CPU baseline + reference arm, train), conv micro-benchmark, ncu launch
# list of the bench command, one full-set capture each of the two spike-conv kernels.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4
echo "== bench infer (default flags)"
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench.json
echo "== bench reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
echo "== bench train"
timeout -k 10 900 python bench.py --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_train.json; tail -2 gpurun_out/bench_train.err
echo "== conv microbench"
timeout -k 10 300 python tools/conv_bench.py --mode fast --ts auto > gpurun_out/conv_bench_auto.txt 2>&1; tail -1 gpurun_out/conv_bench_auto.txt | cut -c1-900
PROF="python bench.py --batch 16 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
CB="python tools/conv_bench.py --mode fast --only 1 --reps 2"
timeout -k 10 300 $CB > gpurun_out/cb_plain.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_spike_conv_ts -s 3 -c 1 -o gpurun_out/prof_ts128 -f $CB > gpurun_out/ncu_full_ts128.log 2>&1
echo "ncu full ts128 rc=$?"
CB="python tools/conv_bench.py --mode fast --only 3 --reps 2"
timeout -k 10 300 $CB > gpurun_out/cb_plain2.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -s 3 -c 1 -o gpurun_out/prof_ss512 -f $CB > gpurun_out/ncu_full_ss512.log 2>&1
echo "ncu full ss512 rc=$?"
