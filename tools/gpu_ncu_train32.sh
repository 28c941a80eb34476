#!/bin/bash
mkdir -p gpurun_out
PROF="python bench.py --mode train --batch 32 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 300 $PROF > gpurun_out/prof_train_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_train_b32.csv $PROF > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_train_b32.csv | head -45
