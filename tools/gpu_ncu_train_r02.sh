#!/bin/bash
# ncu launch list of the training step of the `train` sub-record (resnet18, batch 32, TAL loss, fused optimizer).
mkdir -p gpurun_out
PROF="python bench.py --mode train --model resnet18 --batch 32 --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline"
timeout -k 10 300 $PROF > gpurun_out/prof_train_plain.log 2>&1 &&
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_train_r18_b32.csv $PROF > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu launches rc=$?"
tail -3 gpurun_out/prof_train_plain.log | cut -c1-600
