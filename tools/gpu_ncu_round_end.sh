#!/bin/bash
# Round-end ncu launch lists (per-launch device times, cold cache / serialised: the SHARES are what count):
# inference bench at batch 16, training bench at batch 8 (real loss).
# NOTE: the batch-32 training step under ncu (12 000 launches, 35 GB of stored LIF state) did not finish within 9 minutes
# of GPU time in round 1 -- keep the training capture at batch 8 and ECSY_LIF_STORE=0.
mkdir -p gpurun_out
PROF="python bench.py --batch 16 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 300 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_infer_b16_end.csv $PROF > gpurun_out/ncu_launches.log 2>&1
echo "ncu infer rc=$?"
python tools/summarize_launches.py gpurun_out/launches_infer_b16_end.csv | head -25
PROF="python bench.py --mode train --batch 8 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 300 $PROF > gpurun_out/prof_train_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_train_b8_end.csv $PROF > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu train rc=$?"
python tools/summarize_launches.py gpurun_out/launches_train_b8_end.csv | head -45
