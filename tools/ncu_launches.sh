#!/bin/bash
# ncu launch list of one short bench run (batch 16): per-launch device times.
mkdir -p gpurun_out
PROF="python bench.py --batch 16 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
