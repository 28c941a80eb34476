#!/bin/bash
# Standard GPU pass: parity tests, bench, ncu launch list + one full capture of the top kernel.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15
echo "== bench batch 64"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$NCU" = "1" ]; then
  PROF="python bench.py --batch 16 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
  timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
  timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout -k 10 600 $PROF > gpurun_out/prof_plain2.log 2>&1 &&
  timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_umma_gemm} -s ${NCU_SKIP:-120} -c ${NCU_COUNT:-4} -o gpurun_out/prof -f $PROF > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
