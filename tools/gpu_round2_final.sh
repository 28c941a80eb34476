#!/bin/bash
# Round-2 evidence run: GPU tests, the default bench line (driver's command), the reference arm, the ncu launch list of the
# timed inference step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "mbarrier timeout" | tail -4 > gpurun_out/r02_pytest_gpu.txt
cat gpurun_out/r02_pytest_gpu.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -c 600 gpurun_out/r02_bench_default.json
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
cat gpurun_out/r02_bench_reference_arm.json | cut -c1-400
PROF="python bench.py --steps 2 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline --no-train --no-parity-leg --no-small-batch"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench_b64.csv $PROF > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
