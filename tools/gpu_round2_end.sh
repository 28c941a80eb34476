#!/bin/bash
# End-of-round-2 evidence at HEAD: GPU tests, the driver's default bench line, the reference arm, the ncu launch lists of the
# timed inference step and of one training step.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
t0=$(date +%s)
python -m pytest tests -m gpu -q 2>&1 | grep -v "mbarrier timeout" | tail -4 > gpurun_out/r02_pytest_gpu.txt
cat gpurun_out/r02_pytest_gpu.txt; echo "tests $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/r02_bench_default.err
python -c "import json; d=json.load(open('gpurun_out/r02_bench_default.json')); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), 'train', round(d['train']['images_per_s'],1), round(d['train']['ms_per_step'],1), 'parity', round(d['parity']['parity_precision_images_per_s'],1), d['clocks'])"
t0=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
echo "reference arm $(( $(date +%s) - t0 )) s"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
if [ -z "$NO_NCU" ]; then
PROF="python bench.py --steps 2 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline --no-train --no-parity-leg --no-small-batch"
timeout -k 10 300 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench_b64.csv $PROF > gpurun_out/ncu_launches.log 2>&1
echo "ncu infer launches rc=$?"
if [ -z "$NO_NCU_TRAIN" ]; then
PROF="python bench.py --mode train --model resnet18 --batch 32 --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline"
timeout -k 10 300 $PROF > gpurun_out/prof_train_plain.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_train_r18_b32.csv $PROF > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu train launches rc=$?"
fi
fi
