#!/bin/bash
# Other BASELINE configs: resnet18 (Stack B) inference/training, resnet10, T sweep on resnet18.
mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 600 python bench.py "$@" --no-cpu-baseline --no-train --no-parity-leg --no-small-batch > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/cfg_$name.json')); print(round(d['value'],1),'img/s',round(d['ms_per_step'],1),'ms')" 2>/dev/null)"; tail -1 gpurun_out/cfg_$name.err | cut -c1-200; }
run r18_infer --model resnet18 --steps 3 --warmup 3
run r18_train --model resnet18 --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2
run r10_infer --model resnet10 --steps 3 --warmup 3
run r18_T1 --model resnet18 --T 1 --steps 3 --warmup 3
run r18_T8 --model resnet18 --T 8 --batch 32 --steps 3 --warmup 3
run r34_parity --model resnet34 --precision parity --steps 3 --warmup 3
run r34_gen1_T5 --model resnet34 --T 5 --events --batch 32 --steps 3 --warmup 3
run r34_train --model resnet34 --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2
run res18ee_infer --model res18-ee --batch 32 --steps 3 --warmup 3
# BASELINE configs[3] also names TRAINING on Gen1 event frames (T = 5): only run with GEN1_TRAIN=1 or as the whole list
run r34_gen1_T5_train --model resnet34 --T 5 --events --mode train --batch 16 --steps 3 --warmup 2 --min-warmup 2
run res18ee_gen1_T5_train --model res18-ee --T 5 --events --mode train --batch 16 --steps 3 --warmup 2 --min-warmup 2
