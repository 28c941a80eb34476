#!/bin/bash
# 2-GPU sanity of the current code: DDP training with the fused optimizer + stored LIF state, and inference.
mkdir -p gpurun_out
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 "$@" --no-cpu-baseline; }
timeout 600 bash -c "$(declare -f runN); runN --mode train --batch 32 --steps 3 --warmup 3 --no-e2e" > gpurun_out/scale2_train.json 2> gpurun_out/scale2_train.err; echo "train N=2 rc=$?"; tail -3 gpurun_out/scale2_train.err
timeout 600 bash -c "$(declare -f runN); runN --steps 3 --warmup 3" > gpurun_out/scale2_infer.json 2> gpurun_out/scale2_infer.err; echo "infer N=2 rc=$?"
python - <<PY
import json
for k in ("train", "infer"):
    try:
        d = json.load(open(f"gpurun_out/scale2_{k}.json")); print(k, round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms", d["n_gpus"], "gpus")
    except Exception as e:
        print(k, "ERR", e)
PY
