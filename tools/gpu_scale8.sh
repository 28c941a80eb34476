#!/bin/bash
# 1 -> N GPU weak scaling on ONE box (inference and DDP training), N = $1 (default 8).
N=${1:-8}
mkdir -p gpurun_out
run1() { python bench.py "$@" --no-cpu-baseline; }
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" --no-cpu-baseline; }
timeout 600 bash -c "$(declare -f run1); run1 --steps 5 --warmup 3" > gpurun_out/scale_infer_1.json 2> gpurun_out/scale_infer_1.err; echo "infer N=1 rc=$?"
timeout 900 bash -c "N=$N; $(declare -f runN); runN --steps 5 --warmup 3" > gpurun_out/scale_infer_$N.json 2> gpurun_out/scale_infer_$N.err; echo "infer N=$N rc=$?"
timeout 600 bash -c "$(declare -f run1); run1 --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2 --no-e2e" > gpurun_out/scale_train_1.json 2> gpurun_out/scale_train_1.err; echo "train N=1 rc=$?"
timeout 900 bash -c "N=$N; $(declare -f runN); runN --mode train --batch 32 --steps 3 --warmup 2 --min-warmup 2 --no-e2e" > gpurun_out/scale_train_$N.json 2> gpurun_out/scale_train_$N.err; echo "train N=$N rc=$?"
python - <<PY
import json
for k in ("infer", "train"):
    v = {}
    for n in (1, $N):
        try:
            d = json.load(open(f"gpurun_out/scale_{k}_{n}.json")); v[n] = (d["value"], d["ms_per_step"])
        except Exception as e:
            v[n] = ("ERR", str(e)[:80])
    print(k, v, "speedup", (v[$N][0] / v[1][0]) if isinstance(v[$N][0], float) and isinstance(v[1][0], float) else None)
PY
tail -3 gpurun_out/scale_train_$N.err
