#!/bin/bash
# Same-box GPU comparison: the stock-PyTorch eager port of the reference forward on the B200.
mkdir -p gpurun_out
timeout -k 5 100 python bench.py --impl reference --ref-device cuda --cpu-sample 16 --steps 3 --warmup 2 > gpurun_out/eager_gpu_fp32.json 2> gpurun_out/eager_gpu_fp32.err
echo "rc=$?"; cut -c1-400 gpurun_out/eager_gpu_fp32.json; tail -2 gpurun_out/eager_gpu_fp32.err
timeout -k 5 60 python bench.py --impl reference --ref-device cuda --ref-autocast --cpu-sample 16 --steps 3 --warmup 2 > gpurun_out/eager_gpu_amp.json 2> gpurun_out/eager_gpu_amp.err
echo "rc=$?"; cut -c1-400 gpurun_out/eager_gpu_amp.json; tail -2 gpurun_out/eager_gpu_amp.err
