#!/bin/bash
# DRAM bytes / duration / tensor-pipe activity of every hot-path kernel over one calibration + one timed eval forward
# (batch 64).  tools/make_traffic.py turns the CSV into profiles/traffic.json and the per-kernel HBM / tensor table.
mkdir -p gpurun_out
PROF="python bench.py --batch 64 --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline --no-train --no-parity-leg --no-small-batch"
timeout -k 10 600 $PROF > gpurun_out/prof_plain.log 2>&1 &&
timeout -k 10 2400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:"k_spike_conv_ts|k_umma_gemm|k_dense_tma|k_ecs_step|k_spread_dw|k_lif_first|k_lif_ecs_wave64|k_im2col|k_resample|k_transpose|k_stem_umma|k_to_nhwc4|k_affine_add" -c 3000 --csv --log-file gpurun_out/kernel_metrics.csv $PROF > gpurun_out/ncu_metrics.log 2>&1
echo "ncu metrics rc=$?"
