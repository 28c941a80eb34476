#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_post.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15
