#!/bin/bash
# Last call of the round (2 GPU-minutes left): the pointer-walk staging loop of shift_sum -- the ops that use it + per-layer table.
set -u
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_ops_gpu.py tests/test_zz_fullsize.py -q -m gpu -x -k "c1-64 or c64-1 or c8-1 or c128-1 or c512-1 or c1-8 or G.c1 or G.c4 or D.5 or head or stem" > gpurun_out/gpu_tests_ss.log 2>&1; tail -2 gpurun_out/gpu_tests_ss.log
timeout 40 python tools/conv_bench.py 2 > /dev/null 2>&1; grep -E "layer|G.c1|G.c4|D.5|total" gpurun_out/conv_bench.txt | cut -c1-100
