#!/bin/bash
# 45 GPU-seconds left in the round: the rebuilt library (host-side NULL / dtype checks in the conv entry points) through as
# much of the GPU suite as fits -- the op tests first (every conv entry point through the C ABI), then models, then full size.
set -u
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 36 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_zz_fullsize.py tests/test_options_gpu.py -m gpu -x -v -p no:cacheprovider > gpurun_out/gpu_tests_nullcheck.log 2>&1
echo "rc=$?"
grep -c PASSED gpurun_out/gpu_tests_nullcheck.log; grep -E "FAILED|ERROR" gpurun_out/gpu_tests_nullcheck.log | head -5; tail -2 gpurun_out/gpu_tests_nullcheck.log | cut -c1-200
