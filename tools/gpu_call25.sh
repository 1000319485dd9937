#!/bin/bash
# Dual-item mode of gather_halo_kernel (MRA_HALO_DUAL=1): GPU suite with the default path and with dual items, per-layer
# table and step A/B on one box.  Also the first hardware run of ConvFn without materialised statistics gradients.
set -u
mkdir -p gpurun_out
t0=$SECONDS
echo "== gpu tests (default)";  timeout 300 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
echo "   t=$((SECONDS-t0))s"
echo "== gpu tests (MRA_HALO_DUAL=1)";  MRA_HALO_DUAL=1 timeout 300 python -m pytest tests -q -m gpu --durations=3 > gpurun_out/gpu_tests_dual.log 2>&1; tail -12 gpurun_out/gpu_tests_dual.log | cut -c1-200
echo "   t=$((SECONDS-t0))s"
echo "== conv layers (default)"; timeout 120 python tools/conv_bench.py 2 > /dev/null 2>&1; grep -E "G.rb|D.4|total" gpurun_out/conv_bench.txt | cut -c1-100; cp gpurun_out/conv_bench.txt gpurun_out/conv_bench_default.txt
echo "== conv layers (dual)"; MRA_HALO_DUAL=1 timeout 120 python tools/conv_bench.py 2 > /dev/null 2>&1; grep -E "G.rb|D.4|total" gpurun_out/conv_bench.txt | cut -c1-100; cp gpurun_out/conv_bench.txt gpurun_out/conv_bench_dual.txt
echo "   t=$((SECONDS-t0))s"
line() { python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms/step %.2f Mvox/s %.2f launches %d sm_mhz %s rb_fprop_ms %.4f frac %.3f' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches'], b['clocks']['sm_mhz'], b['roofline']['ms_per_launch'], b['roofline']['frac']))"; }
echo "== bench A (default)";  timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | tee gpurun_out/bench_A1.json | line A1
echo "== bench B (MRA_HALO_DUAL=1)"; MRA_HALO_DUAL=1 timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | tee gpurun_out/bench_B1.json | line B1
echo "   t=$((SECONDS-t0))s"
