#!/bin/bash
# shift_sum rewrite + dead-plane skipping + wgrad flush unroll: GPU suite, per-layer table, step A/B against MRA_HALO_NOSKIP.
set -u
mkdir -p gpurun_out
t0=$SECONDS
echo "== gpu tests";  timeout 400 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
echo "   t=$((SECONDS-t0))s"
line() { python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 ms/step %.2f Mvox/s %.2f launches %d sm_mhz %s rb_fprop_ms %.4f' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches'], b['clocks']['sm_mhz'], b['roofline']['ms_per_launch']))"; }
echo "== bench A (default)";  timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | tee gpurun_out/bench_A1.json | line A1
echo "== bench B (MRA_HALO_NOSKIP=1)"; MRA_HALO_NOSKIP=1 timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | tee gpurun_out/bench_B1.json | line B1
echo "== bench A again";  timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | tee gpurun_out/bench_A2.json | line A2
echo "   t=$((SECONDS-t0))s"
echo "== conv layers"; timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cut -c1-100 gpurun_out/conv_bench.txt | tail -15
echo "   t=$((SECONDS-t0))s"
