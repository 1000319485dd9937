#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 900 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== bench unet b4"; timeout 300 python3 bench.py --workload unet --batch 4 --steps 10 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f Mvox/s %.2f launches %d' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches']))"
echo "== adam kernel time (unet)"; timeout 200 python tools/profile_graph_step.py 4 unet_128 2>&1 | grep -E "graph-replayed|adam_kernel|wgrad_tc|gather_tc" | cut -c1-150
