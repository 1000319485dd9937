#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 900 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== unet layers b4"; timeout 200 python tools/unet_layers.py 4 2>&1 | tail -16 | cut -c1-112
echo "== conv layers"; timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cut -c1-100 gpurun_out/conv_bench.txt | tail -14
echo "== bench unet b4"; timeout 300 python3 bench.py --workload unet --batch 4 --steps 10 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f Mvox/s %.2f launches %d' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches']))"
echo "== bench train"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f Mvox/s %.2f launches %d clocks %s' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches'], b['clocks']))"
