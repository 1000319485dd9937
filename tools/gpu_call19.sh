#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 900 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
echo "== unet layers b4 (split-K)"; timeout 200 python tools/unet_layers.py 4 2>&1 | tail -17
echo "== unet layers b4 (MRA_GATHER_NOKSPLIT=1)"; MRA_GATHER_NOKSPLIT=1 timeout 200 python tools/unet_layers.py 4 2>&1 | tail -2
for v in 0 1; do
  echo "== bench unet b4, NOKSPLIT=$v"
  if [ $v = 1 ]; then export MRA_GATHER_NOKSPLIT=1; else unset MRA_GATHER_NOKSPLIT; fi
  timeout 300 python3 bench.py --workload unet --batch 4 --steps 10 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f Mvox/s %.2f launches %d' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches']))"
done
unset MRA_GATHER_NOKSPLIT
echo "== bench train (regression check)"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f Mvox/s %.2f launches %d' % (b['ms_per_step'], b['value']/1e6, b['gpu_launches']))"
