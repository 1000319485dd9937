#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 600 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== conv layers"; timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cat gpurun_out/conv_bench.txt
echo "== counters G.c1 fprop"
MRA_GATHER_DEBUG=2 ONLY=fprop timeout 100 python tools/layer_bench.py 1 64 7 1 0 0 0 134 2 2>&1 | tail -5
echo "== bench"; timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "rc=$?"; tail -3 gpurun_out/bench_v5.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_v5.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
print("roofline", b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["ms_per_launch"])
PY
echo "== step table"; timeout 120 python tools/profile_step.py > /dev/null 2>&1; head -34 gpurun_out/step_profile.txt | cut -c1-100,190-260
