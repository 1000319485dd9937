#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== batch invariance"; timeout 400 python tools/batch_invariance.py 128 2>&1 | grep -E "generator|G.rb|tc error"
echo "== gpu tests";  timeout 900 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
echo "== bench"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_c17.json 2> gpurun_out/bench_c17.err; echo "rc=$?"; tail -2 gpurun_out/bench_c17.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_c17.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
print("roofline", round(b["roofline"]["frac"], 3), b["roofline"]["ms_per_launch"], [ (round(x["frac"],3), round(x["ms_per_call"],4)) for x in b["roofline_hbm"]])
PY
echo "== G.rb layer"; timeout 100 python tools/layer_bench.py 256 256 3 1 0 0 0 34 2 2>&1 | grep -E "^fprop|^dgrad|^wgrad"
