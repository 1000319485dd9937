#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 900 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/gpu_tests.log 2>&1; tail -9 gpurun_out/gpu_tests.log
grep -h "config 1 losses\|rel-L2: fake_B" gpurun_out/gpu_tests.log
for wpp in 1 2 4; do
  echo "== infer N=1 windows_per_pass=$wpp"
  timeout 300 python3 bench.py --workload infer --steps 3 --warmup 1 --windows-per-pass $wpp > gpurun_out/infer_wpp$wpp.json 2> gpurun_out/infer_wpp$wpp.err; echo "rc=$?"
  python - $wpp <<'PY'
import json, sys
b = json.loads(open("gpurun_out/infer_wpp%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/volume %.2f  Mvox/s %.2f  e2e %.2f launches %d tflops %.0f" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"]/1e6, b["gpu_launches"], b["model_tflops"]))
PY
done
