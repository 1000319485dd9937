#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== eager step profile"; timeout 300 python tools/profile_step.py 2 > gpurun_out/profile_step.log 2>&1; tail -3 gpurun_out/profile_step.log
echo "== graph step profile b2"; timeout 300 python tools/profile_graph_step.py 2 2>&1 | tail -45
echo "== bench"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_c13.json 2> gpurun_out/bench_c13.err; echo "rc=$?"; tail -2 gpurun_out/bench_c13.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_c13.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
print("roofline", round(b["roofline"]["frac"], 3), b["roofline"]["ms_per_launch"], [ (round(x["frac"],3), round(x["ms_per_call"],4)) for x in b["roofline_hbm"]])
PY
