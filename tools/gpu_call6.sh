#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== ncu stem fprop (col kernel)"
ONLY=fprop timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gather_col" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_col_stem -f python tools/layer_bench.py 1 64 7 1 0 0 0 134 2 > gpurun_out/ncu_col.log 2>&1; tail -2 gpurun_out/ncu_col.log
echo "== ncu G.u2 fprop (merged per-tap kernel)"
ONLY=fprop timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gather_tc" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tc_u2 -f python tools/layer_bench.py 128 64 3 2 1 1 1 64 2 > gpurun_out/ncu_u2.log 2>&1; tail -2 gpurun_out/ncu_u2.log
echo "== bench"; timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "rc=$?"
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_v4.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
PY
