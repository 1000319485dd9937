#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 600 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== conv layers (phase kernel everywhere it is eligible)"; MRA_GATHER_PHASE_ALL=1 timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cat gpurun_out/conv_bench.txt; cp gpurun_out/conv_bench.txt gpurun_out/conv_bench_phase.txt
echo "== conv layers (MRA_GATHER_NOPHASE=1)"; MRA_GATHER_NOPHASE=1 timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; grep -E "G.c1|G.c4|G.d1|G.d2|G.u1|G.u2|D.2|D.3|total" gpurun_out/conv_bench.txt; cp gpurun_out/conv_bench.txt gpurun_out/conv_bench_nophase.txt
echo "== counters G.u2 fprop, G.c1"
for L in "128 64 3 2 1 1 1 64" "1 64 7 1 0 0 0 134"; do
  echo "-- layer $L"; MRA_GATHER_PHASE_ALL=1 MRA_GATHER_DEBUG=2 ONLY=fprop timeout 100 python tools/layer_bench.py $L 2 2>&1 | tail -5
done
echo "== bench"; timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "rc=$?"
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_v3.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
PY
