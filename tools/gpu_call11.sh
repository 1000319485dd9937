#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 600 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -6 gpurun_out/gpu_tests.log
echo "== stem/head layers, scatter epilogue vs shift_sum"
for L in "64 1 7 1 0 0 0 134" "1 64 7 1 0 0 0 134" "512 1 4 1 1 0 0 15"; do
  echo "-- layer $L"
  echo -n "scatter   "; KERNELS=1 timeout 100 python tools/layer_bench.py $L 2 2>&1 | grep -E "^fprop|^dgrad|gather_col|shift_sum|scatter_finish|Memset" | tr '\n' ';'; echo
  echo -n "shift_sum "; MRA_COL_NOSCATTER=1 timeout 100 python tools/layer_bench.py $L 2 2>&1 | grep -E "^fprop|^dgrad" | tr '\n' ';'; echo
done
echo "== bench"; timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "rc=$?"; tail -2 gpurun_out/bench_v7.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_v7.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
PY
echo "== bench (MRA_COL_NOSCATTER=1)"; MRA_COL_NOSCATTER=1 timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_v7b.json 2> gpurun_out/bench_v7b.err; echo "rc=$?"
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_v7b.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  launches %d" % (b["ms_per_step"], b["value"] / 1e6, b["gpu_launches"]))
PY
