#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 600 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
for pdl in 1 0 1 0; do
  echo "== bench MRA_PDL=$pdl"
  MRA_PDL=$pdl timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "rc=$?"; tail -2 gpurun_out/bench_pdl$pdl.err
  python - $pdl <<'PY'
import json, sys
b = json.loads(open("gpurun_out/bench_pdl%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"]/1e6, b["gpu_launches"], b["clocks"]))
PY
done
echo "== eager step (no graphs)"; for pdl in 1 0; do MRA_PDL=$pdl timeout 300 python3 bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline --no-anchor --no-graphs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$pdl eager ms/step %.2f' % b['ms_per_step'])"; done
