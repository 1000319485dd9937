#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== nstats bench"; timeout 200 python tools/nstats_bench.py 2 > gpurun_out/nstats_bench.txt 2>&1; cat gpurun_out/nstats_bench.txt
for mode in 1 2 0; do
  echo "== bench MRA_NORM_BWD_FUSED=$mode"
  MRA_NORM_BWD_FUSED=$mode timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_nb$mode.json 2> gpurun_out/bench_nb$mode.err; echo "rc=$?"
  python - $mode <<'PY'
import json, sys
b = json.loads(open("gpurun_out/bench_nb%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["gpu_launches"], b["clocks"]))
PY
done
