#!/bin/bash
# ncu launch list of ONE eager step of the bench command: 3 warm-up steps are skipped at full speed (--launch-skip), the
# window of 2300 launches that follows covers more than one whole optimize_parameters() period (~1900 launches incl. ATen).
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5900 -c 1950 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-anchor --no-graphs > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tools/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final.txt 2>&1; head -30 gpurun_out/launches_final.txt
echo "== bench"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_c16.json 2> gpurun_out/bench_c16.err; echo "rc=$?"; tail -2 gpurun_out/bench_c16.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_c16.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  clocks %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b["clocks"]))
print("roofline", round(b["roofline"]["frac"], 3), b["roofline"]["ms_per_launch"], [ (round(x["frac"],3), round(x["ms_per_call"],4)) for x in b["roofline_hbm"]])
PY
