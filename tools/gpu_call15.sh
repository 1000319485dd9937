#!/bin/bash
# Final-build evidence on one B200: driver command (anchor + cpu baseline), reference arm sample, smoke, per-layer and
# norm tables, ncu --set full of the G.rb / norm kernels at batch 2 and 4, ncu launch list of one eager step.
set -u
mkdir -p gpurun_out
echo "== smoke";  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (driver command)"
timeout 600 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("ms/step %.2f  Mvox/s %.2f  e2e %.2f (%.2f ms) launches %d anchor %s cpu %s" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["e2e"]["ms_per_step"], b["gpu_launches"], b.get("anchor", {}).get("ms_per_step"), b.get("cpu_baseline", {}).get("value")))
print("roofline", round(b["roofline"]["frac"], 3), b["roofline"]["ms_per_launch"], [(round(x["frac"], 3), round(x["ms_per_call"], 4)) for x in b["roofline_hbm"]], b["clocks"])
PY
echo "== reference arm (3 steps)"; timeout 300 python3 bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>/dev/null | cut -c1-600
echo "== conv layers"; timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cat gpurun_out/conv_bench.txt
echo "== norms"; timeout 120 python tools/norm_bench.py > gpurun_out/norm_kernels.txt 2>&1; cat gpurun_out/norm_kernels.txt
echo "== eager step profile"; timeout 300 python tools/profile_step.py 2 > gpurun_out/profile_step.log 2>&1; tail -2 gpurun_out/profile_step.log
echo "== ncu full, batch 2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_halo|wgrad_tc|inorm_" --launch-skip 16 --launch-count 8 \
  -o gpurun_out/r02_final_b2 -f python tools/one_kernel.py 2 3 > gpurun_out/ncu_b2.log 2>&1; tail -1 gpurun_out/ncu_b2.log
echo "== ncu full, batch 4"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_halo|wgrad_tc|inorm_" --launch-skip 16 --launch-count 8 \
  -o gpurun_out/r02_final_b4 -f python tools/one_kernel.py 4 3 > gpurun_out/ncu_b4.log 2>&1; tail -1 gpurun_out/ncu_b4.log
echo "== ncu launch list (eager step, the bench command with --no-graphs)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-anchor --no-graphs > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tools/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final.txt 2>&1; head -14 gpurun_out/launches_final.txt
ls -la gpurun_out/*final*
