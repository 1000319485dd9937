#!/bin/bash
# One gpurun call that re-establishes the measured state of the tree on a fresh B200 box (≈6-7 GPU-minutes):
#   gpurun --timeout 900 -- 'bash tools/round_start.sh'
# Everything lands in gpurun_out/ (copy what is to be judged into profiles/).  Nothing here runs under a profiler
# except the last step, whose printed numbers are not bench values.
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 240 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== smoke";      timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (the driver's own command: 25 steps at batch 2 fill the 50-image pools)"
timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench.json
echo "== reference arm"; timeout 300 python3 bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
echo "== step table"; timeout 120 python tools/profile_step.py > /dev/null 2>&1; head -25 gpurun_out/step_profile.txt | cut -c60-200
echo "== norms";      timeout 120 python tools/norm_bench.py > gpurun_out/norm_kernels.txt 2>&1; cat gpurun_out/norm_kernels.txt
echo "== configs 4/5"; timeout 180 python tools/config_check.py > gpurun_out/configs_4_5.txt 2>&1; cut -c1-200 gpurun_out/configs_4_5.txt
echo "== ncu launch list (one eager step)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -12 gpurun_out/launches.txt
