#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 600 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/gpu_tests.log 2>&1; tail -14 gpurun_out/gpu_tests.log
echo "== nstats bench"; timeout 200 python tools/nstats_bench.py 2 > gpurun_out/nstats_bench.txt 2>&1; cat gpurun_out/nstats_bench.txt
echo "== step table"; timeout 120 python tools/profile_step.py > /dev/null 2>&1; head -40 gpurun_out/step_profile.txt | cut -c1-100,190-260
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
