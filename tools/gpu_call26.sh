#!/bin/bash
# Final evidence of the round on one box: GPU suite (default path), eager / eager / graphs loss spread behind the tolerance of
# test_cuda_graph_step_matches_eager, smoke, the driver's command (anchor + cpu baseline), ncu launch list of one eager step.
set -u
mkdir -p gpurun_out
t0=$SECONDS
echo "== gpu tests";  timeout 300 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
echo "   t=$((SECONDS-t0))s"
echo "== eager vs eager vs graphs"; timeout 120 python tools/graph_compare.py > gpurun_out/graph_compare.txt 2>&1; head -9 gpurun_out/graph_compare.txt | cut -c1-150
echo "== smoke";      timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "   t=$((SECONDS-t0))s"
echo "== bench (driver command)"
timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
python - <<'PY'
import json
try:
    b = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
    print('ms/step %.2f Mvox/s %.2f e2e %.2f launches %d clocks %s' % (b['ms_per_step'], b['value']/1e6, b['e2e']['value']/1e6, b['gpu_launches'], b['clocks']))
    print('roofline frac %.3f ms %.4f traffic %s' % (b['roofline']['frac'], b['roofline']['ms_per_launch'], b['roofline']['traffic']))
    print('hbm', [(round(x['frac'], 3), round(x['ms_per_call'], 4)) for x in b['roofline_hbm']]); print('cpu', b['cpu_baseline']['value'], b['cpu_baseline']['cores']); print('anchor', b['anchor']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
echo "   t=$((SECONDS-t0))s"
echo "== ncu launch list (eager step, the bench command with --no-graphs)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-anchor --no-graphs > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tools/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final.txt 2>&1; head -16 gpurun_out/launches_final.txt | cut -c1-160
echo "   t=$((SECONDS-t0))s"
