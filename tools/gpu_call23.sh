#!/bin/bash
# Re-entry validation of HEAD on a fresh box (the container was re-created, the library rebuilt from source):
# GPU suite, smoke(), the driver's exact single-GPU command (full line with anchor + cpu_baseline), reference arm.
set -u
mkdir -p gpurun_out
t0=$SECONDS
echo "== gpu tests";  timeout 300 python -m pytest tests -q -m gpu -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
echo "   t=$((SECONDS-t0))s"
echo "== smoke";      timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "   t=$((SECONDS-t0))s"
echo "== bench (driver command)"
timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
python - <<'EOF'
import json
try:
    b = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
    print('ms/step %.2f Mvox/s %.2f e2e %.2f launches %d clocks %s' % (b['ms_per_step'], b['value']/1e6, b['e2e']['value']/1e6, b['gpu_launches'], b['clocks']))
    print('roofline', b['roofline']); print('roofline_hbm', b.get('roofline_hbm')); print('cpu_baseline', b.get('cpu_baseline')); print('anchor', b.get('anchor'))
except Exception as e:
    print('bench parse failed', e)
EOF
echo "   t=$((SECONDS-t0))s"
echo "== reference arm"
timeout 200 python3 bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-500 gpurun_out/bench_ref.json
echo "   t=$((SECONDS-t0))s"
