#!/bin/bash
# 2-GPU call: data-parallel variants of the training step (eager / graphs / fused wgrad / no all-reduce), config 4 and
# config 5 at N = 2.  Every run under its own timeout; lines land in gpurun_out/dp2_*.json.
set -u
mkdir -p gpurun_out
N=${1:-2}
run() {   # name, env..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "== $name (${envs[*]:-}) $*"
  env "${envs[@]}" timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/dp${N}_$name.json 2> gpurun_out/dp${N}_$name.err
  echo "rc=$?"
  python - "$name" "$N" <<'PY'
import json, sys
name, n = sys.argv[1], sys.argv[2]
try:
    b = json.loads(open("gpurun_out/dp%s_%s.json" % (n, name)).read().strip().splitlines()[-1])
    print("   %-14s ms/step %8.2f  Mvox/s %8.2f  e2e %8.2f  launches %5d  clocks %s  %s %s" % (
        name, b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], (b.get("clocks") or {}).get("sm_mhz"),
        b["config"].get("launch"), json.dumps(b.get("grad_sync") or b.get("sharded_vs_single_max_abs"))))
except Exception as e:
    print("   %s: no line (%s)" % (name, e))
    import subprocess
    print(subprocess.run("tail -5 gpurun_out/dp%s_%s.err" % (n, name), shell=True, capture_output=True, text=True).stdout)
PY
}
echo "== gpu tests"; timeout 600 python -m pytest tests -q -m gpu --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -6 gpurun_out/gpu_tests.log
run eager -- --steps 10 --warmup 5
run graphs MRA_DP_GRAPHS=1 -- --steps 10 --warmup 5
run fusedw MRA_DP_FUSED_WGRAD=1 -- --steps 10 --warmup 5
run graphs_fusedw MRA_DP_GRAPHS=1 MRA_DP_FUSED_WGRAD=1 -- --steps 10 --warmup 5
run skipnccl MRA_DP_SKIP_ALLREDUCE=1 -- --steps 10 --warmup 5
run unet -- --workload unet --steps 10 --warmup 5
run unet_skipnccl MRA_DP_SKIP_ALLREDUCE=1 -- --workload unet --steps 10 --warmup 5
run unet_fusedw MRA_DP_FUSED_WGRAD=1 -- --workload unet --steps 10 --warmup 5
run unet_graphs_fusedw MRA_DP_GRAPHS=1 MRA_DP_FUSED_WGRAD=1 -- --workload unet --steps 10 --warmup 5
run infer32 -- --workload infer --stride 32 --steps 2 --warmup 1
run infer64 -- --workload infer --stride 64 --steps 2 --warmup 1
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
