#!/bin/bash
set -u
source <(sed -n '/^run() {/,/^}/p' tools/gpu_call_dp2.sh)
N=${1:-2}
mkdir -p gpurun_out
run infer32 -- --workload infer --stride 32 --steps 2 --warmup 1
python -c "
import json; b=json.loads(open('gpurun_out/dp${N}_infer32.json').read().strip().splitlines()[-1]); print(b['sharded_vs_single_max_abs'], b['batched_vs_one_window_per_pass'])"
run train -- --steps 10 --warmup 5
