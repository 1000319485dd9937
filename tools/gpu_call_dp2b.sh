#!/bin/bash
# 2-GPU call: the fused-wgrad data-parallel variants after the bucket-alignment fix (+ unet with graphs).
set -u
source <(sed -n '/^run() {/,/^}/p' tools/gpu_call_dp2.sh)
N=${1:-2}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 600 python -m pytest tests -q -m gpu --durations=3 > gpurun_out/gpu_tests.log 2>&1; tail -6 gpurun_out/gpu_tests.log
run fusedw MRA_DP_FUSED_WGRAD=1 -- --steps 10 --warmup 5
run graphs_fusedw MRA_DP_GRAPHS=1 MRA_DP_FUSED_WGRAD=1 -- --steps 10 --warmup 5
run unet_graphs MRA_DP_GRAPHS=1 -- --workload unet --steps 10 --warmup 5
run unet_fusedw MRA_DP_FUSED_WGRAD=1 -- --workload unet --steps 10 --warmup 5
run unet_graphs_fusedw MRA_DP_GRAPHS=1 MRA_DP_FUSED_WGRAD=1 -- --workload unet --steps 10 --warmup 5
