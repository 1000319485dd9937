#!/bin/bash
# 8-GPU call: the headline data-parallel run (config 3), config 4 (unet_128) and config 5 (sharded sliding window).
set -u
source <(sed -n '/^run() {/,/^}/p' tools/gpu_call_dp2.sh)
N=${1:-8}
mkdir -p gpurun_out
run train -- --steps 20 --warmup 5
run unet -- --workload unet --steps 10 --warmup 5
run unet_skipnccl MRA_DP_SKIP_ALLREDUCE=1 -- --workload unet --steps 10 --warmup 5
run infer32 -- --workload infer --stride 32 --steps 3 --warmup 1
run infer64 -- --workload infer --stride 64 --steps 3 --warmup 1
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader | head -8
