#!/bin/bash
# 8-GPU call on the final build: config 3 (driver-style command), config 4, config 5.
set -u
source <(sed -n '/^run() {/,/^}/p' tools/gpu_call_dp2.sh)
N=${1:-8}
mkdir -p gpurun_out
run train -- --steps 20 --warmup 5
run unet -- --workload unet --steps 10 --warmup 5
run infer32 -- --workload infer --stride 32 --steps 3 --warmup 1
run infer32_w1 -- --workload infer --stride 32 --steps 3 --warmup 1 --windows-per-pass 1
