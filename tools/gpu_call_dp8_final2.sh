#!/bin/bash
# 8-GPU call on the final build (after split-K / stream-K start / flush changes): config 3 and config 4.
set -u
source <(sed -n '/^run() {/,/^}/p' tools/gpu_call_dp2.sh)
N=${1:-8}
mkdir -p gpurun_out
run train -- --steps 20 --warmup 5
run unet -- --workload unet --steps 10 --warmup 5
run infer32 -- --workload infer --stride 32 --steps 3 --warmup 1
