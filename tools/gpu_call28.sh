#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none --import-source on -k regex:shift_sum --launch-skip 2 --launch-count 1 -o gpurun_out/r02_shift_sum -f python tools/head_once.py > gpurun_out/ncu_ss.log 2>&1; tail -2 gpurun_out/ncu_ss.log
ls -la gpurun_out/r02_shift_sum.ncu-rep
