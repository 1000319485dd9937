#!/bin/bash
set -u
for L in "128 64 3 2 1 1 1 64 fprop" "64 128 3 2 1 0 0 128 dgrad" "256 128 3 2 1 1 1 32 fprop" "128 256 3 2 1 0 0 64 dgrad"; do
  set -- $L; which=${9}; args="$1 $2 $3 $4 $5 $6 $7 $8 2"
  echo "== layer $args $which"
  echo -n "default      "; ONLY=$which timeout 60 python tools/layer_bench.py $args 2>&1 | tail -1
  for np in 2 3 4 6; do
    echo -n "phase NP=$np   "; MRA_GATHER_PHASE=1 MRA_PHASE_NP=$np ONLY=$which timeout 60 python tools/layer_bench.py $args 2>&1 | tail -1
  done
done
