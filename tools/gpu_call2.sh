#!/bin/bash
# Round-2 second GPU call: norm-backward statistics from the dgrad epilogue (tests + A/B of the step), per-layer conv
# table, ncu --set full of the G.rb kernels at batch 2 and 4.
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 400 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
echo "== bench (fused norm bwd)"
timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_fused.json",):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.2f  Mvox/s %.2f  e2e %.2f  launches %d  anchor %.2f ms" % (b["ms_per_step"], b["value"] / 1e6, b["e2e"]["value"] / 1e6, b["gpu_launches"], b.get("anchor", {}).get("ms_per_step", 0)))
        for r in b["roofline_hbm"]: print("   %.0f GB/s %.3f  %.3f ms  %s" % (r["achieved"], r["frac"], r["ms_per_call"], r["kernel"][:60]))
        print("   roofline", b["roofline"]["achieved"], b["roofline"]["frac"], b["clocks"])
    except Exception as e: print(f, "unreadable", e)
PY
echo "== bench (two-pass norm bwd, ablation)"
MRA_NORM_BWD_FUSED=0 timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor > gpurun_out/bench_twopass.json 2> gpurun_out/bench_twopass.err; echo "rc=$?"
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_twopass.json").read().strip().splitlines()[-1])
print("two-pass: ms/step %.2f  Mvox/s %.2f launches %d" % (b["ms_per_step"], b["value"] / 1e6, b["gpu_launches"]))
PY
echo "== conv layers"; timeout 200 python tools/conv_bench.py 2 > /dev/null 2>&1; cat gpurun_out/conv_bench.txt
echo "== norms"; timeout 120 python tools/norm_bench.py > gpurun_out/norm_kernels.txt 2>&1; cat gpurun_out/norm_kernels.txt
echo "== ncu full, batch 2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_halo|wgrad_tc|inorm_" --launch-skip 16 --launch-count 8 \
  -o gpurun_out/r02_full_b2 -f python tools/one_kernel.py 2 3 > gpurun_out/ncu_b2.log 2>&1; tail -2 gpurun_out/ncu_b2.log
echo "== ncu full, batch 4"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gather_halo|wgrad_tc|inorm_" --launch-skip 16 --launch-count 8 \
  -o gpurun_out/r02_full_b4 -f python tools/one_kernel.py 4 3 > gpurun_out/ncu_b4.log 2>&1; tail -2 gpurun_out/ncu_b4.log
ls -la gpurun_out/*.ncu-rep
