#!/bin/bash
# Round-2 first GPU call: the statistics-flush fix under compute-sanitizer, the whole GPU suite, and the DRIVER's
# exact single-GPU bench command (VERDICT r1 "next round" item 1).
set -u
mkdir -p gpurun_out
echo "== gpu tests";  timeout 400 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
echo "== memcheck (halo statistics canary + tail-item test)"
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_halo.log \
  python -m pytest tests/test_zz_fullsize.py -q -m gpu -k "halo_statistics or halo_tail" > gpurun_out/memcheck_pytest.log 2>&1
tail -3 gpurun_out/memcheck_pytest.log; tail -4 gpurun_out/memcheck_halo.log
echo "== driver bench command"
timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== reference arm (short)"
timeout 200 python3 bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
nproc; free -g | head -2
