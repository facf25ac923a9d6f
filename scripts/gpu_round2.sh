#!/bin/bash
# Round-2 GPU-box visits.  usage: scripts/gpu_round2.sh <stage> ; outputs under gpurun_out/
set -u
mkdir -p gpurun_out
case "${1:-all}" in
  tests)
    timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log ;;
  newtests)
    timeout 900 python -m pytest tests/test_gpu_pinned.py tests/test_gpu_dropin.py tests/test_gpu_pointops.py tests/test_gpu_offsurface.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/pytest_new.log; tail -40 gpurun_out/pytest_new.log ;;
  timeline)
    timeout 300 python scripts/siren_timeline.py > gpurun_out/siren_timeline.txt 2>&1; tail -70 gpurun_out/siren_timeline.txt ;;
  bench)
    timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -5 ;;
  refarm)
    timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/bench_reference.json ;;
esac
