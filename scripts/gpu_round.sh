#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch lists and full captures of the
# dominant kernels.  Outputs land in gpurun_out/ (copied into profiles/ by hand after reading).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -5
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/bench_reference.json
if [ "${1:-}" = "prof" ]; then
  # launch lists (cold-cache, serialised: compare shares)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench_splat.py --steps 1 > gpurun_out/ncu_c4.log 2>&1
  # full captures of the dominant kernels
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:frnn_query_kernel -s 3 -c 1 -f -o gpurun_out/prof_frnn_query python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_frnn.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_raster_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_raster python bench_splat.py --steps 1 > gpurun_out/ncu_raster.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_occ_backward_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_occ_bwd python bench_splat.py --steps 1 > gpurun_out/ncu_occ.log 2>&1
  ls -la gpurun_out/
fi
