#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms; with "prof": ncu launch lists, full captures of the
# dominant kernels, reference-CUDA timings and a compute-sanitizer pass over the small parity tests.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -5
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_reference.json
if [ "${1:-}" = "prof5" ]; then
  timeout 200 python bench_splat.py --steps 5 > gpurun_out/bench_splat.json 2>/dev/null; tail -c 300 gpurun_out/bench_splat.json; echo
  timeout 200 python bench_trace.py --steps 5 > gpurun_out/bench_trace.json 2>/dev/null; tail -c 300 gpurun_out/bench_trace.json; echo
  timeout 100 python bench_rays.py --steps 20 > gpurun_out/bench_rays.json 2>/dev/null
  timeout 100 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('second run:', d['value'], d['ms_each_step'])"
fi
if [ "${1:-}" = "prof4" ]; then
  # launch list of the C2 bench command (kernel shares of the step)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log | cut -c1-200
fi
if [ "${1:-}" = "prof3" ]; then
  # added after the r01h captures: point-to-ray search of the in-surface sampler, RayTracing on the value-only SIREN mode
  timeout 200 python bench_rays.py --steps 20 > gpurun_out/bench_rays.json 2>/dev/null; tail -c 700 gpurun_out/bench_rays.json; echo
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ray_nearest_point -s 4 -c 1 -f -o gpurun_out/prof_ray_nearest_point python bench_rays.py --steps 1 > gpurun_out/ncu_rays.log 2>&1; tail -2 gpurun_out/ncu_rays.log
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_offsurface.py tests/test_gpu_rays.py -m gpu -q -k "ties_and_empty or 7-33 or 1-1 or no_ray_hits or golden-siren3" > gpurun_out/sanitizer_memcheck_r01i.log 2>&1; echo "memcheck(r01i) rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_r01i.log
  ls gpurun_out/
fi
if [ "${1:-}" = "prof2" ]; then
  # kernels added after the r01g captures: EWA per-point parameters, the renderable mask, forward-only SIREN
  timeout 300 python bench_splat.py --steps 5 > gpurun_out/bench_splat.json 2>/dev/null
  timeout 300 python bench_trace.py --steps 5 > gpurun_out/bench_trace.json 2>/dev/null; tail -c 600 gpurun_out/bench_trace.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench_splat.py --steps 1 > gpurun_out/ncu_c4.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:point_params_x4 -s 3 -c 1 -f -o gpurun_out/prof_ewa_point_params python bench_splat.py --steps 1 > gpurun_out/ncu_ewa.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:renderable_mask -s 3 -c 1 -f -o gpurun_out/prof_renderable_mask python bench_splat.py --steps 1 > gpurun_out/ncu_mask.log 2>&1
  # first launch of the 4th marching pass (3 warm-up passes x 11 launches skipped): all 200 000 rays live
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_sdf_grad -s 33 -c 1 -f -o gpurun_out/prof_siren_forward_only python bench_trace.py --steps 1 > gpurun_out/ncu_siren_fwd.log 2>&1
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ewa.py tests/test_gpu_trace.py -m gpu -q -k "not 40000 and not views0 and not views3" > gpurun_out/sanitizer_memcheck_new.log 2>&1; echo "memcheck(new) rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_new.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_trace.py tests/test_gpu_ewa.py -m gpu -q -k "forward_only_step or trace_step_kernel or unaligned or reference_golden" > gpurun_out/sanitizer_racecheck_new.log 2>&1; echo "racecheck(new) rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_new.log
  ls gpurun_out/
fi
if [ "${1:-}" = "prof" ]; then
  timeout 500 python scripts/ref_cuda_timing.py 2>&1 | grep -v Warning | tail -1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench_splat.py --steps 1 > gpurun_out/ncu_c4.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench_frnn.py --steps 1 > gpurun_out/ncu_c3.log 2>&1
  # the fused SIREN kernel: first launch of the timed step (all 200 000 rows live); 3 warm-up steps x 15 launches are skipped
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_sdf_grad -s 45 -c 1 -f -o gpurun_out/prof_siren python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_siren.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:frnn_query -s 3 -c 1 -f -o gpurun_out/prof_frnn_query_c2 python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_frnn_c2.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:frnn_query -s 3 -c 1 -f -o gpurun_out/prof_frnn_query python bench_frnn.py --steps 1 > gpurun_out/ncu_frnn.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_raster_v2_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_raster python bench_splat.py --steps 1 > gpurun_out/ncu_raster.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_occ_backward_hybrid_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_occ_bwd python bench_splat.py --steps 1 > gpurun_out/ncu_occ.log 2>&1
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_splat.py tests/test_gpu_frnn.py tests/test_gpu_projection.py tests/test_gpu_siren.py -m gpu -q -k "not 500k and not c4_scale and not c2_scale and not world1 and not match_fp64 and not projection_with_fused and not project_resample_and_ragged" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_splat.py -m gpu -q -k "bit_exact_vs_oracle or backward_matches_oracle or blend" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_siren.py -m gpu -q -k "layer_counts or fused_newton" > gpurun_out/sanitizer_racecheck_siren.log 2>&1; echo "racecheck(siren) rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_siren.log
  ls gpurun_out/
fi
