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
  sanitize)
    # compute-sanitizer over the kernels written or rewritten this round (small parity cases)
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_frnn.py tests/test_gpu_splat.py tests/test_gpu_pointops.py tests/test_gpu_siren.py -m gpu -q -x -k "small_bit_exact or clustered or fused_blend or tiled_sweep or backward_matches_oracle or all_sms or layer_counts or fused_newton or device_side" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_frnn.py tests/test_gpu_splat.py tests/test_gpu_siren.py -m gpu -q -x -k "clustered or fused_blend or tiled_sweep or layer_counts" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
    timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_siren.py tests/test_gpu_pointops.py -m gpu -q -x -k "layer_counts or all_sms" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/sanitizer_synccheck.log ;;
  sanitize2)
    # the kernels touched at the end of the round: capped grid parameters + bounding box, median passes, the tile
    # count kernel's scan tail, the three-CTA raster budget, the run-ahead projection step
    timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_frnn.py tests/test_gpu_splat.py tests/test_gpu_pinned.py -m gpu -q -x -k "small_bit_exact or capped_grid or cell_table or nothing_converges or 30000 or fused_blend or edge_cases or dense_overdraw or backward_matches_oracle or tiled_sweep or (c4_scale and raster_v2_default)" > gpurun_out/sanitizer2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer2_memcheck.log
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_splat.py tests/test_gpu_frnn.py -m gpu -q -x -k "fused_blend or dense_overdraw or edge_cases or capped_grid" > gpurun_out/sanitizer2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer2_racecheck.log
    timeout 300 python scripts/step_gaps.py > gpurun_out/step_gaps.txt 2>&1; tail -32 gpurun_out/step_gaps.txt ;;
  quick)
    timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
    timeout 300 python bench.py --steps 10 --warmup 3 --no-side | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2', round(d['ms_per_step'],3), round(d['value']/1e6,3), d['clocks']); [print('  ', k, round(v['ms_per_step'],4)) for k, v in d['kernels'].items()]; print('   glue', d['sdf_callback_and_glue_ms_per_step'])"
    timeout 300 python bench_frnn.py --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3', d['ms_per_step'], d['value'], d['kernels_avg_ms'])"
    timeout 300 python bench_splat.py --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4', d['ms_fwd'], d['ms_fwd_blend_bwd'], d['value']); [print('  ', k, v) for k, v in d['kernels'].items()]" ;;
  pointops)
    timeout 600 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_frnn.py -m gpu -q -x 2>&1 | tail -8
    timeout 300 python bench_frnn.py --steps 10 > gpurun_out/bench_frnn.json 2>gpurun_out/bench_frnn.err; python -c "import json; d=json.load(open('gpurun_out/bench_frnn.json')); print(d['ms_per_step'], d['value'], d['kernels_avg_ms'])"
    timeout 600 python bench_pointops.py > gpurun_out/bench_pointops.json 2> gpurun_out/bench_pointops.err; cat gpurun_out/bench_pointops.json; tail -3 gpurun_out/bench_pointops.err ;;
  frnn)
    timeout 600 python -m pytest tests/test_gpu_frnn.py tests/test_gpu_pointops.py tests/test_gpu_projection.py -m gpu -q -x 2>&1 | tail -15
    timeout 300 python bench_frnn.py --steps 10 > gpurun_out/bench_frnn.json 2>gpurun_out/bench_frnn.err; python -c "import json; d=json.load(open('gpurun_out/bench_frnn.json')); print(d['ms_per_step'], d['value'], d['kernels_avg_ms'])"
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:frnn_query -s 3 -c 1 -f -o gpurun_out/prof_frnn_query python bench_frnn.py --steps 1 > gpurun_out/ncu_frnn.log 2>&1; tail -2 gpurun_out/ncu_frnn.log ;;
  splat)
    timeout 600 python -m pytest tests/test_gpu_splat.py tests/test_gpu_ewa.py tests/test_gpu_offsurface.py -m gpu -q -x 2>&1 | tail -15
    timeout 300 python bench_splat.py --steps 10 > gpurun_out/bench_splat.json 2>gpurun_out/bench_splat.err; python -c "import json; d=json.load(open('gpurun_out/bench_splat.json')); print(d['ms_fwd'], d['ms_fwd_blend_bwd'], d['value']); [print(k, v) for k, v in d['kernels'].items()]" ;;
  stagger)
    # sustained C2 step with the phase stagger of the SIREN kernel off / on, alternating on one box
    for r in 1 2 3; do
      for st in 0 -1; do
        ISOB200_SIREN_STAGGER=$st timeout 300 python bench.py --steps 20 --warmup 3 --no-side | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger $st', round(d['ms_per_step'],3), round(d['value']/1e6,3), round(d['config']['sdf_evaluations_per_s']/1e6,2), d['clocks']['sm_mhz'], d['clocks']['power_w'], d['kernels']['siren_project_step']['ms_per_step'])"
      done
    done ;;
  ab)
    # old (build/ab/libisob200_old.so: `git worktree` of an earlier commit, python -m isopoints_b200.build there, copy the
    # .so) vs new library, alternating on the same box: sustained C2 step
    for r in 1 2 3; do
      for which in old new; do
        if [ $which = old ]; then export ISOB200_LIB=$PWD/build/ab/libisob200_old.so; else unset ISOB200_LIB; fi
        timeout 300 python bench.py --steps 20 --warmup 3 --no-side | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$which', round(d['ms_per_step'],3), round(d['value']/1e6,3), round(d['config']['sdf_evaluations_per_s']/1e6,2), d['clocks']['sm_mhz'], d['kernels']['siren_project_step']['ms_per_step'])"
      done
    done ;;
  siren2)
    timeout 300 python -m pytest tests/test_gpu_siren.py tests/test_gpu_trace.py -m gpu -q -x 2>&1 | tail -3
    ISOB200_SIREN_PAIR_HANDOFF=1 timeout 300 python -m pytest tests/test_gpu_siren.py tests/test_gpu_trace.py tests/test_gpu_pinned.py -m gpu -q -x 2>&1 | tail -3
    timeout 300 python scripts/siren_timeline.py > gpurun_out/siren_timeline.txt 2>&1; head -24 gpurun_out/siren_timeline.txt
    ISOB200_SIREN_PAIR_HANDOFF=1 timeout 300 python scripts/siren_timeline.py > gpurun_out/siren_timeline_pair.txt 2>&1; tail -52 gpurun_out/siren_timeline_pair.txt
    for p in 0 1 0 1; do ISOB200_SIREN_PAIR_HANDOFF=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-side | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair $p', d['ms_per_step'], d['value'], d['config']['sdf_evaluations_per_s'], d['clocks'])"; done ;;
  siren)
    timeout 600 python -m pytest tests/test_gpu_siren.py tests/test_gpu_trace.py tests/test_gpu_pinned.py tests/test_gpu_projection.py tests/test_gpu_rays.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_siren.log; tail -30 gpurun_out/pytest_siren.log
    timeout 300 python scripts/siren_timeline.py > gpurun_out/siren_timeline.txt 2>&1; cat gpurun_out/siren_timeline.txt
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:siren_sdf_grad --csv --log-file gpurun_out/siren_cta_sweep_ncu.csv python scripts/siren_cta_sweep.py --once > /dev/null 2>&1
    cut -d, -f 5,13- gpurun_out/siren_cta_sweep_ncu.csv | tail -21 ;;
  sweep)
    timeout 300 python scripts/siren_cta_sweep.py --stamps > gpurun_out/siren_cta_sweep.txt 2>&1
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:siren_sdf_grad --csv --log-file gpurun_out/siren_cta_sweep_ncu.csv python scripts/siren_cta_sweep.py --once > /dev/null 2>&1
    cat gpurun_out/siren_cta_sweep.txt; cut -d, -f 5,13- gpurun_out/siren_cta_sweep_ncu.csv | tail -30 ;;
  prof)
    # r02 evidence: tests, benches, per-stage stamps, launch lists (C2 / C3 / C4), full captures of the dominant kernels
    timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
    timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -3
    timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null
    timeout 200 python bench_splat.py --steps 10 > gpurun_out/bench_splat.json 2>/dev/null
    timeout 200 python bench_frnn.py --steps 10 > gpurun_out/bench_frnn.json 2>/dev/null
    timeout 200 python bench_trace.py --steps 5 > gpurun_out/bench_trace.json 2>/dev/null
    timeout 100 python bench_rays.py --steps 20 > gpurun_out/bench_rays.json 2>/dev/null
    timeout 300 python bench_pointops.py > gpurun_out/bench_pointops.json 2>/dev/null
    timeout 300 python scripts/siren_timeline.py > gpurun_out/siren_timeline.txt 2>&1
    timeout 300 python scripts/siren_cta_sweep.py > gpurun_out/siren_cta_sweep.txt 2>&1
    export ISO_BENCH_PREROLL=3
    B="python bench.py --steps 1 --warmup 3 --no-side"
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2.csv $B > gpurun_out/ncu_c2.log 2>&1
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench_splat.py --steps 1 > gpurun_out/ncu_c4.log 2>&1
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench_frnn.py --steps 1 > gpurun_out/ncu_c3.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_sdf_grad -s 45 -c 1 -f -o gpurun_out/prof_siren $B > gpurun_out/ncu_siren.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:frnn_query -s 3 -c 1 -f -o gpurun_out/prof_frnn_query_c2 $B > gpurun_out/ncu_frnn_c2.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:frnn_query -s 3 -c 1 -f -o gpurun_out/prof_frnn_query python bench_frnn.py --steps 1 > gpurun_out/ncu_frnn.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_raster_v2_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_raster python bench_splat.py --steps 1 > gpurun_out/ncu_raster.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_occ_backward_tiled_kernel -s 3 -c 1 -f -o gpurun_out/prof_splat_occ_bwd python bench_splat.py --steps 1 > gpurun_out/ncu_occ.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_coop_kernel -s 1 -c 1 -f -o gpurun_out/prof_fps python bench_pointops.py > gpurun_out/ncu_fps.log 2>&1
    ls -la gpurun_out/ | tail -40 ;;
  bench)
    timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -5 ;;
  refarm)
    timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/bench_reference.json ;;
esac
