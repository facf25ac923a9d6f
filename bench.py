#!/usr/bin/env python
"""bench.py -- iso-points hot path on B200: one JSON line per run (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): 200 000 points per GPU, (rand-0.5)*2, and the SDF SURVEY 8d
pins: the reference's own Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, omega 30, linear head) as
constructed right after torch.manual_seed(0) (tests/helpers.pinned_siren; state_dict equality with the
reference class is a CPU test) -- 60 % of the points converge on this random-init field, 7.5 SDF
evaluations per input point (--sdf siren, default: evaluated by the package's fused tcgen05 kernel;
--sdf opaque: the same weights behind an opaque nn.Module, i.e. the autograd callback), UniformProjection(proj_max_iters=10, tol=5e-5, knn_k=8, sample_iters=1)
.project_points(x, sdf, skip_upsampling=True)  = project -> filter -> FRNN(K=9) -> resample ->
3-iteration re-projection.  A "step" is one such pass over one synthetic cloud.
  value : input iso-points / second, whole job, inputs resident in HBM, CUDA-event timed;
  e2e   : same through the public API with HOST buffers (pinned H2D of the cloud, D2H of
          points+normals+mask inside the timed region);
  splat : second headline metric (pixel-splats/s, BASELINE configs[3] "C4") when the splat
          kernels are built, reported in the same line under "splat".
N > 1 (torchrun, one rank per GPU): the cloud is sharded by contiguous point ranges (weak scaling,
200 000 points per rank); projection is embarrassingly parallel, the resample exchanges projected
positions + normals with one all-gather so that every rank searches the full cloud.
Before the K timed steps: max(W, 50) untimed steps with exactly the timed steps' instrumentation (L2 flush,
per-call CUDA events), so that event pools, NVML and the GPU clocks are in steady state -- `ms_each_step` in
the line shows every timed step, `config.untimed_steps_before` the count.
`--impl reference` times the CPU restatement of the reference path (oracle/port.py; the path is
Python + third-party CUDA-only FRNN, so the reference itself cannot run on host cores) on rank 0.
Outside every timed region the line also gains: `ref_cuda` (N = 1: the reference's own Python on its own CUDA
extensions recompiled for sm_100a, oracle/ref_gpu.py in a subprocess, C2 / C3 / C4 -- the GPU-vs-GPU baseline),
`dist_parity` (N > 1: the sharded result equals the single-GPU operator on the concatenated cloud) and `c5`
(BASELINE configs[4], bench_c5.py: the fixed 2 M-point project + resample + splat 16 x 1024^2 problem at this N).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _ncu(capture):
    """Record of the committed `ncu --set full` capture `capture` (profiles/ncu_traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[capture]
    except Exception:
        return {}


def _ncu_traffic(capture):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from that capture, or None."""
    return _ncu(capture).get("dram_bytes_per_launch")



def _issue_ceiling(capture, launch_ms, sm_mhz=1965.0):
    """Ceiling of an instruction-issue-bound kernel: 148 SMs x 4 schedulers x 1 warp instruction per cycle; warp
    instructions per launch from the committed `ncu --set full` capture, launch time measured live."""
    n = _ncu(capture).get("warp_inst")
    if not n or not launch_ms:
        return None
    peak = 148 * 4 * sm_mhz * 1e6 / 1e9
    ach = n / (launch_ms * 1e-3) / 1e9
    return {"warp_inst_per_launch": n, "achieved_ginst_s": ach, "peak_ginst_s": peak, "frac": ach / peak,
            "peak_is": "148 SMs x 4 schedulers x %.0f MHz (max SM clock)" % sm_mhz}


C2_POINTS = 200_000
C2_WORKLOAD = ("C2: 200000 pts/GPU, SURVEY-pinned SIREN 8x256 SDF (reference Siren(n_layers=7) under "
               "torch.manual_seed(0)), project(10 it)+resample(knn_k=8, 1 it)+reproject(3 it)")
L2_FLUSH_BYTES = 256 << 20


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


PREROLL_STEPS = int(os.environ.get("ISO_BENCH_PREROLL", "50"))   # minimum number of untimed steps before the timed region (clock ramp of an idle GPU)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.

    Uses NVML in-process (nvidia_ml_py): forking `nvidia-smi` from a process that holds a CUDA context
    can stall the launching thread for tens of milliseconds -- visible as an outlier step now that the
    Newton loop is enqueued without read-backs and the GPU depends on the host keeping its queue full.
    `nvidia-smi` remains the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0, period=0.004):
        self.sm, self.mx, self.reasons, self.pw = [], [], set(), []
        self.gpu = gpu_index
        self.period = period
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nv = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            try:
                self._plimit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self._h) / 1000.0
            except Exception:
                self._plimit = None
        except Exception:
            self._h = None

    def _sample_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        self.mx.append(self._max)
        try:
            self.pw.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for name, bit in self.REASONS:
            if bits & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [c.strip() for c in line.split(",")]
            self.sm.append(float(r[1])); self.mx.append(float(r[2]))
            try:
                self.pw.append(float(r[3]))
            except Exception:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(self.period if self._h is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def reset(self):
        """Drop what was sampled so far: the sampler is started before the warm-up steps (the first NVML queries
        of a process initialise driver state and were seen to stall the first step after them by tens of ms) and
        reset when the timed region starts, so only samples taken during it are reported."""
        self.sm, self.mx, self.reasons, self.pw = [], [], set(), []

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)),
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w": float(np.median(self.pw)) if self.pw else None,
                "power_w_max": float(max(self.pw)) if self.pw else None,
                "power_limit_w": getattr(self, "_plimit", None),
                "source": "nvml" if self._h is not None else "nvidia-smi"}


def _dist_init(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, dev):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _make_c2(rank, dev, opaque=False):
    """C2 cloud + SDF.  `pinned_siren(0)` = the reference decoder (DSS/models/common.py:90-165) exactly as
    SURVEY 8d pins it (torch.manual_seed(0), default nn.Linear init then uniform_ on the weights), in the
    structure the package evaluates with its fused tcgen05 kernel; `.as_opaque()` holds the same weights
    behind an opaque nn.Module (autograd path)."""
    from tests.helpers import pinned_siren
    g = torch.Generator().manual_seed(1000 + rank)
    x = (torch.rand(1, C2_POINTS, 3, generator=g) - 0.5) * 2
    net = pinned_siren(0)
    return x, (net.as_opaque() if opaque else net)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    from isopoints_b200 import _ext
    from isopoints_b200.levelset_sampling import UniformProjection
    rank, world, local = _dist_init(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False      # parity bar is fp32 1e-4 (SIREN x30 gain)
    torch.backends.cudnn.allow_tf32 = False
    lib = _ext.lib()
    x_host, net = _make_c2(rank, dev, opaque=args.sdf == "opaque")
    net = net.to(dev)
    x_dev = x_host.to(dev)
    x_pin = x_host.pin_memory()
    if world > 1:
        from isopoints_b200.dist import ShardedUniformProjection
        proj = ShardedUniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    else:
        proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def step(x):
        return proj.project_points(x, net, skip_upsampling=True)

    # pinned host buffers for the results, allocated once (outputs have at most C2_POINTS rows)
    pin = {"levelset_points": torch.empty((1, C2_POINTS, 3), dtype=torch.float32).pin_memory(),
           "levelset_normals": torch.empty((1, C2_POINTS, 3), dtype=torch.float32).pin_memory(),
           "mask": torch.empty((1, C2_POINTS), dtype=torch.bool).pin_memory()}

    def step_e2e():
        x = x_pin.to(dev, non_blocking=True)
        out = step(x)
        res = []
        for k in ("levelset_points", "levelset_normals", "mask"):
            dst = pin[k][:, :out[k].shape[1]]
            dst.copy_(out[k], non_blocking=True)
            res.append(dst)
        torch.cuda.synchronize()
        return out, tuple(res)

    clocks = ClockSampler(local)
    clocks.__enter__()            # sampling thread runs through the warm-up too; reset() below
    # warm-up steps carry the instrumentation of the timed ones (L2 flush, per-call CUDA events, row-count
    # records): the first step that created those events was seen to take 40-50 ms on a fresh box
    from isopoints_b200 import siren as _siren
    _ext.PROFILE = {}
    _siren.RECORD = []
    # ... and there are at least PREROLL_STEPS of them (~0.4 s; the same count on every rank -- the sharded step
    # has collectives), so that a GPU that idled at low clocks on a fresh box has ramped up before the first
    # timed step (seen: 13 ms first step of the first process on a box against 8.0-8.5 ms afterwards)
    for k in range(max(args.warmup, PREROLL_STEPS)):
        flush.fill_(k & 0xff)
        wa, wb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wa.record()
        out = step(x_dev)
        wb.record()
        torch.cuda.synchronize()
        _ext.PROFILE = {}
        _siren.RECORD = []
    _barrier(world)

    # ---- timed region: K steps, device resident inputs, L2 flushed between steps ----------
    _ext.PROFILE = {}
    for k_ in _siren.STATS:
        _siren.STATS[k_] = 0
    _siren.RECORD = []
    l0 = lib.isob200_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    try:
        _barrier(world)
        clocks.reset()
        t0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 0xff)
            ev[k][0].record()
            out = step(x_dev)
            ev[k][1].record()
        _barrier(world)
        wall = time.perf_counter() - t0
    finally:
        clocks.__exit__(None, None, None)
    launches = lib.isob200_launch_count() - l0
    prof = _ext.PROFILE
    _ext.PROFILE = None
    _siren.resolve_record()
    _siren.RECORD = None
    siren_stats = dict(_siren.STATS)
    siren_stats["flops"] = _siren.algorithmic_flops(siren_stats["rows"], 7)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms = sum(step_ms) / args.steps
    ms = _max_over_ranks(ms, world, dev)
    converged = float(out["mask"].float().mean())
    n_out = int(out["mask"].shape[1])

    kern = {}
    for name, pairs in prof.items():
        ts = [a.elapsed_time(b) for a, b in pairs]
        kern[name.replace("isob200_", "")] = {"calls_per_step": len(ts) / args.steps,
                                              "ms_per_step": sum(ts) / args.steps,
                                              "avg_ms": sum(ts) / len(ts)}
    own_ms = sum(v["ms_per_step"] for v in kern.values())

    # ---- e2e: host buffers, H2D + D2H inside the timed region ------------------------------
    # (same per-call event instrumentation as the device-timed loop above, so the two numbers are comparable)
    _ext.PROFILE = {}
    _siren.RECORD = []
    for _ in range(2):
        step_e2e()
        _ext.PROFILE = {}
        _siren.RECORD = []
    _barrier(world)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_e2e = 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        o, res = step_e2e()
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t1
    e2e_ms = _max_over_ranks(t_e2e / args.steps * 1e3, world, dev)
    _ext.PROFILE = None
    _siren.RECORD = None
    # ---- outside every timed region: the projection's own converged fraction (SURVEY 8d asks for it) --------
    p0 = proj._project_points(net, x_dev, torch.tensor([C2_POINTS], device=dev), num_points_list=[C2_POINTS])
    converged_projection = float(p0.mask.float().mean())
    torch.cuda.synchronize()
    h2d = x_pin.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in res)

    peaks, peak_src = _peaks()
    sd = kern.get("siren_project_step") or kern.get("siren_sdf_grad")
    q = kern.get("frnn_find_nbrs")
    roof = None
    frnn_roof = None
    if q:
        # the FRNN query (K = knn_k + 1 = 9, int64 idx + f32 dist out); SURVEY 8d: 16 + 12K bytes per query
        K = 9
        alg = n_out * (16 + 12 * K)
        achieved = alg / (q["avg_ms"] * 1e-3) / 1e9
        frnn_roof = {"kernel": "frnn_query_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": _ncu_traffic("prof_frnn_query_c2"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "avg_launch_ms": q["avg_ms"],
                     "limiter": "instruction issue, not HBM (candidates are served by L1/L2)",
                     "issue": _issue_ceiling("prof_frnn_query_c2", q["avg_ms"]),
                     "ncu": _ncu("prof_frnn_query_c2")}
        roof = frnn_roof
    if sd:
        # dominant kernel of the step: the fused SIREN SDF + gradient kernel (tensor-core bound).
        # achieved = fp32-equivalent algorithmic FLOPs of the timed launches / their CUDA-event time;
        # every fp32-equivalent product costs three fp16 tcgen05 MMAs (hi*hi + lo*hi + hi*lo).
        flops = siren_stats["flops"] / max(siren_stats["calls"], 1)
        avg_ms = sd["avg_ms"]
        achieved = flops / (avg_ms * 1e-3) / 1e12
        peak_tf = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1500.0
        roof = {"kernel": "siren_sdf_grad_kernel", "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": _ncu_traffic("prof_siren"),
                "peak_source": peak_src + " (dense bf16 cuBLAS, sustained; fp16 MMA runs at the same rate)",
                "algorithmic_flops_per_launch": flops, "avg_launch_ms": avg_ms,
                "rows_per_launch": siren_stats["rows"] / max(siren_stats["calls"], 1),
                "tensor_pipe_frac": 3 * achieved / peak_tf,
                "note": "fp32 accuracy from 3 fp16 MMAs per product: frac <= 1/3 by construction; "
                        "tensor_pipe_frac counts the issued MMA work; both this kernel and the cuBLAS reference of "
                        "`peak` run at the board's power limit, so tensor_pipe_frac is also the fraction of the energy "
                        "roofline (MAC/s per watt relative to a pure GEMM)",
                "limiter": "board power: sw_power_cap holds the SM clock below its maximum for the whole timed region "
                           "(see clocks); a kernel build with 8.6 % fewer cycles per tile gives the same step time at a "
                           "2 % lower clock (A/B on one box, profiles/r02a_siren_ab.txt) -- the ceiling is energy per "
                           "evaluation (3 x 672 fp16 MMAs per 128 rows), and cuBLAS bf16 itself sustains 62 % of nominal "
                           "under the same cap (MEASURED_PEAKS.json)",
                "ncu": _ncu("prof_siren")}

    line = {
        "metric": "iso-points/sec (project+resample)", "value": C2_POINTS * world / (ms * 1e-3), "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C2_WORKLOAD, "points_per_gpu": C2_POINTS, "sdf": args.sdf,
                   "sdf_eval": ("fused tcgen05 kernel, fp32-equivalent via fp16 hi/lo split" if sd else
                                "opaque nn.Module through autograd, fp32, TF32 off"),
                   "l2": "flushed between steps (256 MiB write)", "untimed_steps_before": max(args.warmup, PREROLL_STEPS),
                   "converged_frac_projection": converged_projection,
                   "valid_frac_after_resample": converged,
                   "points_after_filter": n_out,
                   "sdf_evaluations_per_point": siren_stats["rows"] / args.steps / C2_POINTS if sd else None,
                   "sdf_evaluations_per_s": (siren_stats["rows"] / args.steps / (sd["ms_per_step"] * 1e-3)) if sd else None,
                   "cpu_baseline_frnn": "O(n^2) single-thread brute force on the CPU sample (the reference's frnn_bf_cpu)",
                   "parallelism": "point-sharded x%d" % world},
        "e2e": {"value": C2_POINTS * world / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof,
        "roofline_frnn": frnn_roof,
        "own_kernels_ms_per_step": own_ms,
        "sdf_callback_and_glue_ms_per_step": ms - own_ms,
        "kernels": kern,
        "wall_s_timed_region": wall,
        "ms_each_step": [round(t, 3) for t in step_ms],
    }
    if world > 1:
        line["dist_parity"] = dist_parity(rank, world, dev, net, x_dev, out)
    if not args.no_c5 and not args.no_side:
        try:
            import bench_c5
            del flush
            torch.cuda.empty_cache()
            line["c5"] = bench_c5.run(rank, world, dev, steps=5, check=True)   # 2 steps were noisy (allocator growth)
        except Exception as e:      # never lose the headline line to the side record
            line["c5"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0:
        if world == 1 and not args.no_side:
            cpu_baseline(sample_points=max(500, args.cpu_sample // 10))       # warm-up (thread pool, allocator)
            line["cpu_baseline"] = cpu_baseline(sample_points=args.cpu_sample)
            line["c1"] = bench_c1(dev)
            import bench_frnn
            import bench_splat
            line["frnn"] = bench_frnn.run(args, dev, peaks, peak_src)
            line["splat"] = bench_splat.run(args, dev, peaks, peak_src)
            import bench_trace
            line["trace"] = bench_trace.run(dev, steps=3)
            try:
                import bench_pointops
                line["pointops"] = bench_pointops.run(dev)
            except Exception as e:      # never lose the headline line to a side record
                line["pointops"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            if not args.no_ref_cuda:
                line["ref_cuda"] = ref_cuda()
                rc = line["ref_cuda"]
                if rc.get("c2_ms"):
                    rc["c2_speedup_vs_reference_gpu"] = rc["c2_ms"] / ms
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def ref_cuda(timeout=420):
    """The reference itself on THIS GPU (its Python + its CUDA extensions recompiled for sm_100a): oracle/ref_gpu.py
    in a subprocess, after and outside every timed region of this process."""
    try:
        r = subprocess.run([sys.executable, "-m", "oracle.ref_gpu"], cwd=ROOT, capture_output=True, text=True,
                           timeout=timeout)
        lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        if not lines:
            return {"unavailable": "no output (rc %d): %s" % (r.returncode, r.stderr.strip()[-300:])}
        return json.loads(lines[-1])
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def dist_parity(rank, world, dev, net, x_dev, out):
    """N > 1, outside the timed region: the concatenation of the ranks' sharded results equals the single-GPU
    operator on the concatenated cloud (positions within 1e-4 rel, masks equal) -- tests/run_dist_gpu.py's check
    on the bench workload itself."""
    from isopoints_b200.dist import all_gather_varlen
    from isopoints_b200.levelset_sampling import UniformProjection
    pts, _ = all_gather_varlen(out["levelset_points"][0].contiguous())
    msk, _ = all_gather_varlen(out["mask"][0].float()[:, None].contiguous())
    xs, _ = all_gather_varlen(x_dev[0].contiguous())
    ref = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1).project_points(
        xs[None], net, skip_upsampling=True)
    same_shape = ref["levelset_points"].shape[1] == pts.shape[0]
    if same_shape:
        agree = ref["mask"][0] == msk[:, 0].bool()
        close = torch.isclose(ref["levelset_points"][0], pts, rtol=1e-4, atol=1e-5).all(-1)
        frac = float((agree & close).float().mean())
    else:
        frac = 0.0
    t = torch.tensor([frac], dtype=torch.float64, device=dev)
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"rows": int(pts.shape[0]), "rows_single_gpu": int(ref["levelset_points"].shape[1]),
            "agree_frac": float(t.item()), "ok": bool(float(t.item()) > 0.999),
            "what": "sharded project+resample (all ranks' rows concatenated) vs the single-GPU operator on the "
                    "concatenated %d-point cloud: mask equal and position within 1e-4 rel, per row" % xs.shape[0]}


def bench_c1(dev, reps=20):
    """BASELINE configs[0] ("C1"): 4 096 points, analytic unit-sphere SDF, 10 projection iterations --
    through the opaque nn.Module callback (11 SDF evaluations) and through the fully fused built-in
    sphere kernel (one launch, 37 B/point)."""
    from isopoints_b200.levelset_sampling import UniformProjection
    from tests.helpers import SphereSDF
    torch.manual_seed(0)
    x = ((torch.rand(1, 4096, 3) - 0.5) * 1.5).to(dev)
    proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5)
    out = {}
    for name, sdf in (("opaque_module", SphereSDF().to(dev)), ("fused_sphere_kernel", SphereSDF(analytic=True).to(dev))):
        f = lambda: proj.project_points(x, sdf, skip_resampling=True, skip_upsampling=True)
        for _ in range(3):
            r = f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            r = f()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out[name] = {"ms": ms, "points_per_s": 4096 / (ms * 1e-3), "converged": float(r["mask"].float().mean())}
    return out


# ---------------------------------------------------------------------------------------------
def cpu_baseline(sample_points=10_000, threads=None):
    """The oracle port of the same workload on the host cores, on a bounded sample."""
    from oracle import port
    from tests.helpers import pinned_siren
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1000)
    x = ((torch.rand(1, C2_POINTS, 3, generator=g) - 0.5) * 2)[0, :sample_points].contiguous()
    net = pinned_siren(0).as_opaque()
    frnn_fn, frnn_how = _cpu_frnn()
    t0 = time.perf_counter()
    p, n, v = port.project_points_packed(net, x, proj_max_iters=10, proj_tolerance=5e-5)
    t1 = time.perf_counter()
    if int(v.sum()) >= 18:
        port.resample(net, p[v], n[v], sample_iters=1, knn_k=8, proj_tolerance=5e-5, frnn_fn=frnn_fn)
    t2 = time.perf_counter()
    return {"value": sample_points / (t2 - t0), "unit": "points/s", "cores": threads, "kind": "port",
            "sample": "first %d of the 200000 C2 points, same SIREN; project %.1fs + resample %.1fs; "
                      "torch-CPU fp32 with %d threads (FRNN stage: %s)"
                      % (sample_points, t1 - t0, t2 - t1, threads, frnn_how)}


def _cpu_frnn():
    """FRNN stage of the CPU baseline: the reference's OWN CPU implementation (frnn._C.frnn_bf_cpu,
    external/FRNN/frnn/csrc/bruteforce/bruteforce_cpu.cpp, compiled unmodified into oracle/_ref) when that
    library is present, else the numpy restatement in oracle/port.py.  Same indices either way."""
    try:
        from oracle import ref_native
        probe = torch.rand(1, 32, 3)
        ref_native.frnn_bf_cpu(probe, probe, torch.tensor([32]), torch.tensor([32]), 4, 0.5)

        def fn(points, K, r):
            t = torch.as_tensor(points)[None].contiguous()
            n = torch.tensor([t.shape[1]])
            return ref_native.frnn_bf_cpu(t, t, n, n, K, r)[0][0].numpy()
        return fn, "the reference's own frnn_bf_cpu, O(n^2) brute force, 1 thread"
    except Exception:
        return None, "numpy brute force O(n^2), 1 thread"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    vals = []
    base = None
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(sample_points=max(1000, args.cpu_sample // 4))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        base = cpu_baseline(sample_points=args.cpu_sample)
        vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "iso-points/sec (project+resample)", "value": v, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": args.cpu_sample / v * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C2_WORKLOAD, "points_per_gpu": C2_POINTS,
                       "sample": "each step = the first %d of the 200000 C2 points (bounded CPU sample), same SIREN, "
                                 "torch-CPU fp32 autograd SDF" % args.cpu_sample},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=10_000)
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 side record (2 M points, 16 x 1024^2)")
    ap.add_argument("--no-side", action="store_true",
                    help="profiling runs: skip every side record (c5, cpu_baseline, c1, frnn, splat, trace, ref_cuda)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-on-GPU side record")
    ap.add_argument("--sdf", default="siren", choices=["siren", "opaque"],
                    help="siren: the reference's Siren decoder structure (fused SDF kernel); "
                         "opaque: same weights behind an opaque nn.Module (autograd SDF)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
        run_ours(args)


if __name__ == "__main__":
    main()
