#!/usr/bin/env python
"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  Usage: summarize_ncu.py <round-tag>"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt, tot = {}, collections.Counter(), 0.0
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = re.sub(r"\(.*", "", r[ki])[:110]
        agg[name] = agg.get(name, 0) + v
        cnt[name] += 1
        tot += v
    lines = ["total %.1f us over %d launches (ncu: cold cache, serialised -- compare SHARES)" % (tot, sum(cnt.values())),
             "%12s %7s %6s  kernel" % ("us", "share", "n")]
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:30]:
        lines.append("%12.1f %6.1f%% %6d  %s" % (v, 100 * v / tot, cnt[k], k))
    own = sum(v for k, v in agg.items() if "isob200" in k)
    lines.append("own kernels (isob200::*): %.1f us = %.2f%% of the listed time" % (own, 100 * own / tot))
    return "\n".join(lines)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        res.append("kernel: " + d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                res.append("  %-72s %s %s" % (k, d[k], u.get(k, "")))
        for k in hdr:   # tensor-pipe metrics (only non-zero for the tcgen05 kernel)
            if (("pipe_tensor_cycles_active" in k or "utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak" in k
                 or k == "sm__inst_executed.avg.per_cycle_active") and d.get(k, "0") not in ("0", "")):
                res.append("  %-72s %s %s" % (k[-72:], d[k], u.get(k, "")))
    return "\n".join(res)


def traffic(rep):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, in bytes (first captured launch)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    d, u = dict(zip(hdr, rows[2])), dict(zip(hdr, units))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k].replace(",", "")) * mult.get(u[k], 1)
    name = re.sub(r"[<(].*", "", d["Kernel Name"].replace("void ", "").replace("isob200::", "")).strip()
    extra = {}
    for k, short in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
                     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
                     ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
                      "tensor_ops_pct_of_peak"),
                     ("smsp__inst_executed.sum", "warp_inst"),
                     ("gpu__time_duration.sum", "duration")):
        try:
            extra[short] = float(d[k].replace(",", ""))
        except Exception:
            pass
    try:    # the launch duration in one unit (ncu picks ns / us / ms per report)
        tm = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
        extra["duration_us"] = float(d["gpu__time_duration.sum"].replace(",", "")) * tm.get(u["gpu__time_duration.sum"], 1.0)
    except Exception:
        pass
    return name, tot, extra


def hot_lines(rep, top=14):
    """SASS-level hot spots: share of warp-stall samples and of executed instructions."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return "(source page unavailable)"
    hdr = rows[hi[0]]
    si, ni, ii = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for n, r in enumerate(rows[hi[0] + 1:]):
        try:
            data.append((float(r[ni]), float(r[ii]), n, r[si].strip()))
        except Exception:
            pass
    ts = sum(d[0] for d in data) or 1
    ti = sum(d[1] for d in data) or 1
    lines = ["%d SASS instructions, %d stall samples, %.0f warp-instructions executed" % (len(data), ts, ti),
             "samples%  inst%   #   SASS"]
    for v, e, n, src in sorted(data, reverse=True)[:top]:
        lines.append("%6.1f%% %5.1f%% %4d  %s" % (100 * v / ts, 100 * e / ti, n, src[:100]))
    return "\n".join(lines)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    os.makedirs(PROF, exist_ok=True)
    for name in ("launches_c2", "launches_c4", "launches_c3"):
        p = os.path.join(OUT, name + ".csv")
        if os.path.exists(p):
            open(os.path.join(PROF, "%s_%s.txt" % (tag, name)), "w").write(launches(p) + "\n")
    import json
    tpath = os.path.join(PROF, "ncu_traffic.json")
    tr = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".ncu-rep"):
            rep = os.path.join(OUT, f)
            try:
                name, b, extra = traffic(rep)
                tr[f[:-8]] = {"dram_bytes_per_launch": b, "kernel": name, "capture": "%s_%s" % (tag, f[:-8]), **extra}
            except Exception as e:
                print("traffic:", f, e)
            txt = raw(rep) + "\n\nhottest source lines (share of executed instructions):\n" + hot_lines(rep) + "\n"
            open(os.path.join(PROF, "%s_%s.txt" % (tag, f[:-8])), "w").write(txt)
    json.dump(tr, open(tpath, "w"), indent=1, sort_keys=True)
    for f in ("bench.json", "bench_reference.json", "bench_splat.json", "bench_trace.json", "bench_n2.json",
              "bench_n8.json", "bench_frnn.json", "bench_pointops.json", "bench_rays.json", "ref_cuda_timing.json",
              "siren_timeline.txt", "siren_cta_sweep.txt", "siren_ab.txt", "pytest_gpu.log"):
        p = os.path.join(OUT, f)
        if os.path.exists(p) and not f.endswith(".json"):      # plain-text evidence: verbatim
            open(os.path.join(PROF, "%s_%s" % (tag, f.replace(".log", ".txt"))), "w").write(open(p).read())
            continue
        if os.path.exists(p):
            text = open(p).read()
            try:                                   # a pretty-printed file is one JSON document: keep it whole
                json.loads(text)
                keep = text.strip()
            except ValueError:                     # a log with one JSON line at the end
                lines = [l for l in text.splitlines() if l.startswith("{")]
                keep = lines[-1] if lines else None
            if keep:
                open(os.path.join(PROF, "%s_%s" % (tag, f)), "w").write(keep + "\n")


if __name__ == "__main__":
    main()
