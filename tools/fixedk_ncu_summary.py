#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture of ransac_fixedk_kernel (tools/stress_bench.py, F=20000) into the
JSON bench.py cites for `ransac_stress.hypothesis_kernel`: executed FP32 FLOP from the source page (ncu's op counters do
not see the packed FFMA2/FMUL2/FADD2), pipe utilisation, and the time split by kernel phase.

    python tools/fixedk_ncu_summary.py gpurun_out/r2_prof_fixedk_v2.ncu-rep gpurun_out/i_launches_stress.csv > profiles/r2_fixedk_ncu.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep, launches = sys.argv[1], sys.argv[2]
F = 20000


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
h, v = raw[0], raw[2]
get = lambda k: float(v[h.index(k)])
src = page("source")
hdr, data = src[1], src[2:]
isrc, iex, ithr, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
W = {"FFMA2": 4, "FMUL2": 2, "FADD2": 2, "FFMA": 2, "FMUL": 1, "FADD": 1}
flop = collections.Counter()
warp_instr = 0
for r in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0] if r[isrc].strip() else ""
    warp_instr += int(r[iex] or 0)
    if op in W:
        flop[op] += W[op] * int(r[ithr] or 0)
# phases: split at the CTA barriers (table build | main loop | pooled leftovers | finale)
tot = sum(int(r[isamp] or 0) for r in data)
segs, cum, prev = [], 0, 0
for k, r in enumerate(data):
    cum += int(r[isamp] or 0)
    if "BAR.SYNC" in r[isrc] or k == len(data) - 1:
        segs.append(round(100.0 * (cum - prev) / tot, 1)); prev = cum
ms = get("gpu__time_duration.sum")
ms = ms / 1e6 if ms > 1e3 else ms
total_flop = sum(flop.values())
times = collections.defaultdict(list)
for row in csv.reader(open(launches)):
    if len(row) > 5 and row[-1].replace(".", "").isdigit():
        for name in ("ransac_fixedk_kernel", "fixedk_prepare_kernel", "refit_kernel"):
            if name in row[4]:
                times[name].append(float(row[-1]) / 1e3)
med = {k: sorted(x)[len(x) // 2] for k, x in times.items()}
print(json.dumps({
    "capture": "ncu --set full --clock-control none, ransac_fixedk_kernel, F=20000 frames, K=4096, N=53, 40% outliers (tools/stress_bench.py, F=20000)",
    "kernel_ms": ms, "frames": F,
    "fp32_flop_executed_per_frame": total_flop / F, "fp32_flop_by_opcode": dict(flop),
    "executed_fp32_TFLOPs_in_capture": total_flop / (ms * 1e-3) / 1e12,
    "warp_instructions_per_frame": warp_instr / F,
    "sm__pipe_fma_cycles_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "sm__pipe_alu_cycles_active_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "smsp__issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "barrier_stall_per_issue": get("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    "registers": int(get("launch__registers_per_thread")),
    "warp_sample_share_by_phase_pct": dict(zip(["table_build", "main_loop", "pooled_leftovers", "finale"], segs)),
    "note": "FLOP = thread-level executed FFMA2 x4, FMUL2/FADD2 x2, FFMA x2, FMUL/FADD x1 from the ncu source page (ncu's op_ffma counters do not see the packed instructions)",
    "launch_us_F50000": med,
    "hypothesis_kernel_share_of_fit": (med.get("ransac_fixedk_kernel", 0) + med.get("fixedk_prepare_kernel", 0)) / max(1e-9, sum(med.values())),
    "share_source": "ncu launch list of tools/stress_bench.py at F=50000 (profiles/r2_launches_stress.csv): medians per kernel",
}, indent=1))
