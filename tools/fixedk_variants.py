#!/usr/bin/env python
"""ransac_fixedk_kernel variants of the measurement build (EGL_FIXEDK_VARIANT: CTAs per SM asked of the register allocator x
scoring-loop unroll x static / dynamic batch distribution) on the stress shape (F = 20000, K = 4096, N = 53): kernel time and
FMA-pipe utilisation from ncu, and a digest of every output so that the variants can be compared bit for bit.

    python tools/fixedk_variants.py            # parent: one ncu run per variant, one JSON line each
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {0: "6 CTAs, unroll 4, static, compare + predicated add (shipped)", 10: "6, 4, static, sign-bit count",
            11: "6, 8, static, sign-bit count", 12: "6, 4, dynamic, sign-bit count"}
# measured before the sign-bit count existed (profiles/r2_fixedk_variants_occupancy.txt): 7 and 8 CTAs per SM, unroll 2 / 8 and the
# dynamic batch distribution were all 1.4 - 11.7 % slower than 6 CTAs / unroll 4 / static


def child() -> None:
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from eagle_b200 import _native as N
    from eagle_b200 import synthetic
    from eagle_b200.engine import GeometryEngine, KeypointSet
    assert N.lib.egl_build_flags() == 1, "needs EAGLE_B200_LIBRARY = the measurement build"
    F, K = 20000, 4096
    eng = GeometryEngine("cuda:0")
    xy, valid, flags, cams = synthetic.stress_point_sets(256, 1920, 1080, seed=1)
    reps = (F + 255) // 256
    xy = np.tile(xy, (reps, 1, 1))[:F]
    on = [i for i in range(57) if i not in (0, 1, 24, 25)]
    order = np.full((F, 64), 255, np.uint8); order[:, :53] = on
    kp = KeypointSet(torch.zeros((F, 57), dtype=torch.int32).cuda(), torch.zeros((F, 57)).cuda(), torch.from_numpy(xy).cuda(),
                     torch.from_numpy(order).cuda(), torch.from_numpy(np.full((F, 2), 53, np.int32)).cuda())
    fit = eng.alloc_fit(F)
    for _ in range(3):
        eng.fit(kp, mode=N.FIT_FIXED_K, K=K, seed=1, out=fit)
    torch.cuda.synchronize()
    h = hashlib.sha256()
    for t in (fit.status, fit.inlier_mask, fit.used_mask, fit.H, fit.info):
        h.update(t.cpu().numpy().tobytes())
    print("DIGEST", h.hexdigest()[:16])


def main() -> None:
    lib = os.path.join(ROOT, "tools", "_variants", "libeagle_b200_variants.so")
    first = None
    for v, what in VARIANTS.items():
        env = dict(os.environ, EAGLE_B200_LIBRARY=lib, EGL_FIXEDK_VARIANT=str(v))
        log = os.path.join(ROOT, "gpurun_out", f"fixedk_variant_{v}.csv")
        cmd = ["ncu", "--metrics", "gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,"
               "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,"
               "launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active",
               "--clock-control", "none", "-k", "regex:ransac_fixedk", "-s", "1", "-c", "2", "--csv", "--log-file", log,
               sys.executable, __file__, "--child"]
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        except subprocess.TimeoutExpired:
            print(json.dumps({"variant": v, "what": what, "error": "timeout"}), flush=True)
            continue
        digest = [ln.split()[1] for ln in r.stdout.splitlines() if ln.startswith("DIGEST")]
        row = {"variant": v, "what": what, "digest": digest[0] if digest else None}
        try:
            import csv
            vals = {}
            for rec in csv.reader(open(log)):
                if len(rec) > 10 and rec[0].isdigit():
                    vals.setdefault(rec[-3], []).append(float(rec[-1].replace(",", "")))
            for k, xs in vals.items():
                row[k.split(".")[0]] = round(sum(xs) / len(xs), 3)
        except Exception as exc:  # noqa: BLE001
            row["error"] = str(exc) + r.stderr[-200:]
        first = first or row.get("digest")
        row["same_outputs_as_variant_0"] = row.get("digest") == first
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
