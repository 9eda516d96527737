#!/usr/bin/env python
"""A/B of egl_upload_frames' staging copy on one GPU box: memcpy / non-temporal stores / rep movsb, default vs
write-combined page-locked rings, slice sizes and thread counts.  Uses the measurement build of the library
(python -m eagle_b200.build --variants -> tools/_variants/), one subprocess per configuration (the ring is allocated once
per process).  Prints one JSON line per configuration; bytes on the device are checked against the host frames."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "_variants", "libeagle_b200_variants.so")
H, W, NF = 1080, 1920, 256


def child(threads: int) -> None:
    import numpy as np
    import torch
    lib = C.CDLL(LIB)
    lib.egl_upload_frames.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int]
    lib.egl_upload_frames.restype = C.c_int
    lib.egl_last_error.restype = C.c_char_p
    assert lib.egl_build_flags() == 1, "not the measurement build"
    base = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames = [np.bitwise_xor(base, np.uint8(i)) for i in range(NF)]
    ptrs = (C.c_void_p * NF)(*[f.ctypes.data for f in frames])
    dev = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
    nb = frames[0].nbytes

    def up(n):
        rc = lib.egl_upload_frames(ptrs, n, nb, dev.data_ptr(), threads)
        assert rc == 0, lib.egl_last_error()

    up(64)
    torch.cuda.synchronize()
    best = 0.0
    rates = []
    for _ in range(4):
        t0 = time.perf_counter()
        up(NF)
        dt = time.perf_counter() - t0
        rates.append(NF * nb / dt / 1e9)
    ok = all(np.array_equal(dev[i].cpu().numpy(), frames[i]) for i in (0, 1, 17, 100, NF - 1))
    print(json.dumps({"copy": int(os.environ.get("EGL_UPLOAD_COPY", 0)), "wc": int(os.environ.get("EGL_UPLOAD_WC", 0)),
                      "slice_kb": int(os.environ.get("EGL_UPLOAD_SLICE_KB", 4096)), "threads": threads,
                      "GBps": [round(r, 1) for r in rates], "median_GBps": round(sorted(rates)[len(rates) // 2], 1), "bytes_identical": ok}))


def main() -> None:
    configs = [(0, 0, 4096, 8), (1, 0, 4096, 8), (2, 0, 4096, 8), (1, 1, 4096, 8), (0, 0, 1024, 8), (1, 0, 1024, 8), (1, 0, 2048, 8),
               (1, 0, 4096, 6), (1, 0, 4096, 12), (0, 0, 4096, 12)]
    for copy, wc, kb, nt in configs:
        env = dict(os.environ, EGL_UPLOAD_COPY=str(copy), EGL_UPLOAD_WC=str(wc), EGL_UPLOAD_SLICE_KB=str(kb))
        try:
            r = subprocess.run([sys.executable, __file__, "--child", str(nt)], env=env, capture_output=True, text=True, timeout=120)
            out = r.stdout.strip().splitlines()
            print(out[-1] if out else json.dumps({"copy": copy, "wc": wc, "slice_kb": kb, "threads": nt, "error": r.stderr[-300:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"copy": copy, "wc": wc, "slice_kb": kb, "threads": nt, "error": "timeout"}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
