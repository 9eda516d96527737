#!/usr/bin/env python
"""K1 at 1080p on 2250 frames: the word-load + dp4a code (kPair, shipped) against the byte-wise code, both from the measurement
build of the library (tools/_variants/, EGL_PREPROCESS_PAIR=0 selects the byte-wise kernel), timed for >= 2 s each so that the
SM clock is the sustained one, outputs compared bit for bit.  One JSON line per variant.  With --once: a single launch of the
shipped library's kernel (for ncu)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "_variants", "libeagle_b200_variants.so")
F, H, W = 2250, 1080, 1920
BYTES = F * (H * W * 3 + 3 * 540 * 960 * 4)


def frames_on_device(torch):
    g = torch.Generator(device="cuda").manual_seed(1)
    return torch.randint(0, 256, (F, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)


def child() -> None:
    import hashlib
    import torch
    lib = C.CDLL(LIB)
    lib.egl_preprocess_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
    assert lib.egl_build_flags() == 1
    fr = frames_on_device(torch)
    out = torch.empty((F, 3, 540, 960), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        assert lib.egl_preprocess_u8(fr.data_ptr(), F, H, W, 3 * W, 3 * W * H, out.data_ptr(), st) == 0

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    burst = []
    for _ in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        burst.append(a.elapsed_time(b))
        time.sleep(0.3)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    n = 600
    a.record()
    for _ in range(n):
        run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    digest = hashlib.sha256(out[:64].cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"pair": int(os.environ.get("EGL_PREPROCESS_PAIR", 1)), "ms_sustained": round(ms, 4), "GBps_sustained": round(BYTES / ms / 1e6, 1),
                      "ms_burst_min": round(min(burst), 4), "GBps_burst": round(BYTES / min(burst) / 1e6, 1), "timed_s": round(ms * n / 1e3, 2),
                      "sha_first_64_frames": digest}))


def once() -> None:
    import torch
    from eagle_b200.engine import GeometryEngine
    eng = GeometryEngine("cuda:0")
    fr = frames_on_device(torch)
    out = torch.empty((F, 3, 540, 960), dtype=torch.float32, device="cuda")
    for _ in range(3):
        eng.preprocess(fr, out=out)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    elif "--once" in sys.argv:
        once()
    else:
        for pair in (1, 0, 1):
            r = subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, EGL_PREPROCESS_PAIR=str(pair)), capture_output=True, text=True, timeout=300)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"pair": pair, "error": r.stderr[-400:]}), flush=True)
