"""Development sweep over kernel launch variants (env switches read by the C ABI at first use)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from eagle_b200.engine import GeometryEngine
eng = GeometryEngine("cuda:0")
def timeit(fn, iters=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
what = os.environ["SWEEP_WHAT"]
if what == "decode":
    F = 2250
    hm = torch.rand((F, 57, 135, 240), device="cuda")
    kp = eng.alloc_keypoints(F)
    t = timeit(lambda: eng.decode(hm, 1920, 1080, out=kp))
    ok = bool((kp.flat.long() == hm.view(F, 57, -1).argmax(2)).all()) and bool((kp.score == hm.view(F, 57, -1).amax(2)).all())
    print("correct" if ok else "WRONG RESULT", end=" ")
    print(f"decode variant={os.environ.get('EGL_DECODE_VARIANT')} F={F}: {t:.3f} ms {F*57*135*240*4/t/1e6:.0f} GB/s")
else:
    for (w, h, F) in [(1280, 720, 1024), (1920, 1080, 1024), (3840, 2160, 256)]:
        fr = torch.randint(0, 256, (F, h, w, 3), dtype=torch.uint8, device="cuda")
        out = torch.empty((F, 3, 540, 960), device="cuda")
        t = timeit(lambda: eng.preprocess(fr, out=out))
        alg = F * ((h if w < 3840 else h // 2) * w * 3 + 3 * 540 * 960 * 4)
        print(f"preprocess rows={os.environ.get('EGL_PREPROCESS_ROWS')} {w}x{h} F={F}: {t:.3f} ms {alg/t/1e6:.0f} GB/s")
        del fr, out
''' % ROOT

for v in (os.environ.get("SWEEP_DECODE", "0 2 4 6 7 8 9 10 11").split()):
    env = dict(os.environ, SWEEP_WHAT="decode", EGL_DECODE_VARIANT=v)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
for r in (os.environ.get("SWEEP_PRE", "").split()):
    env = dict(os.environ, SWEEP_WHAT="pre", EGL_PREPROCESS_ROWS=r)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
