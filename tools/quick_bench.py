"""Ad-hoc per-kernel timing (CUDA events) used while developing; bench.py is the contract bench."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import _native as N  # noqa: E402
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.engine import GeometryEngine, KeypointSet  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    eng = GeometryEngine("cuda:0")
    print("SMs", N.lib.egl_sm_count(), torch.cuda.get_device_name(0))
    F = int(os.environ.get("F", 1024))
    # K2: heatmaps larger than L2 (F*7.39MB)
    hm = torch.rand((F, 57, 135, 240), device="cuda") * 0.05
    clip = synthetic.make_clip(32, 1920, 1080, seed=0, ghost_prob=0.05)
    small = torch.from_numpy(clip["heatmaps"]).cuda()
    for i in range(0, F, 32):
        hm[i:i + 32] = small[: min(32, F - i)]
    kp = eng.alloc_keypoints(F)
    t = timeit(lambda: eng.decode(hm, 1920, 1080, out=kp))
    gb = F * 57 * 135 * 240 * 4 / 1e9
    print(f"decode   F={F}: {t:.3f} ms  {gb / t * 1e3:.0f} GB/s  {F / t * 1e3:.0f} frames/s")
    logits = torch.logit(hm.clamp(1e-6, 1 - 1e-6))
    t2 = timeit(lambda: eng.decode(logits, 1920, 1080, out=kp, from_logits=True))
    t3 = timeit(lambda: eng.decode(torch.sigmoid(logits), 1920, 1080, out=kp))
    print(f"decode from logits (fused sigmoid) F={F}: {t2:.3f} ms  {gb / t2 * 1e3:.0f} GB/s   vs torch.sigmoid + decode: {t3:.3f} ms")
    del logits
    eng.decode(hm, 1920, 1080, out=kp)
    # F1 + K3 + select + K4
    eng.synthesize(kp)
    fit = eng.alloc_fit(F)
    t = timeit(lambda: eng.fit(kp, out=fit))
    print(f"fit cv2  F={F}: {t:.3f} ms  {F / t * 1e3:.0f} frames/s")
    foot, count = synthetic.objects_to_arrays(clip["objects"], 23)
    foot = torch.from_numpy(np.tile(foot, (F // 32 + 1, 1, 1))[:F]).cuda(); count = torch.from_numpy(np.tile(count, F // 32 + 1)[:F]).cuda()
    pr = eng.alloc_projection(F, 23)
    def tail():
        h, a = eng.select(fit.status, 1)
        eng.project(fit.H, foot, count, 1920, 1080, h_index=h, out=pr)
    t = timeit(tail)
    print(f"select+project F={F}: {t:.3f} ms")
    def synth():
        eng.decode(hm[:64], 1920, 1080, out=None)
    kp2 = eng.decode(hm, 1920, 1080)
    t = timeit(lambda: eng.synthesize(kp2))
    print(f"synthesize F={F}: {t:.3f} ms (idempotent re-run)")
    # K1
    for (w, h) in [(1280, 720), (1920, 1080), (3840, 2160)]:
        Fp = 256 if w < 3840 else 96
        fr = torch.randint(0, 256, (Fp, h, w, 3), dtype=torch.uint8, device="cuda")
        out = torch.empty((Fp, 3, 540, 960), device="cuda")
        t = timeit(lambda: eng.preprocess(fr, out=out))
        alg = Fp * (h * w * 3 + 3 * 540 * 960 * 4) if w < 3840 else Fp * (h // 2 * w * 3 + 3 * 540 * 960 * 4)
        print(f"preprocess {w}x{h} F={Fp}: {t:.3f} ms  {alg / t / 1e6:.0f} GB/s  {Fp / t * 1e3:.0f} frames/s")
        del fr, out
    # fixed-K stress
    Fs, K = 2048, 4096
    xy, valid, flags, cams = synthetic.stress_point_sets(64, 1920, 1080, seed=1)
    xy = np.tile(xy, (Fs // 64, 1, 1))
    on = [i for i in range(57) if i not in (0, 1, 24, 25)]
    order = np.full((Fs, 64), 255, np.uint8); order[:, :53] = on
    kps = KeypointSet(torch.zeros((Fs, 57), dtype=torch.int32).cuda(), torch.zeros((Fs, 57)).cuda(), torch.from_numpy(xy).cuda(),
                      torch.from_numpy(order).cuda(), torch.from_numpy(np.full((Fs, 2), 53, np.int32)).cuda())
    fits = eng.alloc_fit(Fs)
    t = timeit(lambda: eng.fit(kps, mode=N.FIT_FIXED_K, K=K, seed=1, out=fits), iters=3, warm=1)
    flop = Fs * K * (627 + 21 * 53)
    print(f"fit fixedK F={Fs} K={K}: {t:.3f} ms  {flop / t / 1e9:.2f} TFLOP/s(alg)  {Fs / t * 1e3:.0f} frames/s  inliers {fits.info[:4,1].tolist()}")


if __name__ == "__main__":
    main()
