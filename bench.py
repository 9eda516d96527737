#!/usr/bin/env python
"""Benchmark of the per-frame geometry path (decode -> RANSAC homography -> projection, with the
uint8 -> network-input preprocessing in front) on synthetic data of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one clip: 2250 frames of 1920x1080 (BASELINE.json
configs[1]: a 90 s clip at 25 fps) per GPU.  N > 1 (launched by torch.distributed.run, one rank per
GPU) shards a clip of N x 2250 frames into contiguous frame ranges, one per rank -- frames are
independent, there is no data-path collective -- and gathers the per-frame results to rank 0 over
NCCL inside the timed region ("weak" scaling: per-GPU work fixed).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = frames/s
through the same C-ABI calls starting from pinned HOST buffers (H2D of frames, heatmaps and foot
points and D2H of every result inside the timed region); `roofline` describes the dominant kernel
(the heatmap arg-max, HBM-bound); `cpu_baseline` is the oracle port of the reference's cv2/numpy
path timed on this box's host cores.  `--impl reference` times that CPU path alone, on all cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_OUT = sys.stdout

W, H = 1920, 1080
FRAMES_PER_GPU = 2250
POOL = 64            # distinct synthetic frames (landmark layouts / boxes); tiled to the clip length
MAX_OBJ = 23         # 22 players + ball
HM_BYTES = 57 * 135 * 240 * 4  # algorithmic bytes per frame of the dominant kernel
WORKLOAD = "1080p 90 s clip @25 fps (2250 frames/GPU): preprocess + heatmap decode + synthesis + RANSAC homography + projection"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's cv2/numpy path (oracle/pipeline.py + preprocess.py)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n, with_pre = args
    import cv2
    cv2.setNumThreads(1)
    import torch
    torch.set_num_threads(1)
    from eagle_b200 import synthetic
    from oracle import pipeline, preprocess
    pool = synthetic.make_clip(min(n, 16), W, H, seed=seed, with_frames=with_pre, ghost_prob=0.05)
    reps = (n + len(pool["objects"]) - 1) // len(pool["objects"])
    idx = (list(range(len(pool["objects"]))) * reps)[:n]
    hm = [pool["heatmaps"][i] for i in idx]
    ob = [pool["objects"][i] for i in idx]
    t0 = time.perf_counter()
    if with_pre:
        for i in idx:
            preprocess.preprocess_reference_calls(pool["frames"][i])
    pipeline.get_coordinates(hm, ob, W, H)
    return time.perf_counter() - t0


def cpu_path_fps(n_frames: int, workers: int, with_pre: bool = True):
    """frames/s of the CPU path over n_frames (split contiguously over `workers` processes)."""
    if workers <= 1:
        dt = _cpu_worker((1, n_frames, with_pre))
        return n_frames / dt, dt
    import multiprocessing as mp
    per = [n_frames * (r + 1) // workers - n_frames * r // workers for r in range(workers)]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(r + 1, per[r], with_pre) for r in range(workers) if per[r] > 0])
        dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    sample = 16 * cores
    for _ in range(max(0, args.warmup - 2)):  # data generation dominates a warm-up; one is plenty
        cpu_path_fps(cores, cores)
    vals = []
    for _ in range(args.steps):
        # the workers generate their own inputs before their timers start; wall time below includes it,
        # so time inside the workers instead: use the max of per-worker compute times
        import multiprocessing as mp
        per = [sample // cores] * cores
        with mp.get_context("fork").Pool(cores) as pool:
            times = pool.map(_cpu_worker, [(r + 1, per[r], True) for r in range(cores)])
        vals.append(sample / max(times))
    v = statistics.median(vals)
    line = {"impl": "reference", "metric": "frames/sec, decode->RANSAC homography->projection", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (cv2, numpy)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "sample_frames_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} frames/step over {cores} processes (oracle port of coordinate_model.py:221-415: "
                                       "cv2.resize+normalise, np.argmax decode, cv2.findHomography cascade, cv2.perspectiveTransform)"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_inputs(torch, dev, n_frames, seed):
    """Synthetic clip in HBM: uint8 frames, heatmaps (distinct noise per frame, POOL landmark layouts), foot points."""
    from eagle_b200 import synthetic
    pool = synthetic.make_clip(POOL, W, H, seed=seed, ghost_prob=0.05)
    foot_p, count_p = synthetic.objects_to_arrays(pool["objects"], MAX_OBJ)
    g = torch.Generator(device=dev); g.manual_seed(seed)
    bumps = torch.from_numpy(pool["heatmaps"]).to(dev)
    bumps = torch.clamp(bumps - 0.05, min=0)  # strip the pool's own background; re-noised per frame below
    hm = torch.empty((n_frames, 57, 135, 240), dtype=torch.float32, device=dev)
    for s in range(0, n_frames, POOL):
        n = min(POOL, n_frames - s)
        hm[s:s + n] = torch.rand((n, 57, 135, 240), generator=g, device=dev) * 0.05
        hm[s:s + n] = torch.maximum(hm[s:s + n], bumps[:n])
    frames = torch.empty((n_frames, H, W, 3), dtype=torch.uint8, device=dev)
    for s in range(0, n_frames, 50):
        n = min(50, n_frames - s)
        frames[s:s + n] = torch.randint(0, 256, (n, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    reps = (n_frames + POOL - 1) // POOL
    foot = torch.from_numpy(np.tile(foot_p, (reps, 1, 1))[:n_frames].copy()).to(dev)
    count = torch.from_numpy(np.tile(count_p, reps)[:n_frames].copy()).to(dev)
    return frames, hm, foot, count, pool


def main():
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 from C (NCCL prints its version
    # banner there) are redirected to stderr, and the line is printed through a private copy of fd 1.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-propagated", action="store_true", help="skip the sparse-keypoint-cadence measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from eagle_b200 import _native as N
    from eagle_b200.engine import GeometryEngine
    from eagle_b200.sharding import frame_range, gather_to_rank0, pack_results

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    lo, hi = frame_range(F * world, rank, world)
    assert hi - lo == F
    eng = GeometryEngine(dev)
    frames, hm, foot, count, pool = build_inputs(torch, dev, F, seed=1000 + rank)
    x = torch.empty((F, 3, N.MODEL_H, N.MODEL_W), dtype=torch.float32, device=dev)
    kp = eng.alloc_keypoints(F); fit = eng.alloc_fit(F); proj = eng.alloc_projection(F, MAX_OBJ)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches = {"n": 0}

    def step(i=None):
        """The hot path over this rank's F frames, inputs resident in HBM.  8 kernel launches, one stream.

        K1's output feeds the keypoint network, which is not part of this path, so nothing in the step
        depends on it; it is simply run last."""
        if i is not None: ev[i][0].record()
        eng.decode(hm, W, H, 0.3, out=kp)                                # K2  (2 launches)
        if i is not None: ev[i][1].record()
        eng.synthesize(kp)                                               # F1  (1)
        eng.fit(kp, out=fit)                                             # K3  (2)
        h_index, attempted = eng.select(fit.status, 1)                   # cadence (1)
        eng.project(fit.H, foot, count, W, H, h_index=h_index, out=proj)  # K4  (1)
        if i is not None: ev[i][2].record()
        eng.preprocess(frames, out=x)                                    # K1  (1 launch)
        if i is not None: ev[i][3].record()
        launches["n"] += 8
        if world > 1:
            rec = pack_results([fit.H, fit.inlier_mask, fit.status, proj.coords, proj.in_bounds, proj.bounds])
            gather_to_rank0(rec, [F] * world)
        return h_index, attempted

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches["n"] = 0
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step(i)
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    gpu_launches = launches["n"]
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = F * world / (ms_step * 1e-3)

    # per-kernel split of the step (same events, same stream) -- explains `value`
    pre_ms = statistics.mean(ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps))   # decode (argmax+postprocess), alone
    tail_ms = statistics.mean(ev[i][1].elapsed_time(ev[i][2]) for i in range(args.steps))  # synth+fit+select+project
    k1_ms = statistics.mean(ev[i][2].elapsed_time(ev[i][3]) for i in range(args.steps))    # preprocess
    peak, peak_src = peaks()
    achieved = F * HM_BYTES / (pre_ms * 1e-3) / 1e9
    roofline = {"kernel": "egl::argmax_ldg_kernel (K2 heatmap decode; interval also holds postprocess_kernel, <1%)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": HM_BYTES, "launch_ms": pre_ms, "traffic": None}
    tr = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_at_F2250")
        except Exception:
            pass

    # ---- e2e: pinned host buffers -> H2D -> same kernels -> D2H of every result, chunked + double-buffered
    e2e = None
    if rank == 0 or world > 1:
        CH = 125
        nh = 2 * CH
        h_frames = torch.empty((nh, H, W, 3), dtype=torch.uint8).pin_memory(); h_frames.copy_(frames[:nh].cpu())
        h_hm = torch.empty((nh, 57, 135, 240), dtype=torch.float32).pin_memory(); h_hm.copy_(hm[:nh].cpu())
        h_foot = torch.empty((nh, MAX_OBJ, 2), dtype=torch.float32).pin_memory(); h_foot.copy_(foot[:nh].cpu())
        h_cnt = torch.empty((nh,), dtype=torch.int32).pin_memory(); h_cnt.copy_(count[:nh].cpu())
        copy_s = torch.cuda.Stream(dev); comp_s = torch.cuda.Stream(dev)
        bufs = []
        for b in range(2):
            bufs.append(dict(fr=torch.empty((CH, H, W, 3), dtype=torch.uint8, device=dev), hm=torch.empty((CH, 57, 135, 240), device=dev),
                             foot=torch.empty((CH, MAX_OBJ, 2), device=dev), cnt=torch.empty((CH,), dtype=torch.int32, device=dev),
                             x=torch.empty((CH, 3, 540, 960), device=dev), kp=eng.alloc_keypoints(CH), fit=eng.alloc_fit(CH),
                             proj=eng.alloc_projection(CH, MAX_OBJ), ready=torch.cuda.Event(), done=torch.cuda.Event()))
        rec_like = None
        nchunks = F // CH
        h_out = None

        def e2e_step(copy_heatmaps=True):
            nonlocal h_out, rec_like
            outs = []
            for c in range(nchunks):
                b = bufs[c & 1]; s = (c & 1) * CH
                with torch.cuda.stream(copy_s):
                    copy_s.wait_event(b["done"])          # previous use of this buffer finished
                    b["fr"].copy_(h_frames[s:s + CH], non_blocking=True)
                    if copy_heatmaps:
                        b["hm"].copy_(h_hm[s:s + CH], non_blocking=True)
                    b["foot"].copy_(h_foot[s:s + CH], non_blocking=True); b["cnt"].copy_(h_cnt[s:s + CH], non_blocking=True)
                    b["ready"].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(b["ready"])
                    eng.preprocess(b["fr"], out=b["x"])
                    eng.decode(b["hm"], W, H, 0.3, out=b["kp"]); eng.synthesize(b["kp"]); eng.fit(b["kp"], out=b["fit"])
                    hi_, at_ = eng.select(b["fit"].status, 1)
                    eng.project(b["fit"].H, b["foot"], b["cnt"], W, H, h_index=hi_, out=b["proj"])
                    rec = pack_results([b["kp"].xy, b["kp"].order, b["kp"].count, b["fit"].H, b["fit"].inlier_mask, b["fit"].status,
                                        hi_, b["proj"].coords, b["proj"].coords_i, b["proj"].in_bounds, b["proj"].bounds])
                    if h_out is None:
                        h_out = torch.empty((nchunks, CH, rec.shape[1]), dtype=torch.uint8).pin_memory()
                    h_out[c].copy_(rec, non_blocking=True)
                    b["done"].record(comp_s)
            comp_s.synchronize()
            return h_out

        for _ in range(2):
            e2e_step()
        barrier()
        e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e_dt = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([e_dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_dt = float(t.item())
        # variant: heatmaps stay on the device (where the keypoint network leaves them); only frames and boxes cross PCIe
        for b in bufs:
            b["hm"].copy_(hm[:CH])
        e2e_step(False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step(False)
        torch.cuda.synchronize()
        e_dt2 = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([e_dt2], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_dt2 = float(t.item())
        h2d = nchunks * CH * (H * W * 3 + HM_BYTES + MAX_OBJ * 8 + 4)
        d2h = int(h_out.numel())
        e2e = {"value": nchunks * CH * world / e_dt, "unit": "frames/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": e_dt * 1e3,
               "heatmaps_on_device": {"value": nchunks * CH * world / e_dt2, "unit": "frames/s",
                                      "h2d_bytes_per_step": nchunks * CH * (H * W * 3 + MAX_OBJ * 8 + 4) * world,
                                      "note": "same, but the heatmaps are already in HBM (they are the keypoint network's output); frames + boxes from host"},
               "note": "pinned host -> H2D (frames u8 + heatmaps f32 + foot points) -> 8 kernels/chunk -> D2H of all "
                                                  "per-frame results; 125-frame chunks, double-buffered on two streams; PCIe-bound"}

    # ---- the same path at the reference's default cadence (main.py:27: network every 8th frame, Lucas-Kanade
    # propagation in between, homography once a second): frames + head heatmaps resident in HBM, one clip of F
    # frames per GPU.  Reported beside the headline, not as it: BASELINE.json's metric is the per-frame stage.
    propagated = None
    if H == 1080 and W == 1920 and not args.no_propagated:
        from eagle_b200 import synthetic
        from eagle_b200.propagation import PropagatedPath
        fps_ref = 25
        kint, hint = max(1, int(fps_ref / 3)), max(1, int(fps_ref / 1))
        del x
        torch.cuda.empty_cache()
        pf, heads = synthetic.tiled_flow_clip_device(F, kint, dev, paths=3, seed=rank, out_frames=frames)
        prop = PropagatedPath(eng)

        def prop_once():
            o = prop.run(pf, heads, None, kint, hint, False)
            eng.project(o["H"], foot, count, W, H, h_index=o["h_index"], out=proj)
            return o

        for _ in range(2):
            o = prop_once()
        barrier()
        p_steps = max(2, min(args.steps, 5))
        pe0 = torch.cuda.Event(enable_timing=True); pe1 = torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(p_steps):
            o = prop_once()
        pe1.record()
        barrier()
        p_ms = pe0.elapsed_time(pe1) / p_steps
        ga = torch.cuda.Event(enable_timing=True); gb = torch.cuda.Event(enable_timing=True)
        ga.record(); eng.gray_pyramid(pf, 2, out=prop.pyr); gb.record(); torch.cuda.synchronize()
        g_ms = ga.elapsed_time(gb)
        if world > 1:
            t = torch.tensor([p_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p_ms = float(t.item())
        pyr_bytes = F * (H * W * 3 + int(N.lib.egl_pyramid_bytes(H, W, 2)))
        propagated = {"value": F * world / (p_ms * 1e-3), "unit": "frames/s", "ms_per_clip": p_ms, "frames_per_gpu": F,
                      "workload": f"1080p clip, keypoint interval {kint} (network heads decoded, {kint - 1} of {kint} frames carried by "
                                  f"pyramidal Lucas-Kanade + the reference's filters), homography interval {hint}, rendered pitch frames",
                      "mean_keypoints_per_frame": float(o["count"][:, 0].float().mean().item()),
                      "frames_with_homography": int((o["h_index"] >= 0).sum().item()), "stats": dict(prop.stats),
                      "gray_pyramid": {"ms": g_ms, "bound": "hbm", "achieved": pyr_bytes / (g_ms * 1e-3) / 1e9, "unit": "GB/s",
                                       "frac": pyr_bytes / (g_ms * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_frame": H * W * 3 + int(N.lib.egl_pyramid_bytes(H, W, 2))},
                      "parity": "bit-exact with cv2.calcOpticalFlowPyrLK / the reference's dict (tests/test_gpu_flow.py)"}
        # end to end for this cadence: frames (uint8), the heads' heatmaps and the foot points come from pinned host
        # memory every clip, every per-frame result goes back; only 1 frame in 8 brings a heatmap along
        BL = 16 * kint
        nbl = (F + BL - 1) // BL
        st_fr = torch.empty((2, BL, H, W, 3), dtype=torch.uint8).pin_memory(); st_fr.copy_(pf[:2 * BL].view(2, BL, H, W, 3).cpu())
        st_hm = torch.empty((2, BL // kint, 57, 135, 240), dtype=torch.float32).pin_memory()
        st_hm.copy_(heads[:2 * (BL // kint)].view(2, BL // kint, 57, 135, 240).cpu())
        st_foot = torch.empty(foot.shape, dtype=foot.dtype).pin_memory(); st_foot.copy_(foot.cpu())
        st_cnt = torch.empty(count.shape, dtype=count.dtype).pin_memory(); st_cnt.copy_(count.cpu())
        h_res = None

        def prop_e2e():
            nonlocal h_res
            for b in range(nbl):
                n = min(BL, F - b * BL)
                pf[b * BL:b * BL + n].copy_(st_fr[b & 1, :n], non_blocking=True)
                nh_ = (n + kint - 1) // kint
                heads[b * (BL // kint):b * (BL // kint) + nh_].copy_(st_hm[b & 1, :nh_], non_blocking=True)
            foot.copy_(st_foot, non_blocking=True); count.copy_(st_cnt, non_blocking=True)
            o = prop.run(pf, heads, None, kint, hint, False)
            eng.project(o["H"], foot, count, W, H, h_index=o["h_index"], out=proj)
            rec = pack_results([o["xy"], o["order"], o["count"], o["src"], o["H"], o["fit_ok"], o["h_index"], proj.coords, proj.coords_i,
                                proj.in_bounds, proj.bounds])
            if h_res is None:
                h_res = torch.empty(rec.shape, dtype=torch.uint8).pin_memory()
            h_res.copy_(rec, non_blocking=True)
            torch.cuda.synchronize()

        prop_e2e()
        barrier()
        t_e = time.perf_counter()
        for _ in range(2):
            prop_e2e()
        pe_dt = (time.perf_counter() - t_e) / 2
        if world > 1:
            t = torch.tensor([pe_dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pe_dt = float(t.item())
        propagated["e2e"] = {"value": F * world / pe_dt, "unit": "frames/s", "ms_per_clip": pe_dt * 1e3,
                             "h2d_bytes_per_step": world * (F * H * W * 3 + ((F + kint - 1) // kint) * HM_BYTES + F * (MAX_OBJ * 8 + 4)),
                             "d2h_bytes_per_step": world * int(h_res.numel()),
                             "note": "pinned host -> H2D of all frames, the chain heads' heatmaps and the foot points -> PropagatedPath + "
                                     "projection -> D2H of every per-frame result; PCIe-bound"}
        if rank == 0 and not args.no_cpu_baseline:
            # the reference's CPU statements for the same cadence (cv2.calcOpticalFlowPyrLK, cv2.cvtColor, cv2.fitLine,
            # cv2.findHomography, cv2.perspectiveTransform), one thread, on the first chains of the same clip
            from oracle import pipeline as _pipe
            n_cpu = min(F, 50 * kint)
            fr_h = pf[:n_cpu].cpu().numpy()
            hm_h = {i: heads[i // kint].cpu().numpy() for i in range(0, n_cpu, kint)}
            objs_h = [pool["objects"][i % len(pool["objects"])] for i in range(n_cpu)]
            t_c = time.perf_counter()
            _pipe.get_coordinates_propagated(list(fr_h), hm_h, objs_h, fps_ref, 1, 3, library_calls=True)
            dt_c = time.perf_counter() - t_c
            propagated["cpu_baseline"] = {"value": n_cpu / dt_c, "unit": "frames/s", "cores": 1, "kind": "port",
                                          "sample": f"first {n_cpu} frames of the same clip, one thread: the reference's cv2/numpy calls for this "
                                                    "cadence (oracle/pipeline.py, library_calls=True); K1 and the network not included on either side"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline:
        n_cpu = 2048  # ~15 s of single-thread CPU work
        v, dt = cpu_path_fps(n_cpu, 1)
        cpu = {"value": v, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"{n_cpu} frames of the same 1080p workload, 1 process / 1 thread: oracle port of the reference path "
                         "(cv2.resize+normalise, np.argmax decode, synthesis, cv2.findHomography cascade, cv2.perspectiveTransform)"}

    line = {"metric": "frames/sec, decode->RANSAC homography->projection", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 heatmaps / u8 frames / f64 refit", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "frames_per_gpu": F, "heatmaps": [57, 135, 240], "objects_per_frame": MAX_OBJ,
                       "fit": "cv2-compatible adaptive RANSAC (cap 2000) + LS refit + LM", "l2_policy": "inputs larger than L2 "
                       f"({F * (H * W * 3 + HM_BYTES) / 1e9:.1f} GB streamed per step vs 126 MB L2)", "parallelism": f"frame-range x{world}"},
            "kernel_ms": {"decode_K2": pre_ms, "synth_fit_select_project": tail_ms, "preprocess_K1": k1_ms},
            "stage_fps_without_preprocess": F * world / ((pre_ms + tail_ms) * 1e-3),
            "preprocess_roofline": {"bound": "hbm", "achieved": F * (H * W * 3 + 3 * 540 * 960 * 4) / (k1_ms * 1e-3) / 1e9, "unit": "GB/s",
                                    "frac": F * (H * W * 3 + 3 * 540 * 960 * 4) / (k1_ms * 1e-3) / 1e9 / peak},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "propagated_cadence": propagated}
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
