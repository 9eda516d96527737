#!/usr/bin/env python
"""Benchmark of the per-frame geometry path (decode -> RANSAC homography -> projection, with the
uint8 -> network-input preprocessing in front) on synthetic data of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one clip: 2250 frames of 1920x1080 (BASELINE.json
configs[1]: a 90 s clip at 25 fps) per GPU.  N > 1 (launched by torch.distributed.run, one rank per
GPU) shards a clip of N x 2250 frames into contiguous frame ranges, one per rank -- frames are
independent, there is no data-path collective -- and gathers the per-frame results to rank 0 over
NCCL inside the timed region ("weak" scaling: per-GPU work fixed).

Prints ONE JSON line (rank 0):
  value        frames/s over exactly K steps, inputs resident in HBM (CUDA events, max over ranks); the small kernels after
               the decode run on a side stream next to K1, which nothing in the step depends on
  sustained    the same step repeated for >= 2 s (clocks sampled throughout); kernel split and rooflines come from here
  roofline     the kernel with the largest share of the step (K1 preprocess, HBM-bound); roofline_decode = K2
  e2e          frames/s through the PUBLIC API: CoordinateModel.get_coordinates(list of host frames) -> reference-format
               dict, with a device-resident stand-in for the keypoint network; H2D of every frame, D2H of every result
               and the dict assembly are inside the timed region (api_e2e has the split)
  full_match   BASELINE configs[2]: a 135 000-frame clip frame-sharded over the N ranks through run_sharded
               (chunks streamed through resident buffers, NCCL gather, clip-wide cadence + projection on rank 0)
  ransac_stress, sweep_4k   BASELINE configs[3] and configs[4] (N = 1 only)
  cpu_baseline the reference's CPU path on this box's host cores (the unmodified reference when /root/reference is
               mounted, else the oracle port, which is pinned JSON-identical to it)
`--impl reference` times that CPU path alone, on all cores.
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
MATCH_FRAMES = 135000   # BASELINE configs[2]: 90 min at 25 fps
POOL = 64               # distinct synthetic frames (landmark layouts / boxes); tiled to the clip length
MAX_OBJ = 23            # 22 players + ball
HM_BYTES = 57 * 135 * 240 * 4                       # algorithmic bytes per frame of K2
K1_BYTES = H * W * 3 + 3 * 540 * 960 * 4            # algorithmic bytes per frame of K1 at 1080p: uint8 in + float32 out
WORKLOAD = "1080p 90 s clip @25 fps (2250 frames/GPU): preprocess + heatmap decode + synthesis + RANSAC homography + projection"
MIN_TIMED_S = 2.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return None


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
# CPU arm: the reference's cv2/numpy path.  The unmodified reference (oracle/ref_harness.py) when /root/reference is
# mounted -- it never is on the GPU box -- else the oracle port (oracle/pipeline.py + preprocess.py), which the CPU
# tests hold JSON-identical to it.
# ------------------------------------------------------------------------------------------------
def _reference_available() -> bool:
    try:
        from oracle import ref_harness
        return bool(ref_harness.reference_available())
    except Exception:
        return False


def _cpu_worker(args):
    seed, n, with_pre, use_ref = args
    import cv2
    cv2.setNumThreads(1)
    import torch
    torch.set_num_threads(1)
    from eagle_b200 import synthetic
    pool = synthetic.make_clip(min(n, 16), W, H, seed=seed, with_frames=with_pre or use_ref, ghost_prob=0.05)
    reps = (n + len(pool["objects"]) - 1) // len(pool["objects"])
    idx = (list(range(len(pool["objects"]))) * reps)[:n]
    hm = [pool["heatmaps"][i] for i in idx]
    ob = [pool["objects"][i] for i in idx]
    if use_ref:  # the reference's own get_coordinates: transforms (cv2.resize + normalise), get_keypoints, cv2 fit, projection, dict
        from oracle import ref_harness
        fr = [pool["frames"][i] for i in idx]
        t0 = time.perf_counter()
        ref_harness.run_reference(fr, np.stack(hm), ob, fps=1)
        return time.perf_counter() - t0
    from oracle import pipeline, preprocess
    t0 = time.perf_counter()
    if with_pre:
        for i in idx:
            preprocess.preprocess_reference_calls(pool["frames"][i])
    pipeline.get_coordinates(hm, ob, W, H)
    return time.perf_counter() - t0


def cpu_path_fps(n_frames: int, workers: int, with_pre: bool = True, use_ref: bool = False):
    """frames/s of the CPU path over n_frames split over `workers` processes (each times its own compute; the slowest counts)."""
    if workers <= 1:
        dt = _cpu_worker((1, n_frames, with_pre, use_ref))
        return n_frames / dt, dt
    import multiprocessing as mp
    per = [n_frames * (r + 1) // workers - n_frames * r // workers for r in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        times = pool.map(_cpu_worker, [(r + 1, per[r], with_pre, use_ref) for r in range(workers) if per[r] > 0])
    dt = max(times)
    return n_frames / dt, dt


def cpu_kind(use_ref: bool):
    if use_ref:
        return "reference", ("unmodified eagle.models.coordinate_model.CoordinateModel.get_coordinates from /root/reference "
                             "(oracle/ref_harness.py: stub detector / network, everything else the reference's own statements)")
    return "port", ("oracle port of coordinate_model.py:221-415 (cv2.resize+normalise, np.argmax decode, synthesis, cv2.findHomography "
                    "cascade, cv2.perspectiveTransform, dict); /root/reference is not mounted on this box")


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    use_ref = _reference_available()
    kind, what = cpu_kind(use_ref)
    sample = 16 * cores
    for _ in range(max(0, args.warmup - 2)):  # data generation dominates a warm-up; one is plenty
        cpu_path_fps(cores, cores, True, use_ref)
    vals = [cpu_path_fps(sample, cores, True, use_ref)[0] for _ in range(args.steps)]
    v = statistics.median(vals)
    line = {"impl": "reference", "metric": "frames/sec, decode->RANSAC homography->projection", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (cv2, numpy)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "sample_frames_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": f"{sample} frames/step over {cores} processes: {what}"},
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


def bind_cores(local_rank: int, local_world: int):
    """One disjoint slice of the visible cores per rank, taken BEFORE any page-locked allocation so that staging buffers
    are first touched by the cores that will fill them.  (The 8-GPU boxes of this pool present one NUMA node, 32 vCPUs,
    every GPU with the same affinity, so there is no closer/farther core to choose -- the slice only stops the ranks'
    copy threads from migrating over each other.)"""
    cores = sorted(os.sched_getaffinity(0))
    if local_world <= 1 or len(cores) < local_world:
        return cores
    per = len(cores) // local_world
    mine = cores[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cores
    return mine


def main():
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 from C (NCCL prints its version
    # banner there) are redirected to stderr, and the line is printed through a private copy of fd 1.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-propagated", action="store_true", help="skip the sparse-keypoint-cadence measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip full_match / ransac_stress / sweep_4k")
    ap.add_argument("--match-frames", type=int, default=MATCH_FRAMES)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    my_cores = bind_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    os.environ["OMP_NUM_THREADS"] = str(len(my_cores))   # torchrun pins it to 1; the host-side tensor ops may use this rank's cores
    import torch
    import torch.distributed as dist
    from eagle_b200 import _native as N
    from eagle_b200.engine import GeometryEngine
    from eagle_b200.sharding import frame_range, gather_to_rank0, pack_results

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    torch.set_num_threads(max(1, len(my_cores)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    F = args.frames
    lo, hi = frame_range(F * world, rank, world)
    assert hi - lo == F
    eng = GeometryEngine(dev)
    frames, hm, foot, count, pool = build_inputs(torch, dev, F, seed=1000 + rank)
    x = torch.empty((F, 3, N.MODEL_H, N.MODEL_W), dtype=torch.float32, device=dev)
    kp = eng.alloc_keypoints(F); fit = eng.alloc_fit(F); proj = eng.alloc_projection(F, MAX_OBJ)
    launches = {"n": 0}

    # the latency-bound tail (synthesis, fit, cadence, projection: a few hundred microseconds of small kernels) runs on a
    # high-priority side stream UNDER the HBM-bound K1, which does not depend on it; the streams join at the end of the step
    side = torch.cuda.Stream(dev, priority=-1)

    def step(ev=None):
        """The hot path over this rank's F frames, inputs resident in HBM.  8 kernel launches on two streams.

        K1's output feeds the keypoint network, which is not part of this path, so nothing in the step depends on it:
        it is launched after K2 on the main stream and the tail of the geometry path runs next to it."""
        main = torch.cuda.current_stream(dev)
        if ev is not None: ev[0].record()
        eng.decode(hm, W, H, 0.3, out=kp)                                # K2  (2 launches)
        if ev is not None: ev[1].record()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if ev is not None: ev[4].record()
            eng.synthesize(kp)                                               # F1  (1)
            eng.fit(kp, out=fit)                                             # K3  (2)
            h_index, attempted = eng.select(fit.status, 1)                   # cadence (1)
            eng.project(fit.H, foot, count, W, H, h_index=h_index, out=proj)  # K4  (1)
            if ev is not None: ev[5].record()
        if ev is not None: ev[2].record()
        eng.preprocess(frames, out=x)                                    # K1  (1 launch)
        if ev is not None: ev[3].record()
        main.wait_stream(side)
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

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches["n"] = 0
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step()
    t1.record()
    barrier()
    gpu_launches = launches["n"]
    ms_step = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    value = F * world / (ms_step * 1e-3)

    # ---- the same step for >= MIN_TIMED_S: what the kernel split and the rooflines are computed from
    n_sus = max(args.steps, int(MIN_TIMED_S * 1e3 / ms_step) + 1)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(n_sus)]
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(n_sus):
        step(evs[i])
    s1.record()
    barrier()
    sus_ms = max_over_ranks(s0.elapsed_time(s1)) / n_sus
    clocks = sampler.stop() if sampler else None
    k2_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)     # decode (argmax + postprocess)
    tail_ms = statistics.mean(e[4].elapsed_time(e[5]) for e in evs)   # synth + fit + select + project (side stream, under K1)
    k1_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)     # preprocess
    del evs
    peak, peak_src = peaks()

    def captured_traffic(summary_file):
        """DRAM bytes per launch from the kept ncu --set full capture of this kernel -- only if the kernel source is still
        the file the capture was taken from (sha-256 recorded with it) and the launch is the captured shape; else None."""
        import hashlib
        d = profile_json(summary_file)
        if not d or F != FRAMES_PER_GPU:
            return None
        for rel, sha in d.get("source_sha256", {}).items():
            try:
                if hashlib.sha256(open(os.path.join(ROOT, rel), "rb").read()).hexdigest() != sha:
                    return None
            except OSError:
                return None
        return d.get("traffic_bytes_per_launch")

    def hbm_roofline(kernel, bytes_per_frame, ms, profile_file, summary_file=None):
        a = F * bytes_per_frame / (ms * 1e-3) / 1e9
        traffic = captured_traffic(summary_file) if summary_file else None
        note = (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch of this shape in the kept ncu --set full capture (profiles/{summary_file}, "
                f"raw page profiles/{profile_file}; the kernel source is unchanged since: sha-256 checked); algorithmic bytes per launch {F * bytes_per_frame}"
                if traffic is not None else f"not measured in this run; the ncu capture of this kernel is kept in profiles/{profile_file}")
        return {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": bytes_per_frame, "launch_ms": ms, "share_of_step": ms / sus_ms,
                "traffic": traffic, "traffic_note": note}

    roofline = hbm_roofline("egl::preprocess_kernel<4,true,false,true> (K1: uint8 BGR frames -> float32 network input; the synthesis/fit/"
                            "cadence/projection kernels run next to it on a side stream)", K1_BYTES, k1_ms, "r2_prof_preprocess_pair_raw.csv", "r2_k1_ncu.json")
    roofline_decode = hbm_roofline("egl::argmax_ldg_kernel (K2 heatmap decode; interval also holds postprocess_kernel, <1%)", HM_BYTES, k2_ms,
                                   "r1_prof_argmax_ldg_raw.csv")
    sustained = {"value": F * world / (sus_ms * 1e-3), "unit": "frames/s", "steps": n_sus, "ms_per_step": sus_ms,
                 "timed_s": sus_ms * n_sus * 1e-3}

    # ---- host->device copy ceiling of this box with all ranks copying at once (explains e2e at N > 1)
    probe_n = 256 * 1024 * 1024
    h_probe = torch.empty(probe_n, dtype=torch.uint8, pin_memory=True)
    d_probe = torch.empty(probe_n, dtype=torch.uint8, device=dev)
    d_probe.copy_(h_probe, non_blocking=True)
    barrier()
    pa = torch.cuda.Event(enable_timing=True); pb = torch.cuda.Event(enable_timing=True)
    pa.record()
    for _ in range(12):
        d_probe.copy_(h_probe, non_blocking=True)
    pb.record()
    barrier()
    my_h2d = 12 * probe_n / (pa.elapsed_time(pb) * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([my_h2d], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank_h2d = [float(v.item()) for v in allv]
    else:
        per_rank_h2d = [my_h2d]
    del h_probe, d_probe
    # ... and the rate of the upload the API uses for frames held in ordinary (pageable) numpy arrays, again with all ranks
    # at once and the thread count the API will use
    from eagle_b200.coordinate_model import upload_threads
    n_stage, n_thr = 64, upload_threads(len(my_cores))
    src_frames = [np.full((H, W, 3), i, dtype=np.uint8) for i in range(n_stage)]
    up_dst = torch.empty((n_stage, H, W, 3), dtype=torch.uint8, device=dev)
    eng.upload_frames(src_frames, up_dst, threads=n_thr)
    barrier()
    ts = time.perf_counter()
    for _ in range(6):
        eng.upload_frames(src_frames, up_dst, threads=n_thr)
    my_stage = 6 * n_stage * H * W * 3 / (time.perf_counter() - ts) / 1e9
    del up_dst
    barrier()
    if world > 1:
        t = torch.tensor([my_stage], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank_stage = [float(v.item()) for v in allv]
    else:
        per_rank_stage = [my_stage]
    del src_frames
    h2d_probe = {"per_rank_GBps": per_rank_h2d, "aggregate_GBps": sum(per_rank_h2d), "cores_per_rank": len(my_cores),
                 "pageable_upload_per_rank_GBps": per_rank_stage, "pageable_upload_aggregate_GBps": sum(per_rank_stage),
                 "pageable_upload_threads_per_rank": n_thr,
                 "note": "per_rank_GBps: pure cudaMemcpyAsync from page-locked memory, all ranks at once -- the ceiling of any host-fed "
                         "number on this box; pageable_upload: egl_upload_frames alone on ordinary (pageable) numpy frames, all ranks at "
                         "once -- the ceiling of the host-fed number when the caller's frames are what the reference passes"}

    # ---- e2e through the public API: host frames -> CoordinateModel.get_coordinates -> reference-format dict
    from eagle_b200.coordinate_model import CoordinateModel, upload_threads
    POOL_HOST = 256     # distinct host frames (1.6 GB), cycled to the clip length: every frame is still staged and copied
    host_pool = frames[:POOL_HOST].cpu().numpy()
    host_frames = [host_pool[i % POOL_HOST] for i in range(F)]
    objs_pool = pool["objects"]
    state = {"i": 0, "h": 0}

    def detector(_frame):       # stand-in for YOLO + BoT-SORT: the pre-generated detections, in frame order
        o = objs_pool[state["i"] % POOL]
        state["i"] += 1
        return o

    def network(xin):           # stand-in for HRNet: device-resident heatmaps, in frame order (a view, no copy)
        n = xin.shape[0]
        s = state["h"] % F
        if s + n > F:
            s = 0
        state["h"] = s + n
        return hm[s:s + n]

    model = CoordinateModel(keypoint_model=network, detect_objects=detector, device=dev, chunk=150)
    model.network_batch = 150
    model.copy_threads = upload_threads(len(my_cores))

    def api_once():
        state["i"] = 0; state["h"] = 0
        return model.get_coordinates(host_frames, fps=25, num_homography=25, num_keypoint_detection=25, verbose=False)

    api_once()
    barrier()
    api_steps = 0
    ta = time.perf_counter()
    while True:
        res = api_once()
        api_steps += 1
        if time.perf_counter() - ta >= MIN_TIMED_S and api_steps >= 2:
            break
    torch.cuda.synchronize()
    api_dt = max_over_ranks((time.perf_counter() - ta) / api_steps)
    st = dict(model.last_stats)
    model.profile = True
    api_once()
    torch.cuda.synchronize()
    prof = dict(model.last_stats)
    model.profile = False
    assert len(res) == F and res[F - 1]["Keypoints"], "the API did not return every frame"
    api_e2e = {"value": F * world / api_dt, "unit": "frames/s", "ms_per_clip": api_dt * 1e3, "clips_timed": api_steps,
               "call": "CoordinateModel.get_coordinates(list of 2250 host frames, fps=25, num_homography=25, num_keypoint_detection=25)",
               "us_per_frame": {"assemble_dict": st["assemble_s"] / F * 1e6, "upload_host_to_device": st["upload_s"] / F * 1e6,
                                "wait_for_upload_after_detector": st["stage_s"] / F * 1e6, "detector_stand_in": st["detect_s"] / F * 1e6,
                                "kernels": prof["kernels_ms"] / F * 1e3, "d2h": prof["d2h_ms"] / F * 1e3},
               "h2d_GBps_achieved": st["h2d_bytes"] / api_dt / 1e9,
               "h2d_GBps_inside_upload_calls": st["h2d_bytes"] / max(st["upload_s"], 1e-9) / 1e9,
               "upload_threads": model.copy_threads,
               "note": "wall clock, dict returned; frames are pageable numpy arrays (256 distinct, cycled) moved by egl_upload_frames "
                       "(worker threads: 4 MiB slices through page-locked rings, H2D per slice) while the detector stand-in runs; "
                       "kernels and the packed D2H on their own streams, assembly (C extension) in a worker thread; the legs overlap, "
                       "so they do not add up to the total"}
    # the same call with the frames handed over in page-locked memory (a (F,H,W,3) uint8 torch tensor): no staging copy
    Fp = 450
    pinned = torch.empty((Fp, H, W, 3), dtype=torch.uint8, pin_memory=True)
    pinned.copy_(frames[:Fp])
    torch.cuda.synchronize()

    def api_pinned():
        state["i"] = 0; state["h"] = 0
        return model.get_coordinates(pinned, fps=25, num_homography=25, num_keypoint_detection=25, verbose=False)

    api_pinned()
    barrier()
    p_steps = 0
    tp = time.perf_counter()
    while time.perf_counter() - tp < MIN_TIMED_S or p_steps < 2:
        resp = api_pinned()
        p_steps += 1
    torch.cuda.synchronize()
    pin_dt = max_over_ranks((time.perf_counter() - tp) / p_steps)
    assert len(resp) == Fp
    api_e2e["frames_in_pinned_memory"] = {"value": Fp * world / pin_dt, "unit": "frames/s", "frames_per_call": Fp, "calls_timed": p_steps,
                                          "h2d_GBps_achieved": Fp * H * W * 3 / pin_dt / 1e9,
                                          "note": "same call, frames passed as one page-locked (F,H,W,3) uint8 torch tensor: the host staging "
                                                  "copy drops out and the PCIe link is the limit"}
    del resp, pinned
    e2e = {"value": api_e2e["value"], "unit": "frames/s", "h2d_bytes_per_step": st["h2d_bytes"] * world, "d2h_bytes_per_step": st["d2h_bytes"] * world,
           "ms_per_step": api_dt * 1e3, "through": "CoordinateModel.get_coordinates (public API, dict returned)"}
    del res
    if model._stream is not None:
        model._stream.close()
    del model, host_frames, host_pool

    # ---- BASELINE configs[2]: the 135 000-frame match, frame-sharded over the ranks through run_sharded
    full_match = None
    if not args.no_extras:
        from eagle_b200.coordinate_model import GeometryPath
        from eagle_b200.sharding import run_sharded
        M = args.match_frames
        mlo, mhi = frame_range(M, rank, world)
        Fm = mhi - mlo
        path = GeometryPath(dev)
        objs_local = [objs_pool[(mlo + i) % POOL] for i in range(Fm)]
        objs_all = [objs_pool[i % POOL] for i in range(M)] if rank == 0 else None

        def chunks(t):
            for s in range(0, Fm, F):
                yield t[:min(F, Fm - s)]

        mstats = {}

        def match_once(assemble):
            return run_sharded(path, chunks(hm), objs_local, W, H, 25, 25, gather_objects=False, objects_on_rank0=objs_all, assemble=assemble,
                               frames_local=chunks(frames), stats=mstats if assemble else None)   # the split only in the untimed-loop run

        match_once(False)
        barrier()
        tm = time.perf_counter()
        match_once(False)
        barrier()
        # run_sharded holds collectives, so every rank must go round the same number of times: the count comes from
        # one agreed estimate, not from each rank's own clock
        m_steps = max(2, int(MIN_TIMED_S / max_over_ranks(time.perf_counter() - tm)) + 1)
        tm = time.perf_counter()
        for _ in range(m_steps):
            match_once(False)
        barrier()
        m_dt = max_over_ranks((time.perf_counter() - tm) / m_steps)
        barrier()
        tm = time.perf_counter()
        out = match_once(True)
        barrier()
        md_dt = max_over_ranks(time.perf_counter() - tm)
        if rank == 0:
            assert len(out) == M
        del out
        full_match = {"workload": f"{M} frames (90 min @25 fps) at 1080p, frame-sharded over {world} GPU(s): {Fm} frames per rank streamed in "
                                  f"{F}-frame chunks through the resident buffers (inputs cycled), K1 + decode + synthesis + fit per rank, "
                                  "NCCL gather of the per-frame records, clip-wide cadence (interval 25) + projection on rank 0",
                      "value": M / m_dt, "unit": "frames/s", "s_per_match": m_dt, "matches_timed": m_steps, "scaling": "strong",
                      "with_dict_on_rank0": {"value": M / md_dt, "unit": "frames/s", "s_per_match": md_dt,
                                             "note": "same, plus one D2H of every result and the reference-format dict of all frames "
                                                     "assembled on rank 0 (detections already on rank 0)"},
                      "rank0_host_split_s": dict(mstats),
                      "timing": "wall clock around run_sharded incl. the gather, barrier + synchronize on both sides, max over ranks"}
        del objs_local, objs_all

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
        p_steps = max(5, int(MIN_TIMED_S * 1e3 / 8.0))
        pe0 = torch.cuda.Event(enable_timing=True); pe1 = torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(p_steps):
            o = prop_once()
        pe1.record()
        barrier()
        p_ms = max_over_ranks(pe0.elapsed_time(pe1) / p_steps)
        ga = torch.cuda.Event(enable_timing=True); gb = torch.cuda.Event(enable_timing=True)
        ga.record(); eng.gray_pyramid(pf, 2, out=prop.pyr); gb.record(); torch.cuda.synchronize()
        g_ms = ga.elapsed_time(gb)
        pyr_bytes = F * (H * W * 3 + int(N.lib.egl_pyramid_bytes(H, W, 2)))
        propagated = {"value": F * world / (p_ms * 1e-3), "unit": "frames/s", "ms_per_clip": p_ms, "frames_per_gpu": F, "clips_timed": p_steps,
                      "workload": f"1080p clip, keypoint interval {kint} (network heads decoded, {kint - 1} of {kint} frames carried by "
                                  f"pyramidal Lucas-Kanade + the reference's filters), homography interval {hint}, rendered pitch frames",
                      "mean_keypoints_per_frame": float(o["count"][:, 0].float().mean().item()),
                      "frames_with_homography": int((o["h_index"] >= 0).sum().item()), "stats": dict(prop.stats),
                      "gray_pyramid": {"ms": g_ms, "bound": "hbm", "achieved": pyr_bytes / (g_ms * 1e-3) / 1e9, "unit": "GB/s",
                                       "frac": pyr_bytes / (g_ms * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_frame": H * W * 3 + int(N.lib.egl_pyramid_bytes(H, W, 2))},
                      "parity": "bit-exact with cv2.calcOpticalFlowPyrLK / the reference's dict (tests/test_gpu_flow.py)"}
        if rank == 0 and not args.no_cpu_baseline:
            # the reference's CPU statements for the same cadence (cv2.calcOpticalFlowPyrLK, cv2.cvtColor, cv2.fitLine,
            # cv2.findHomography, cv2.perspectiveTransform), one thread, on the first chains of the same clip
            from oracle import pipeline as _pipe
            n_cpu = min(F, 25 * kint)
            fr_h = pf[:n_cpu].cpu().numpy()
            hm_h = {i: heads[i // kint].cpu().numpy() for i in range(0, n_cpu, kint)}
            objs_h = [pool["objects"][i % len(pool["objects"])] for i in range(n_cpu)]
            t_c = time.perf_counter()
            _pipe.get_coordinates_propagated(list(fr_h), hm_h, objs_h, fps_ref, 1, 3, library_calls=True)
            dt_c = time.perf_counter() - t_c
            propagated["cpu_baseline"] = {"value": n_cpu / dt_c, "unit": "frames/s", "cores": 1, "kind": "port",
                                          "sample": f"first {n_cpu} frames of the same clip, one thread: the reference's cv2/numpy calls for this "
                                                    "cadence (oracle/pipeline.py, library_calls=True); K1 and the network not included on either side"}
        del pf, heads, prop

    # ---- BASELINE configs[3] and configs[4]: single-GPU configurations, measured at N = 1
    ransac_stress = sweep_4k = None
    if world == 1 and not args.no_extras:
        del frames, hm
        torch.cuda.empty_cache()
        ransac_stress = bench_ransac_stress(torch, eng, N)
        sweep_4k = bench_sweep_4k(torch, eng, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline:
        use_ref = _reference_available()
        kind, what = cpu_kind(use_ref)
        n_cpu = 2048  # ~15 s of single-thread CPU work
        v, dt = cpu_path_fps(n_cpu, 1, True, use_ref)
        cpu = {"value": v, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"{n_cpu} frames of the same 1080p workload, 1 process / 1 thread: {what}"}

    line = {"metric": "frames/sec, decode->RANSAC homography->projection", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 heatmaps / u8 frames / f64 refit", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "frames_per_gpu": F, "heatmaps": [57, 135, 240], "objects_per_frame": MAX_OBJ,
                       "fit": "cv2-compatible adaptive RANSAC (cap 2000) + LS refit + LM", "l2_policy": "inputs larger than L2 "
                       f"({F * (H * W * 3 + HM_BYTES) / 1e9:.1f} GB streamed per step vs 126 MB L2)", "parallelism": f"frame-range x{world}",
                       "streams": "K2, then K1 on the main stream with synthesis/fit/cadence/projection on a high-priority side stream next to it"},
            "sustained": sustained,
            "kernel_ms": {"decode_K2": k2_ms, "synth_fit_select_project": tail_ms, "preprocess_K1": k1_ms,
                          "note": "CUDA events on the stream each part is launched on; the tail runs on a high-priority side stream "
                                  "under K1, so the three do not add up to the step"},
            "stage_fps_without_preprocess": F * world / ((k2_ms + tail_ms) * 1e-3),
            "roofline": roofline, "roofline_decode": roofline_decode, "cpu_baseline": cpu, "e2e": e2e, "api_e2e": api_e2e,
            "h2d_probe": h2d_probe, "gpu_launches": gpu_launches, "clocks": clocks, "full_match": full_match,
            "ransac_stress": ransac_stress, "sweep_4k": sweep_4k, "propagated_cadence": propagated}
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_ransac_stress(torch, eng, N):
    """BASELINE configs[3]: 50 000 frames x K = 4096 hypotheses x 53 landmarks, 40 % outliers, fixed-K mode."""
    from eagle_b200 import synthetic
    from eagle_b200.engine import KeypointSet
    Fs, K = 50000, 4096
    dev = eng.device
    xy, valid, flags, cams = synthetic.stress_point_sets(256, W, H, seed=1)
    reps = (Fs + 255) // 256
    xy = np.tile(xy, (reps, 1, 1))[:Fs]; flags = np.tile(flags, (reps, 1))[:Fs]
    on = [i for i in range(57) if i not in (0, 1, 24, 25)]
    order = np.full((Fs, 64), 255, np.uint8); order[:, :53] = on
    kp = KeypointSet(torch.zeros((Fs, 57), dtype=torch.int32, device=dev), torch.zeros((Fs, 57), device=dev), torch.from_numpy(xy).to(dev),
                     torch.from_numpy(order).to(dev), torch.from_numpy(np.full((Fs, 2), 53, np.int32)).to(dev))
    fit = eng.alloc_fit(Fs)
    for _ in range(3):
        eng.fit(kp, mode=N.FIT_FIXED_K, K=K, seed=1, out=fit)
    torch.cuda.synchronize()
    iters = max(3, int(MIN_TIMED_S * 1e3 / 6.5) + 1)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        eng.fit(kp, mode=N.FIT_FIXED_K, K=K, seed=1, out=fit)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    inl = fit.inlier_mask.cpu().numpy()
    want = np.array([sum(1 << c for c in on if not flags[f, c]) for f in range(256)], dtype=np.int64)
    want = np.tile(want, reps)[:Fs]
    out = {"workload": f"{Fs} frames x K={K} hypotheses x 53 landmarks, 40 % outliers (BASELINE configs[3]); hypothesis kernel "
                       "(ransac_fixedk_kernel, packed FP32) + FP64 refit of every winner",
           "value": Fs / (ms * 1e-3), "unit": "frames/s", "ms_per_batch": ms, "batches_timed": iters,
           "frames_recovering_planted_inliers": float((inl == want).mean())}
    prof = profile_json("r2_fixedk_ncu.json")
    peak32 = profile_json("fp32_peak.json")
    if prof and peak32:
        # executed FP32 operations of the hypothesis kernel per frame (ncu sass counters: FFMA x 2 + FMUL + FADD, packed
        # instructions counted per lane) against the FFMA2 microbenchmark of tools/microbench/fma_peak.cu
        hyp_ms = ms * prof["hypothesis_kernel_share_of_fit"]
        tf = Fs * prof["fp32_flop_executed_per_frame"] / (hyp_ms * 1e-3) / 1e12
        out["hypothesis_kernel"] = {"ms_estimated": hyp_ms, "executed_fp32_TFLOPs": tf, "peak_TFLOPs": peak32["fp32_fma2_tflops"],
                                    "frac": tf / peak32["fp32_fma2_tflops"], "fma_pipe_pct_ncu": prof["sm__pipe_fma_cycles_active_pct"],
                                    "source": "profiles/r2_fixedk_ncu.json (ncu --set full of this launch shape), profiles/fp32_peak.json; "
                                              "the share of the fit spent in the hypothesis kernel is the ncu launch list's, not this run's"}
    return out


def bench_sweep_4k(torch, eng, peak):
    """BASELINE configs[4]: 3840x2160 frames, K1 and K2 over batch sizes 1..256; distinct buffers per launch so that
    nothing is served from L2.  K1 at 4K touches every other source row (scale 4: taps on rows 4y+1, 4y+2)."""
    rows = []
    t_start = time.perf_counter()

    def timeit(fn, n_bufs, budget_s=0.12):
        for i in range(3):
            fn(i % n_bufs)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(0); b.record(); torch.cuda.synchronize()
        iters = max(5, min(2000, int(budget_s * 1e3 / max(a.elapsed_time(b), 1e-3))))
        a.record()
        for i in range(iters):
            fn(i % n_bufs)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    for Fb in [1, 2, 4, 8, 16, 32, 64, 128, 256]:
        nb = max(2, min(16, int(400e6 // (Fb * 2160 * 3840 * 3)) + 2))
        fr = [torch.randint(0, 256, (Fb, 2160, 3840, 3), dtype=torch.uint8, device=eng.device) for _ in range(nb)]
        out = torch.empty((Fb, 3, 540, 960), device=eng.device)
        t1 = timeit(lambda i: eng.preprocess(fr[i], out=out), nb)
        alg1 = Fb * (2160 // 2 * 3840 * 3 + 3 * 540 * 960 * 4)
        del fr, out
        nbh = max(2, min(16, int(400e6 // (Fb * HM_BYTES)) + 2))
        hms = [torch.rand((Fb, 57, 135, 240), device=eng.device) for _ in range(nbh)]
        kp = eng.alloc_keypoints(Fb)
        t2 = timeit(lambda i: eng.decode(hms[i], 3840, 2160, out=kp), nbh)
        alg2 = Fb * HM_BYTES
        del hms
        rows.append({"batch": Fb, "K1_ms": t1, "K1_GBps": alg1 / t1 / 1e6, "K1_frac": alg1 / t1 / 1e6 / peak, "K2_ms": t2,
                     "K2_GBps": alg2 / t2 / 1e6, "K2_frac": alg2 / t2 / 1e6 / peak})
    return {"workload": "3840x2160 frames: K1 preprocess and K2 heatmap decode per launch over batch sizes 1..256 (BASELINE configs[4]); "
                        "K1 bytes = the source rows its taps touch (half the frame) + the float32 output",
            "rows": rows, "timed_s": time.perf_counter() - t_start}


if __name__ == "__main__":
    main()
