"""Frame-range sharding of a clip across the GPUs of one box and the gather of results to rank 0.

Frames are independent units on this path (every frame has its own heatmaps and boxes), so the
clip is cut into contiguous frame ranges, one per rank, and there is NO data-path collective: the
only communication is one gather of small fixed-size per-frame results to rank 0
(torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced range [lo, hi) of rank ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n_frames * rank // world, n_frames * (rank + 1) // world


def _row_bytes(t: torch.Tensor) -> int:
    per = t.element_size()
    for d in t.shape[1:]:
        per *= int(d)
    return per


def pack_results(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Concatenate per-frame result arrays (F, ...) of any dtypes into one (F, bytes) uint8 record
    tensor so that a single collective moves everything."""
    F = tensors[0].shape[0]
    if F == 0:   # an idle rank still takes part in the gather, with zero rows of the right width
        return torch.empty((0, sum(_row_bytes(t) for t in tensors)), dtype=torch.uint8, device=tensors[0].device)
    cols = [t.contiguous().view(torch.uint8).view(F, _row_bytes(t)) for t in tensors]
    return torch.cat(cols, dim=1)


def unpack_results(records: torch.Tensor, like: Sequence[torch.Tensor]) -> list[torch.Tensor]:
    """Inverse of pack_results for records (F, bytes); ``like`` gives dtypes and per-frame shapes."""
    F = records.shape[0]
    out, o = [], 0
    for t in like:
        per = t.element_size()
        for d in t.shape[1:]:
            per *= int(d)
        # clone(contiguous_format), not contiguous(): a one-row slice counts as contiguous while keeping the record stride
        col = records[:, o:o + per].clone(memory_format=torch.contiguous_format)
        out.append(col.view(t.dtype).view((F,) + tuple(t.shape[1:])))
        o += per
    return out


_PINNED = {}
_NP_DTYPE = {torch.uint8: "uint8", torch.int32: "int32", torch.int64: "int64", torch.float32: "float32", torch.float64: "float64",
             torch.int8: "int8", torch.int16: "int16", torch.bool: "bool"}


def download(tensors: Sequence[torch.Tensor]) -> list:
    """Device tensors (F, ...) -> numpy arrays through ONE device-to-host copy (instead of one per tensor): the tensors are
    laid end to end on the device (planar, so that every array is one contiguous block on the host and no strided
    unpacking is needed there), copied into a reusable page-locked buffer, and copied out of it as owned arrays."""
    import numpy as np
    if tensors[0].shape[0] == 0 or not tensors[0].is_cuda:
        return [t.cpu().numpy() for t in tensors]
    flat = torch.cat([t.contiguous().view(torch.uint8).reshape(-1) for t in tensors])
    n = flat.numel()
    buf = _PINNED.get("buf")
    if buf is None or buf.numel() < n:
        buf = _PINNED["buf"] = torch.empty(max(n, 1 << 20), dtype=torch.uint8, pin_memory=True)
    buf[:n].copy_(flat, non_blocking=True)
    torch.cuda.current_stream(flat.device).synchronize()
    host = buf.numpy()
    out, o = [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        out.append(np.array(host[o:o + nb], copy=True).view(_NP_DTYPE[t.dtype]).reshape(tuple(t.shape)))
        o += nb
    return out


def gather_to_rank0(records: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor | None:
    """Gather per-rank record tensors (F_r, B) to rank 0 in rank order.  ``counts[r]`` = F_r.

    Ranks may hold different F_r (balanced ranges differ by one): every rank pads to max(counts)
    for the collective and rank 0 trims.  Returns the (sum F_r, B) tensor on rank 0, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return records
    fmax = max(counts)
    B = records.shape[1]
    if records.shape[0] < fmax:
        pad = torch.zeros((fmax - records.shape[0], B), dtype=records.dtype, device=records.device)
        records = torch.cat([records, pad])
    if rank == 0:
        bufs = [torch.empty((fmax, B), dtype=records.dtype, device=records.device) for _ in range(world)]
        dist.gather(records, bufs, dst=0, group=group)
        return torch.cat([b[:counts[r]] for r, b in enumerate(bufs)])
    dist.gather(records, None, dst=0, group=group)
    return None


_SIDE: dict = {}


def _side_stream(device):
    """One high-priority stream per device for the small kernels that run under the bandwidth-bound ones."""
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device, priority=-1)
    return _SIDE[key]


class FewLandmarksError(RuntimeError):
    """A frame decoded fewer than four landmarks: the reference rescues it by optical flow from its neighbours
    (coordinate_model.py:287-311), which the per-frame sharded path cannot do.  Use run_sharded_propagated with
    keypoint_interval=1 (or CoordinateModel.get_coordinates, which reroutes by itself)."""


def run_sharded(path, heatmaps_local, objects_local: Sequence[dict], width: int, height: int, fps: int,
                homography_interval: int = 1, group=None, *, gather_objects: bool = True,
                objects_on_rank0: Sequence[dict] | None = None, assemble: bool = True, frames_local=None,
                stats: dict | None = None):
    """The geometry path for one clip sharded by contiguous frame range over the ranks of ``group``.

    Every rank calls this with ITS frames' heatmaps and detector dicts, in rank order of the clip.  ``heatmaps_local``
    is one (F_r,57,h,w) CUDA tensor or an iterable of such tensors (consecutive chunks of the rank's range: a full
    match does not fit in HBM at once, so the caller streams chunks through fixed buffers); ``frames_local``, if given,
    is the matching tensor / iterable of (n,H,W,3) uint8 frames, which then go through K1 as well (the network input
    is produced and dropped: the network is not part of this path).  The bandwidth- and compute-heavy kernels (decode,
    synthesis, RANSAC + refit) run on each rank's range; the small per-frame results are gathered to rank 0, which
    evaluates the homography cadence over the WHOLE clip (it carries state across shard boundaries exactly as the
    reference's sequential loop does), projects and assembles the reference-format dict.  Returns that dict on rank 0
    and None elsewhere (``assemble=False``: the per-frame arrays instead of the dict).  ``gather_objects=False`` (the same
    on EVERY rank) skips the pickled gather of the detection dicts; rank 0 then passes the whole clip's detections as
    ``objects_on_rank0`` (e.g. the detector ran there).

    Precondition (checked, FewLandmarksError on every rank): every frame decodes at least four landmarks.
    Works with NCCL (GPU tensors) and with gloo (records staged through the host, used by the single-GPU tests)."""
    from .boxes import max_objects, objects_to_arrays
    from .coordinate_model import assemble_frames
    from .engine import FitResult, KeypointSet

    import time
    t_mark = [time.perf_counter()]

    def lap(name, sync=False):   # wall-clock split for ``stats``; sync=True charges the device work enqueued so far to this lap
        if stats is not None:
            if sync and torch.cuda.is_available():
                torch.cuda.synchronize()
            now = time.perf_counter()
            stats[name] = stats.get(name, 0.0) + now - t_mark[0]
            t_mark[0] = now

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    e = path.engine
    objects_local = objects_local if type(objects_local) is list else list(objects_local)
    chunks = [heatmaps_local] if torch.is_tensor(heatmaps_local) else heatmaps_local
    fchunks = None if frames_local is None else ([frames_local] if torch.is_tensor(frames_local) else frames_local)
    kps, fits = [], []
    x = None
    fit_iter = iter(fchunks) if fchunks is not None else None
    # Streamed chunks: synthesis + fit of chunk c (a few hundred microseconds of small, latency-bound kernels) run on a
    # high-priority side stream under the HBM-bound K1 / decode of chunk c+1.
    overlap = e.device.type == "cuda" and not torch.is_tensor(heatmaps_local)
    main = torch.cuda.current_stream(e.device) if overlap else None
    side = _side_stream(e.device) if overlap else None
    for hm in chunks:
        if fit_iter is not None:
            fr = next(fit_iter)
            if x is None or x.shape[0] < fr.shape[0]:
                x = torch.empty((fr.shape[0], 3, 540, 960), dtype=torch.float32, device=e.device)
            e.preprocess(fr, out=x[:fr.shape[0]])
        kp = e.decode(hm, width, height, path.keypoint_conf)
        kps.append(kp)
        if overlap:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if path.synthesis:
                    e.synthesize(kp)
                fits.append(e.fit(kp, mode=path.fit_mode, K=path.max_iters, thr=path.thr))
        else:
            if path.synthesis:
                e.synthesize(kp)
            fits.append(e.fit(kp, mode=path.fit_mode, K=path.max_iters, thr=path.thr))
    if overlap:
        main.wait_stream(side)
        for f in fits:   # allocated on the side stream, read on the main stream from here on
            for t in (f.H, f.used_mask, f.inlier_mask, f.status):
                t.record_stream(main)
    cat = lambda xs: torch.cat(xs) if len(xs) > 1 else xs[0]
    kp = KeypointSet(None, None, cat([k.xy for k in kps]), cat([k.order for k in kps]), cat([k.count for k in kps]))
    fit = FitResult(cat([f.H for f in fits]), cat([f.used_mask for f in fits]), cat([f.inlier_mask for f in fits]),
                    cat([f.status for f in fits]), None)
    F_r = kp.xy.shape[0]
    if len(objects_local) != F_r:
        raise ValueError(f"rank {rank}: {F_r} frames of heatmaps but {len(objects_local)} detection dicts")
    lap("enqueue_s")
    # the foot points are packed while the kernels above still run (nothing has synchronised yet)
    P_local = max(1, max_objects(objects_local))
    foot_h, cnt_h = objects_to_arrays(objects_local, P_local)
    # per-rank frame counts, the clip-wide maximum object count and the < 4 landmarks flag (host-side, tiny)
    lap("pack_foot_s")
    few = bool((kp.count[:, 1] < 4).any().item()) if F_r else False
    lap("wait_kernels_s")
    local_meta = (F_r, P_local, few)
    metas = [None] * world
    if world > 1:
        dist.all_gather_object(metas, local_meta, group=group)
    else:
        metas = [local_meta]
    if any(m[2] for m in metas):
        raise FewLandmarksError("a frame of the clip decoded fewer than four landmarks; the per-frame sharded path does not apply "
                                f"(ranks affected: {[r for r, m in enumerate(metas) if m[2]]})")
    counts = [m[0] for m in metas]
    P = max(m[1] for m in metas)
    if P > P_local:   # another rank saw a busier frame: same record width everywhere
        import numpy as np
        foot_h = np.concatenate([foot_h, np.zeros((F_r, P - P_local, 2), np.float32)], axis=1)
    foot = torch.from_numpy(foot_h).to(e.device); cnt = torch.from_numpy(cnt_h).to(e.device)
    like = [kp.xy, kp.order, kp.count, fit.H, fit.used_mask, fit.inlier_mask, fit.status, foot, cnt]
    rec = pack_results(like)
    backend = dist.get_backend(group) if (dist.is_initialized() and world > 1) else None
    if backend == "gloo":
        rec = rec.cpu()
    all_rec = gather_to_rank0(rec, counts, group) if world > 1 else rec
    lap("gather_s", sync=True)
    objs_all = None
    if assemble and gather_objects:
        objs_all = [None] * world
        if world > 1:
            dist.gather_object(objects_local, objs_all if rank == 0 else None, dst=0, group=group)
        else:
            objs_all = [objects_local]
    if rank != 0:
        return None
    all_rec = all_rec.to(e.device)
    xy, order, count, H, used, inl, status, foot_a, cnt_a = unpack_results(all_rec, like)
    h_index, attempted = e.select(status.contiguous(), homography_interval)
    proj = e.project(H.contiguous(), foot_a.contiguous(), cnt_a.contiguous(), width, height, h_index=h_index)
    lap("rank0_cadence_project_s", sync=True)
    arrays = download([xy, order, count, used, inl, status, attempted, h_index, proj.coords_i, proj.in_bounds, proj.bounds])
    lap("rank0_download_s")
    if not assemble:
        return arrays
    if gather_objects:
        objects_per_frame = [o for part in objs_all for o in part]
    elif objects_on_rank0 is None:
        raise ValueError("run_sharded(gather_objects=False) needs the clip's detections as objects_on_rank0 on rank 0")
    else:
        objects_per_frame = list(objects_on_rank0)
    if len(objects_per_frame) != sum(counts):
        raise ValueError(f"{sum(counts)} frames in the clip but {len(objects_per_frame)} detection dicts on rank 0")
    res = assemble_frames(objects_per_frame, fps, 0, *arrays)
    lap("assemble_s")
    return res


def chain_range(n_frames: int, keypoint_interval: int, rank: int, world: int) -> tuple[int, int]:
    """Frame range [lo, hi) of ``rank`` when the clip is cut between chains (a chain = one network frame and the
    frames propagated from it): boundaries are multiples of the keypoint interval.  The chains are dealt out so
    that the LOW ranks are never empty (a clip of fewer chains than ranks leaves the high ranks idle): rank 0
    assembles the result and every non-empty rank's predecessor holds the frame before its first."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    nc = (n_frames + keypoint_interval - 1) // keypoint_interval
    base, rem = divmod(nc, world)
    c0 = rank * base + min(rank, rem)
    c1 = c0 + base + (1 if rank < rem else 0)
    return min(c0 * keypoint_interval, n_frames), min(c1 * keypoint_interval, n_frames)


def run_sharded_propagated(engine, frames_local: torch.Tensor, head_heatmaps_local: torch.Tensor, detect, objects_local: Sequence[dict],
                           fps: int, keypoint_interval: int, homography_interval: int, first_frame: int, calibration: bool = False,
                           keypoint_conf: float = 0.3, group=None):
    """The sparse keypoint cadence (eagle_b200.propagation) for one clip cut between chains over the ranks.

    Rank r holds the frames of its chain_range -- plus, when first_frame > 0, ONE leading frame (the last frame
    of rank r-1) -- and the network's heatmaps of its chain heads.  Every rank runs its parallel pass at once;
    then the boundary state (final keypoint set + retry flag, < 1 KB) is handed from rank to rank while each
    repairs the few chains that depend on it, and the per-frame results are gathered to rank 0, which looks up
    the homography cadence over the whole clip, projects and assembles the dict.  Returns it on rank 0."""
    from .boxes import objects_to_arrays
    from .coordinate_model import assemble_frames
    from .engine import KeypointSet
    from .propagation import PropagatedPath, finalize

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    backend = dist.get_backend(group) if (dist.is_initialized() and world > 1) else None
    dev = engine.device
    height, width = frames_local.shape[1], frames_local.shape[2]
    prop = PropagatedPath(engine, keypoint_conf)
    # a rank past the end of a short clip (chain_range leaves the high ranks empty) only passes the boundary state on
    n_local = frames_local.shape[0] - (1 if first_frame > 0 else 0) if frames_local is not None else 0
    idle = n_local <= 0
    if idle and rank == 0:
        raise ValueError("run_sharded_propagated: rank 0 holds no frames (empty clip)")
    if not idle:
        from .propagation import FirstPieceTooShort
        try:
            # rank 0's forward scan for the first usable frame (:290-297) may not leave its own range: if it would have
            # to, the clip cannot be sharded this way and every rank is told so (instead of a silently different dict)
            prop.start(frames_local, head_heatmaps_local, detect, keypoint_interval, homography_interval, calibration,
                       first_frame=first_frame, clip_continues=rank < world - 1)
            short = False
        except FirstPieceTooShort:
            short = True
    else:
        short = False
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, short, group=group)
        short = any(flags)
    if short:
        raise FirstPieceTooShort("frame 0 has fewer than four landmarks and no frame of rank 0's range has four: the reference's "
                                 "forward scan (coordinate_model.py:290-297) would continue into the next rank's frames; "
                                 "run the clip on fewer ranks or through CoordinateModel.get_coordinates")

    def carry_like():
        return [torch.zeros((1, 57, 2), dtype=torch.int32, device=dev), torch.zeros((1, 64), dtype=torch.uint8, device=dev),
                torch.zeros((1, 2), dtype=torch.int32, device=dev), torch.zeros((1, 64), dtype=torch.uint8, device=dev),
                torch.zeros((1, 1), dtype=torch.int32, device=dev)]

    carry = None
    if rank > 0:
        buf = pack_results(carry_like())
        buf = buf.cpu() if backend == "gloo" else buf
        dist.recv(buf, src=rank - 1, group=group)
        xy, order, count, src, retry = unpack_results(buf.to(dev), carry_like())
        carry = {"retry": int(retry.item()), "kp": KeypointSet(None, None, xy, order, count, src)}
    if not idle:
        prop.repair(carry)
    if rank < world - 1:
        co = carry if idle else prop.carry_out
        msg = pack_results([co["kp"].xy, co["kp"].order, co["kp"].count, co["kp"].src,
                            torch.tensor([[co["retry"]]], dtype=torch.int32, device=dev)])
        dist.send(msg.cpu() if backend == "gloo" else msg, dst=rank + 1, group=group)
    if idle:
        z = lambda *shape, dtype: torch.zeros(shape, dtype=dtype, device=dev)
        out = {"xy": z(0, 57, 2, dtype=torch.int32), "order": z(0, 64, dtype=torch.uint8), "count": z(0, 2, dtype=torch.int32),
               "src": z(0, 64, dtype=torch.uint8), "H": z(0, 9, dtype=torch.float64), "fit_ok": z(0, dtype=torch.uint8),
               "status": z(0, dtype=torch.int32), "inlier_mask": z(0, dtype=torch.int64)}
    else:
        out = prop.outputs()

    F_r = out["xy"].shape[0]
    local_meta = (F_r, max([1] + [sum(len(v) for v in o.values()) for o in objects_local]), dict(prop.stats))
    metas = [None] * world
    if world > 1:
        dist.all_gather_object(metas, local_meta, group=group)
    else:
        metas = [local_meta]
    counts = [m[0] for m in metas]
    P = max(m[1] for m in metas)
    foot_h, cnt_h = objects_to_arrays(list(objects_local), P)
    foot = torch.from_numpy(foot_h).to(dev); cnt = torch.from_numpy(cnt_h).to(dev)
    like = [out["xy"], out["order"], out["count"], out["src"], out["H"], out["fit_ok"], out["status"], out["inlier_mask"], foot, cnt]
    rec = pack_results(like)
    if backend == "gloo":
        rec = rec.cpu()
    all_rec = gather_to_rank0(rec, counts, group) if world > 1 else rec
    objs_all = [None] * world
    if world > 1:
        dist.gather_object(list(objects_local), objs_all if rank == 0 else None, dst=0, group=group)
    else:
        objs_all = [list(objects_local)]
    if rank != 0:
        return None
    xy, order, count, src, H, fit_ok, status, inl, foot_a, cnt_a = unpack_results(all_rec.to(dev), like)
    whole = finalize(engine, [{"xy": xy, "order": order, "count": count, "src": src, "H": H.contiguous(), "fit_ok": fit_ok.contiguous(),
                               "status": status.contiguous(), "inlier_mask": inl}])
    proj = engine.project(whole["H"], foot_a.contiguous(), cnt_a.contiguous(), width, height, h_index=whole["h_index"])
    objects_per_frame = [o for part in objs_all for o in part]
    xy_h, order_h, count_h, hi_h, ci_h, ib_h, bd_h, src_h = download([xy, order, count, whole["h_index"], proj.coords_i, proj.in_bounds,
                                                                      proj.bounds, src])
    res = assemble_frames(objects_per_frame, fps, 0, xy_h, order_h, count_h, None, None, None, None, hi_h, ci_h, ib_h, bd_h, kp_src=src_h)
    run_sharded_propagated.last_stats = [m[2] for m in metas]
    return res
