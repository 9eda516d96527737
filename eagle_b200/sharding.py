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


def pack_results(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Concatenate per-frame result arrays (F, ...) of any dtypes into one (F, bytes) uint8 record
    tensor so that a single collective moves everything."""
    F = tensors[0].shape[0]
    cols = [t.contiguous().view(F, -1).view(torch.uint8) for t in tensors]
    return torch.cat(cols, dim=1)


def unpack_results(records: torch.Tensor, like: Sequence[torch.Tensor]) -> list[torch.Tensor]:
    """Inverse of pack_results for records (F, bytes); ``like`` gives dtypes and per-frame shapes."""
    F = records.shape[0]
    out, o = [], 0
    for t in like:
        per = t.element_size()
        for d in t.shape[1:]:
            per *= int(d)
        out.append(records[:, o:o + per].contiguous().view(t.dtype).view((F,) + tuple(t.shape[1:])))
        o += per
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
