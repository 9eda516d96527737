"""Frame-range sharding + gather to rank 0 on CPU: world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eagle_b200.sharding import frame_range, gather_to_rank0, pack_results, unpack_results


def test_frame_ranges_partition_the_clip():
    for n in (0, 1, 7, 2250, 135000):
        for w in (1, 2, 3, 4, 8):
            r = [frame_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_range(10, 4, 4)


def test_pack_unpack_roundtrip():
    a = torch.randn(5, 9, dtype=torch.float64); b = torch.arange(5, dtype=torch.int32); c = torch.randint(0, 255, (5, 23, 2), dtype=torch.uint8)
    rec = pack_results([a, b, c])
    assert rec.shape == (5, 72 + 4 + 46)
    x, y, z = unpack_results(rec, [a, b, c])
    assert torch.equal(x, a) and torch.equal(y, b) and torch.equal(z, c)


def _worker(rank, world, port, n_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = frame_range(n_frames, rank, world)
    H = torch.arange(lo, hi, dtype=torch.float64).view(-1, 1).repeat(1, 9)
    status = torch.arange(lo, hi, dtype=torch.int32) % 3
    counts = [frame_range(n_frames, r, world)[1] - frame_range(n_frames, r, world)[0] for r in range(world)]
    got = gather_to_rank0(pack_results([H, status]), counts)
    if rank == 0:
        h, s = unpack_results(got, [H, status])
        ok = torch.equal(h[:, 0], torch.arange(n_frames, dtype=torch.float64)) and torch.equal(s, torch.arange(n_frames, dtype=torch.int32) % 3)
        out.put(bool(ok))
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 11), (3, 10)])
def test_gather_to_rank0_gloo(world, n_frames):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_chain_range_cuts_between_chains():
    """Ranges of the sparse keypoint cadence: contiguous, covering, every boundary a multiple of the interval."""
    from eagle_b200.sharding import chain_range
    for n in (1, 7, 8, 9, 26, 2250, 135000):
        for k in (1, 3, 8, 25):
            for world in (1, 2, 3, 8):
                prev = 0
                for r in range(world):
                    lo, hi = chain_range(n, k, r, world)
                    assert lo == prev and lo <= hi <= n and (lo % k == 0 or lo == n)
                    prev = hi
                assert prev == n
                # a short clip leaves the HIGH ranks idle, never rank 0 (which assembles) nor a rank between two busy ones
                sizes = [chain_range(n, k, r, world)[1] - chain_range(n, k, r, world)[0] for r in range(world)]
                assert sizes[0] > 0 and all(a > 0 or b == 0 for a, b in zip(sizes, sizes[1:]))
