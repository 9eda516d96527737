"""torchrun --nproc-per-node N tools/sharded_check.py : run_sharded over NCCL on N GPUs, checked against
the golden dict the unmodified reference produced (tests/golden/ref_cadence_720p.npz)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.coordinate_model import GeometryPath  # noqa: E402
from eagle_b200.sharding import frame_range, run_sharded  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_cadence_720p.npz"))
n, w, h = int(g["n_frames"]), int(g["width"]), int(g["height"])
clip = synthetic.make_clip(n, w, h, seed=int(g["seed"]), ghost_prob=0.05)
lo, hi = frame_range(n, rank, world)
path = GeometryPath(f"cuda:{local}")
res = run_sharded(path, torch.from_numpy(clip["heatmaps"][lo:hi]).cuda(), clip["objects"][lo:hi], w, h, fps=int(g["fps"]), homography_interval=5)
if rank == 0:
    ok = json.dumps(res, default=float, sort_keys=True) == str(g["result_json"])
    print(f"run_sharded over NCCL, world={world}: dict identical to the reference's: {ok}")
    assert ok
dist.destroy_process_group()
