"""torchrun --nproc-per-node N tools/sharded_flow_check.py : the sparse keypoint cadence sharded between chains
over NCCL on N GPUs (eagle_b200.sharding.run_sharded_propagated), checked against the dict the unmodified
reference produced for the same clip (tests/golden/ref_flow_360p.npz: network every 8th frame, 26 frames)."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.engine import GeometryEngine  # noqa: E402
from eagle_b200.sharding import chain_range, run_sharded_propagated  # noqa: E402
from oracle.ref_harness import stamp_frames  # noqa: E402  (checker side: the golden was made from stamped frames)

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_flow_360p.npz"))
n, w, h, fps = int(g["n_frames"]), int(g["width"]), int(g["height"]), int(g["fps"])
clip = synthetic.make_flow_clip(n, w, h, seed=int(g["seed"]), pan_px=float(g["pan_px"]))
frames = np.stack(stamp_frames(clip["frames"]))
assert hashlib.sha256(frames.tobytes()).hexdigest() == str(g["frames_sha256"])
k = max(1, int(fps / int(g["num_keypoint_detection"]))); hint = max(1, int(fps / int(g["num_homography"])))
eng = GeometryEngine(f"cuda:{local}")
lo, hi = chain_range(n, k, rank, world)
halo = 1 if lo > 0 else 0
dev_frames = torch.from_numpy(np.ascontiguousarray(frames[lo - halo:hi])).cuda()
hm = torch.from_numpy(clip["heatmaps"]).cuda()
heads = hm[lo:hi:k].contiguous()                       # stand-in for the network on this rank's chain heads
res = run_sharded_propagated(eng, dev_frames, heads, lambda i: hm[i:i + 1].contiguous(), clip["objects"][lo:hi], fps, k, hint, first_frame=lo)
if rank == 0:
    ok = json.dumps(res, default=float, sort_keys=True) == str(g["result_json"])
    print(f"run_sharded_propagated over NCCL, world={world}, ranges by chain: dict identical to the reference's: {ok}")
    assert ok
dist.destroy_process_group()
