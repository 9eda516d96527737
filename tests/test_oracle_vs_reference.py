"""Live check of the oracle against the UNMODIFIED reference (authoring container only)."""
import json
import warnings

import numpy as np
import pytest

from eagle_b200 import synthetic
from oracle import pipeline, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("w,h,seed", [(1280, 720, 21), (1920, 1080, 22)])
def test_oracle_pipeline_equals_reference(w, h, seed):
    warnings.simplefilter("ignore")
    clip = synthetic.make_clip(5, w, h, seed=seed, with_frames=True, ghost_prob=0.08)
    ref, rec = ref_harness.run_reference(clip["frames"], clip["heatmaps"], clip["objects"])
    trace = []
    got = pipeline.get_coordinates(clip["heatmaps"], clip["objects"], w, h, trace=trace)
    assert json.dumps(ref, default=float, sort_keys=True) == json.dumps(got, default=float, sort_keys=True)
    for f, t in zip(rec.fits, trace):
        assert np.array_equal(f["img_pts"], t["img_pts"]) and np.array_equal(f["H"], t["H"])


def test_reference_processor_consumes_the_dict():
    """The consumer of the path's output -- eagle/processor.py Processor.create_dataframe/format_data --
    accepts the dict in the format this repository emits (the oracle dict is JSON-identical to the CUDA
    path's, see tests/test_gpu_parity.py)."""
    warnings.simplefilter("ignore")
    import sys
    ref_harness.load_reference()
    from eagle.processor import Processor
    clip = synthetic.make_clip(30, 1280, 720, seed=5, with_frames=True, ghost_prob=0.05)
    coords = pipeline.get_coordinates(clip["heatmaps"], clip["objects"], 1280, 720, fps=5, num_homography=1)
    coords = json.loads(json.dumps(coords, default=float))              # what main.py writes/reads back
    coords = {int(k): v for k, v in coords.items()}
    proc = Processor(coords, list(clip["frames"]), 5, filter_ball_detections=False)
    raw = proc.create_dataframe()
    assert len(raw) == 30 and "Ball" in raw.columns and any(c.startswith("Player_") for c in raw.columns)
    df, team_mapping = proc.process_data(smooth=False)      # main.py:35
    out = proc.format_data(df)                               # main.py:41
    assert len(out) == len(df) > 0 and {"Boundaries", "Coordinates", "Coordinates_video"} <= set(out.columns)
