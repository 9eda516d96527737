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
