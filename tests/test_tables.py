"""Landmark tables: product copy == oracle copy == C header == (when mounted) the reference."""
import os
import re

import numpy as np

from eagle_b200 import pitch
from oracle import landmarks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_matches_oracle():
    assert pitch.NUM_LANDMARKS == len(landmarks.INDEX_TO_NAME) == 57
    for i, name in enumerate(pitch.LANDMARK_NAMES):
        assert landmarks.INDEX_TO_NAME[i] == name
        assert tuple(pitch.WORLD_XYZ[i]) == landmarks.WORLD[name]
    assert list(pitch.OFF_PLANE) == landmarks.NOT_ON_PLANE
    assert pitch.ON_PLANE_MASK == sum(1 << i for i in range(57) if i not in landmarks.NOT_ON_PLANE)
    assert (pitch.PITCH_LENGTH_M, pitch.PITCH_WIDTH_M) == (landmarks.PITCH_X_MAX, landmarks.PITCH_Y_MAX)


def test_reference_tables_if_mounted():
    from oracle import ref_harness
    if not ref_harness.reference_available():
        import pytest
        pytest.skip("/root/reference not mounted")
    import sys
    sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    from eagle.utils import pitch as ref
    assert ref.INTERSECTION_TO_PITCH_POINTS == landmarks.INDEX_TO_NAME
    assert ref.GROUND_TRUTH_POINTS == landmarks.WORLD
    assert ref.NOT_ON_PLANE == landmarks.NOT_ON_PLANE
    assert [ref.PITCH_POINTS_TO_INTERSECTION[n] for n in ref.GROUND_TRUTH_POINTS] == landmarks.WORLD_DICT_ORDER


def test_cuda_constant_table_matches():
    """The __constant__ world table compiled into the kernels is generated from pitch.py."""
    path = os.path.join(ROOT, "eagle_b200", "csrc", "pitch_table.inc")
    txt = open(path).read()
    vals = [float(v) for v in re.findall(r"\{\s*([-0-9.eE+]+)f?,\s*([-0-9.eE+]+)f?\s*\}", txt) for v in v]
    got = np.array(vals, dtype=np.float64).reshape(-1, 2)
    assert got.shape == (57, 2)
    assert np.array_equal(got.astype(np.float32), pitch.WORLD_XY_F32)
