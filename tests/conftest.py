import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_sessionstart(session):
    """Make sure the in-tree CUDA extension exists (nvcc cross-compiles without a GPU) so that a fresh
    checkout can run the suite; a stale library is rebuilt.  Building is skipped silently when nvcc is
    absent -- the tests that need the library then fail with its own "build me first" error."""
    try:
        from eagle_b200.build import build_native
        build_native()
    except Exception as e:  # noqa: BLE001
        print(f"[conftest] could not build libeagle_b200.so: {e}")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


CADENCE_GOLDENS = ("ref_cadence_720p.npz", "ref_cadence_retry_720p.npz")


def cadence_clip(golden_path, with_frames=False):
    """The clip a cadence golden was minted from (oracle/make_golden.py::cadence_fixture): seeded synthetic clip, the
    frames listed in ``blank`` replaced by four off-plane peaks so that their fit fails and the reference retries."""
    import numpy as np
    from eagle_b200 import synthetic
    g = np.load(golden_path)
    clip = synthetic.make_clip(int(g["n_frames"]), int(g["width"]), int(g["height"]), seed=int(g["seed"]), with_frames=with_frames,
                               ghost_prob=0.05)
    synthetic.blank_heatmaps(clip["heatmaps"], [int(b) for b in g["blank"]])
    return g, clip


def cascade_clip(golden_path, with_frames=False):
    """The clip of tests/golden/ref_cascade_clip_720p.npz (oracle/make_golden.py::cascade_clip_fixture): a seeded synthetic
    clip in which the stored heatmap peaks replace four frames, so that their homography comes from cv2.RHO / cv2.LMEDS."""
    import numpy as np
    from eagle_b200 import synthetic
    g = np.load(golden_path)
    clip = synthetic.make_clip(int(g["n_frames"]), int(g["width"]), int(g["height"]), seed=int(g["seed"]), with_frames=with_frames,
                               ghost_prob=0.05)
    for f, pk in zip(g["hard_frames"], g["hard_peaks"]):
        clip["heatmaps"][f] = 0.0
        for c in range(57):
            if pk[c, 0]:
                clip["heatmaps"][f, c, pk[c, 1], pk[c, 2]] = 0.9
    return g, clip
