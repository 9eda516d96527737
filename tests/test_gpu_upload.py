"""egl_upload_frames: the host frames of get_coordinates(frames, ...) (coordinate_model.py:188,221: a Python list of
pageable HxWx3 uint8 arrays) land on the device byte for byte, whatever the thread count, frame count or frame size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def engine():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eagle_b200.engine import GeometryEngine
    return GeometryEngine("cuda:0")


@pytest.mark.parametrize("shape,n,threads", [((1080, 1920, 3), 37, 8), ((720, 1280, 3), 5, 16), ((33, 47, 3), 70, 3),
                                             ((2160, 3840, 3), 3, 4), ((1, 1, 3), 1, 8), ((1080, 1920, 3), 2, 1)])
def test_upload_is_byte_exact(engine, shape, n, threads):
    rng = np.random.default_rng(n)
    frames = [rng.integers(0, 256, shape, dtype=np.uint8) for _ in range(n)]
    dev = torch.zeros((n + 1,) + shape, dtype=torch.uint8, device="cuda")
    out = engine.upload_frames(frames, dev, threads=threads)
    assert out.shape[0] == n
    got = dev.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], frames[i]), i
    assert not got[n].any()          # nothing written past the last frame


def test_upload_rejects_bad_arguments(engine):
    from eagle_b200 import _native as N
    dev = torch.zeros((2, 8, 8, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(Exception):
        engine.upload_frames([np.zeros((8, 8, 3), np.uint8), np.zeros((4, 8, 3), np.uint8)], dev)   # ragged frame sizes
    with pytest.raises(Exception):
        engine.upload_frames([np.zeros((16, 8, 3), np.uint8)[::2]], dev)                             # not contiguous
    assert N.lib.egl_upload_frames(None, 1, 10, None, 1) == 1 and b"null" in N.lib.egl_last_error()
    assert engine.upload_frames([], dev).shape[0] == 0


def test_repeated_uploads_reuse_the_rings(engine):
    rng = np.random.default_rng(0)
    dev = torch.zeros((16, 270, 480, 3), dtype=torch.uint8, device="cuda")
    for rep in range(5):
        frames = [rng.integers(0, 256, (270, 480, 3), dtype=np.uint8) for _ in range(16)]
        engine.upload_frames(frames, dev, threads=2 + rep)
        assert np.array_equal(dev.cpu().numpy(), np.stack(frames))


def test_two_concurrent_uploads_are_byte_exact(engine):
    """Two calls at once are served from separate rings (the streaming layer starts chunk c+1's upload while chunk c's
    drains); a third one queues behind them."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(3)
    sets = [[rng.integers(0, 256, (540, 960, 3), dtype=np.uint8) for _ in range(24)] for _ in range(3)]
    devs = [torch.zeros((24, 540, 960, 3), dtype=torch.uint8, device="cuda") for _ in range(3)]
    for rep in range(3):
        with ThreadPoolExecutor(3) as pool:
            list(pool.map(lambda k: engine.upload_frames(sets[k], devs[k], threads=3), range(3)))
        for k in range(3):
            assert np.array_equal(devs[k].cpu().numpy(), np.stack(sets[k])), (rep, k)
            devs[k].zero_()


def test_api_with_two_uploads_in_flight_returns_the_same_dict():
    """DenseStream.uploads_in_flight = 2 changes when the uploads are started, not what is computed."""
    import json
    from eagle_b200 import synthetic
    from eagle_b200.coordinate_model import CoordinateModel
    clip = synthetic.make_clip(23, 640, 360, seed=9, with_frames=True, ghost_prob=0.05)
    hm = torch.from_numpy(clip["heatmaps"]).cuda()
    out = []
    for inflight in (1, 2):
        st = {"i": 0, "h": 0}

        def detector(_f):
            st["i"] += 1
            return clip["objects"][st["i"] - 1]

        def network(x):
            s = st["h"]
            st["h"] += x.shape[0]
            return hm[s:s + x.shape[0]]

        model = CoordinateModel(keypoint_model=network, detect_objects=detector, chunk=5)
        model.network_batch = 5
        model.uploads_in_flight = inflight
        res = model.get_coordinates(list(clip["frames"]), fps=5, num_homography=5, num_keypoint_detection=5, verbose=False)
        assert len(res) == 23
        out.append(json.dumps(res, default=float))
        model._stream.close()
    assert out[0] == out[1]
