"""CPU tests of the host glue written in C (eagle_b200/csrc/assemble.c): the record assembly and the foot-point packing
must return exactly what their readable Python statements return -- same keys, order, values and Python types."""
import json

import numpy as np
import pytest

from eagle_b200 import _native as N
from eagle_b200 import coordinate_model as cm
from eagle_b200 import synthetic
from eagle_b200.boxes import max_objects, objects_to_arrays, objects_to_arrays_py
from eagle_b200.streaming import record_bytes, record_layout, unpack_record


def typed(d):
    if isinstance(d, dict):
        return {k: typed(v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return (type(d).__name__, [typed(v) for v in d])
    return (type(d).__name__, d if d == d else "nan")


def random_case(seed, F=96, P=23):
    rng = np.random.default_rng(seed)
    clip = synthetic.make_clip(16, 1280, 720, seed=seed, ghost_prob=0.05)
    objs = [clip["objects"][i % 16] for i in range(F)]
    objs[3] = {"Player": {}, "Goalkeeper": {}}                                            # nobody in the frame
    objs[4] = {"Player": {np.int64(9): {"BBox": np.array([1.7, 2, 70000, -4]), "Confidence": np.float32(0.5), "Bottom_center": (3, 4.5)}},
               "Goalkeeper": {}, "Ball": {"1": {"BBox": [5, 6, 7, 8], "Confidence": 0.25, "Bottom_center": [6, 8]}}}
    xy = rng.integers(-50, 4000, (F, 57, 2)).astype(np.int32)
    order = np.stack([rng.permutation(57) for _ in range(F)]).astype(np.uint8)
    order = np.concatenate([order, np.full((F, 7), 255, np.uint8)], axis=1)
    count = np.stack([rng.integers(0, 58, F), rng.integers(0, 58, F)], axis=1).astype(np.int32)
    inl = rng.integers(0, 2 ** 57, F).astype(np.int64)
    status = rng.integers(0, 3, F).astype(np.int32)
    att = rng.integers(0, 2, F).astype(np.uint8)
    hidx = rng.integers(-1, F, F).astype(np.int32)
    ci = rng.integers(-10, 120, (F, P, 2)).astype(np.int64)
    ib = (rng.random((F, P)) < 0.8).astype(np.uint8)
    bd = rng.normal(50, 30, (F, 4))
    bd[rng.random(F) < 0.2] = np.nan
    src = rng.integers(0, 3, (F, 64)).astype(np.uint8)
    return objs, xy, order, count, inl, status, att, hidx, ci, ib, bd, src


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_assembler_equals_the_python_statement(seed):
    objs, xy, order, count, inl, status, att, hidx, ci, ib, bd, src = random_case(seed)
    want = cm.assemble_frames_py(objs, 25, 11, xy, order, count, inl, inl, status, att, hidx, ci, ib, bd)
    got = cm.assemble_frames(objs, 25, 11, xy, order, count, inl, inl, status, att, hidx, ci, ib, bd)
    assert list(got) == list(want) and typed(got) == typed(want)
    assert json.dumps(got, default=float) == json.dumps(want, default=float)
    want = cm.assemble_frames_py(objs, 30, 0, xy, order, count, None, None, None, None, hidx, ci, ib, bd, kp_src=src)
    got = cm.assemble_frames(objs, 30, 0, xy, order, count, None, None, None, None, hidx, ci, ib, bd, kp_src=src)
    assert typed(got) == typed(want) and json.dumps(got, default=float) == json.dumps(want, default=float)


def test_c_assembler_appends_chunks_and_reports_bad_input():
    objs, xy, order, count, inl, status, att, hidx, ci, ib, bd, _ = random_case(5, F=40)
    whole = cm.assemble_frames(objs, 25, 0, xy, order, count, inl, inl, status, att, hidx, ci, ib, bd)
    out = {}
    for s in (0, 16, 32):
        e = min(40, s + 16)
        cm.assemble_frames(objs[s:e], 25, s, xy[s:e], order[s:e], count[s:e], inl[s:e], inl[s:e], status[s:e], att[s:e], hidx[s:e], ci[s:e],
                           ib[s:e], bd[s:e], out=out)
    assert typed(out) == typed(whole)
    with pytest.raises(ValueError):   # more detections than the projection arrays hold
        cm.assemble_frames(objs, 25, 0, xy, order, count, inl, inl, status, att, hidx, ci[:, :2], ib[:, :2], bd)
    with pytest.raises(KeyError):
        cm.assemble_frames([{"Player": {1: {"Confidence": 1.0, "Bottom_center": [1, 2]}}}], 25, 0, xy[:1], order[:1], count[:1], inl[:1], inl[:1],
                           status[:1], att[:1], hidx[:1], ci[:1], ib[:1], bd[:1])
    assert cm.assemble_frames([], 25, 0, xy[:0], order[:0], count[:0], inl[:0], inl[:0], status[:0], att[:0], hidx[:0], ci[:0], ib[:0], bd[:0]) == {}


def test_foot_point_packing_equals_the_python_statement():
    objs = random_case(7)[0]
    P = max_objects(objs)
    assert P == max(sum(len(v) for v in o.values()) for o in objs)
    f1, c1 = objects_to_arrays(objs, P)
    f2, c2 = objects_to_arrays_py(objs, P)
    assert np.array_equal(f1, f2) and np.array_equal(c1, c2) and f1.dtype == np.float32 and c1.dtype == np.int32
    foot = np.full((len(objs), P + 2, 2), 7.0, np.float32); cnt = np.zeros(len(objs), np.int32)
    objects_to_arrays(objs, P + 2, out=(foot, cnt))
    assert np.array_equal(foot[:, :P], f2) and not foot[:, P:].any() and np.array_equal(cnt, c2)
    with pytest.raises(ValueError):
        objects_to_arrays(objs, P - 1)


def test_packed_record_round_trip():
    P, n = 5, 9
    rng = np.random.default_rng(0)
    cols = {name: rng.integers(0, 100, (n,) + shape).astype(dt) for name, dt, shape in record_layout(P)}
    buf = np.concatenate([np.ascontiguousarray(cols[name]).reshape(n, -1).view(np.uint8) for name, _, _ in record_layout(P)], axis=1)
    assert buf.shape == (n, record_bytes(P))
    back = unpack_record(np.concatenate([buf, np.zeros((3, buf.shape[1]), np.uint8)]), n, P)
    for name, dt, shape in record_layout(P):
        assert back[name].dtype == dt and np.array_equal(back[name], cols[name])


def test_records_are_untracked_by_the_cyclic_collector():
    """The assembler's containers are acyclic plain data and leave the collector's lists (a generation-2 pass over the
    records of a full match costs seconds); a dict the caller extends with a collectable value tracks itself again."""
    import gc

    import numpy as np

    from eagle_b200.coordinate_model import assemble_frames
    F, P = 3, 2
    xy = np.arange(F * 57 * 2, dtype=np.int32).reshape(F, 57, 2); order = np.tile(np.arange(64, dtype=np.uint8), (F, 1))
    count = np.full((F, 2), 6, np.int32); inl = np.full(F, (1 << 40) - 1, np.int64); status = np.zeros(F, np.int32)
    att = np.array([1, 0, 1], np.uint8); hi = np.array([0, 0, -1], np.int32)
    ci = np.ones((F, P, 2), np.int64); ib = np.array([[1, 0]] * F, np.uint8); bd = np.array([[1.0, 2.0, 3.0, 4.0]] * F)
    objs = [{"Player": {1: {"BBox": [1, 2, 3, 4], "Confidence": 0.5, "Bottom_center": [2, 4]}}, "Goalkeeper": {},
             "Ball": {7: {"BBox": (5, 6, 7, 8), "Confidence": 0.25, "Bottom_center": [6, 8]}}} for _ in range(F)]
    res = assemble_frames(objs, 25, 0, xy, order, count, None, inl, status, att, hi, ci, ib, bd)
    for r in res.values():
        parts = [r, r["Coordinates"], r["Keypoints"], r["Boundaries"], *r["Coordinates"].values(), *r["Keypoints"].values()]
        parts += [o for c in r["Coordinates"].values() for o in c.values()] + [o["BBox"] for c in r["Coordinates"].values() for o in c.values()]
        parts += [b for b in r["Boundaries"] if b is not None]
        assert not any(gc.is_tracked(p) for p in parts)
    assert gc.is_tracked(res)
    r = res[0]
    r["Coordinates"]["Extra"] = {"k": []}
    assert gc.is_tracked(r["Coordinates"])
    assert objs[0]["Ball"][7]["Bottom_center"] == [6, 8] and gc.is_tracked(objs[0]["Ball"][7]["Bottom_center"])   # the caller's own lists are left alone
