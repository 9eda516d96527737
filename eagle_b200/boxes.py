"""Packing of the detector's per-frame dicts into the flat foot-point arrays the C ABI takes."""
from __future__ import annotations

import numpy as np


def objects_to_arrays_py(objs: list, max_points: int):
    """Readable statement of objects_to_arrays (the checker in tests/test_assemble.py)."""
    F = len(objs)
    foot = np.zeros((F, max_points, 2), np.float32)
    count = np.zeros(F, np.int32)
    for i, o in enumerate(objs):
        k = 0
        for cls in o:
            for _id, d in o[cls].items():
                if k >= max_points:
                    raise ValueError(f"frame {i}: more than {max_points} objects")
                foot[i, k] = d["Bottom_center"]
                k += 1
        count[i] = k
    return foot, count


def max_objects(objs: list) -> int:
    """Largest number of detections in one frame."""
    from . import _assemble
    return int(_assemble.max_objects(objs if type(objs) is list else list(objs)))


def objects_to_arrays(objs: list, max_points: int, out=None):
    """Pack detect_objects() dicts into the flat arrays the C ABI takes.

    Returns foot (F, P, 2) float32 (Bottom_center per object, reference iteration order: class
    dict order, then id order) and count (F,) int32.  ``out`` = (foot, count) numpy views to fill in place
    (e.g. page-locked staging memory).  The loop runs in the C extension (csrc/assemble.c: 2 us per frame)."""
    from . import _assemble
    objs = objs if type(objs) is list else list(objs)
    F = len(objs)
    if out is None:
        foot = np.zeros((F, max_points, 2), np.float32)
        count = np.zeros(F, np.int32)
    else:
        foot, count = out
        foot[...] = 0
    _assemble.pack_foot_points(objs, foot, count, int(max_points))
    return foot, count
