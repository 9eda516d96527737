"""Packing of the detector's per-frame dicts into the flat foot-point arrays the C ABI takes."""
from __future__ import annotations

import numpy as np


def objects_to_arrays(objs: list, max_points: int):
    """Pack detect_objects() dicts into the flat arrays the C ABI takes.

    Returns foot (F, P, 2) float32 (Bottom_center per object, reference iteration order: class
    dict order, then id order) and count (F,) int32.
    """
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
