"""Pitch-landmark tables for the geometry path, indexed by heatmap channel.

One row per landmark, in heatmap-channel order: (name, world_x_m, world_y_m, world_z_m) on the
UEFA 105 x 68 m pitch.  Same data as the reference's three dicts
(eagle/utils/pitch.py:1-60 names<->indices, :65 NOT_ON_PLANE, :209-267 world coordinates), folded
into a single channel-ordered table because the CUDA kernels index by channel, never by name.
tests/test_tables.py checks this table against the oracle's copy, the C header and (when
/root/reference is mounted) the reference module itself.
"""
from __future__ import annotations

import numpy as np

PITCH_LENGTH_M = 105  # x extent; reference PITCH_WIDTH (coordinate_model.py:18)
PITCH_WIDTH_M = 68    # y extent; reference PITCH_HEIGHT (coordinate_model.py:19)

NUM_LANDMARKS = 57

LANDMARKS = (
    ("L_GOAL_TL_POST", 0.0, 30.34, -2.44),  # 0
    ("L_GOAL_TR_POST", 0.0, 37.66, -2.44),  # 1
    ("L_GOAL_BL_POST", 0.0, 30.34, 0.0),  # 2
    ("L_GOAL_BR_POST", 0.0, 37.66, 0.0),  # 3
    ("L_GOAL_AREA_BR_CORNER", 5.5, 24.84, 0.0),  # 4
    ("L_GOAL_AREA_TR_CORNER", 5.5, 43.16, 0.0),  # 5
    ("L_GOAL_AREA_BL_CORNER", 0.0, 24.84, 0.0),  # 6
    ("L_GOAL_AREA_TL_CORNER", 0.0, 43.16, 0.0),  # 7
    ("L_PENALTY_AREA_BR_CORNER", 16.5, 13.84, 0.0),  # 8
    ("L_PENALTY_AREA_TR_CORNER", 16.5, 54.16, 0.0),  # 9
    ("L_PENALTY_AREA_BL_CORNER", 0.0, 13.84, 0.0),  # 10
    ("L_PENALTY_AREA_TL_CORNER", 0.0, 54.16, 0.0),  # 11
    ("BL_PITCH_CORNER", 0.0, 0.0, 0.0),  # 12
    ("TL_PITCH_CORNER", 0.0, 68.0, 0.0),  # 13
    ("B_TOUCH_AND_HALFWAY_LINES_INTERSECTION", 52.5, 0.0, 0.0),  # 14
    ("T_TOUCH_AND_HALFWAY_LINES_INTERSECTION", 52.5, 68.0, 0.0),  # 15
    ("R_PENALTY_AREA_BL_CORNER", 88.5, 13.84, 0.0),  # 16
    ("R_PENALTY_AREA_TL_CORNER", 88.5, 54.16, 0.0),  # 17
    ("R_PENALTY_AREA_BR_CORNER", 105.0, 13.84, 0.0),  # 18
    ("R_PENALTY_AREA_TR_CORNER", 105.0, 54.16, 0.0),  # 19
    ("R_GOAL_AREA_BL_CORNER", 99.5, 24.84, 0.0),  # 20
    ("R_GOAL_AREA_TL_CORNER", 99.5, 43.16, 0.0),  # 21
    ("R_GOAL_AREA_BR_CORNER", 105.0, 24.84, 0.0),  # 22
    ("R_GOAL_AREA_TR_CORNER", 105.0, 43.16, 0.0),  # 23
    ("R_GOAL_TL_POST", 105.0, 37.66, -2.44),  # 24
    ("R_GOAL_TR_POST", 105.0, 30.34, -2.44),  # 25
    ("R_GOAL_BL_POST", 105.0, 37.66, 0.0),  # 26
    ("R_GOAL_BR_POST", 105.0, 30.34, 0.0),  # 27
    ("BR_PITCH_CORNER", 105.0, 0.0, 0.0),  # 28
    ("TR_PITCH_CORNER", 105.0, 68.0, 0.0),  # 29
    ("CENTER_CIRCLE_TANGENT_TR", 61.31243189346428, 36.462426470588234, 0.0),  # 30
    ("CENTER_CIRCLE_TANGENT_TL", 43.68756810653572, 36.46242647058824, 0.0),  # 31
    ("CENTER_CIRCLE_TANGENT_BR", 61.31243189346428, 31.537573529411766, 0.0),  # 32
    ("CENTER_CIRCLE_TANGENT_BL", 43.68756810653572, 31.53757352941176, 0.0),  # 33
    ("CENTER_CIRCLE_TR", 58.97002704785691, 40.47002704785691, 0.0),  # 34
    ("CENTER_CIRCLE_TL", 46.02997295214309, 40.47002704785691, 0.0),  # 35
    ("CENTER_CIRCLE_BR", 58.97002704785691, 27.52997295214309, 0.0),  # 36
    ("CENTER_CIRCLE_BL", 46.02997295214309, 27.52997295214309, 0.0),  # 37
    ("CENTER_CIRCLE_R", 61.65, 34.0, 0.0),  # 38
    ("CENTER_CIRCLE_L", 43.35, 34.0, 0.0),  # 39
    ("T_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION", 52.5, 43.15, 0.0),  # 40
    ("B_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION", 52.5, 24.85, 0.0),  # 41
    ("CENTER_MARK", 52.5, 34.0, 0.0),  # 42
    ("LEFT_CIRCLE_R", 20.15, 34.0, 0.0),  # 43
    ("BL_16M_LINE_AND_PENALTY_ARC_INTERSECTION", 16.5, 26.687510683768487, 0.0),  # 44
    ("TL_16M_LINE_AND_PENALTY_ARC_INTERSECTION", 16.5, 41.31248931623151, 0.0),  # 45
    ("LEFT_CIRCLE_TANGENT_T", 19.9906727467215, 35.70008928040832, 0.0),  # 46
    ("LEFT_CIRCLE_TANGENT_B", 19.9906727467215, 32.29991071959168, 0.0),  # 47
    ("L_PENALTY_MARK", 11.0, 34.0, 0.0),  # 48
    ("L_MIDDLE_PENALTY", 16.5, 34.0, 0.0),  # 49
    ("RIGHT_CIRCLE_L", 84.85, 34.0, 0.0),  # 50
    ("BR_16M_LINE_AND_PENALTY_ARC_INTERSECTION", 88.5, 26.687510683768487, 0.0),  # 51
    ("TR_16M_LINE_AND_PENALTY_ARC_INTERSECTION", 88.5, 41.31248931623151, 0.0),  # 52
    ("RIGHT_CIRCLE_TANGENT_T", 85.0093272532785, 35.70008928040832, 0.0),  # 53
    ("RIGHT_CIRCLE_TANGENT_B", 85.0093272532785, 32.29991071959168, 0.0),  # 54
    ("R_PENALTY_MARK", 94.0, 34.0, 0.0),  # 55
    ("R_MIDDLE_PENALTY", 88.5, 34.0, 0.0),  # 56
)
assert len(LANDMARKS) == NUM_LANDMARKS

LANDMARK_NAMES = tuple(r[0] for r in LANDMARKS)
LANDMARK_INDEX = {n: i for i, n in enumerate(LANDMARK_NAMES)}

#: (57, 3) float64 world coordinates in metres.
WORLD_XYZ = np.array([r[1:] for r in LANDMARKS], dtype=np.float64)
#: (57, 2) float32 — what the reference hands to cv2.findHomography (coordinate_model.py:349).
WORLD_XY_F32 = WORLD_XYZ[:, :2].astype(np.float32)

#: Channels whose landmark is off the ground plane (cross-bar ends): reference NOT_ON_PLANE.
OFF_PLANE = tuple(int(i) for i in np.nonzero(WORLD_XYZ[:, 2] != 0.0)[0])
assert OFF_PLANE == (0, 1, 24, 25)

#: bit i set <=> landmark i may be used for the homography (coordinate_model.py:338-344).
ON_PLANE_MASK = sum(1 << i for i in range(NUM_LANDMARKS) if i not in OFF_PLANE)

#: Iteration order of the reference's GROUND_TRUTH_POINTS dict literal (pitch.py:209-267) as channel
#: indices.  _build_pitch_groups (coordinate_model.py:76-94) walks that dict, so the order of the
#: world-x / world-y line families used by the keypoint synthesis follows it.
REFERENCE_DICT_ORDER = (42, 13, 12, 29, 28, 48, 55, 11, 9, 10, 8, 17, 19, 16, 18, 7, 5, 6, 4, 21, 23, 20,
                        22, 0, 1, 2, 3, 24, 25, 26, 27, 15, 14, 40, 41, 45, 44, 52, 51, 30, 31, 32, 33, 34,
                        35, 36, 37, 38, 39, 43, 50, 46, 47, 49, 53, 54, 56)


def line_families():
    """World-y and world-x landmark families and their crossing labels, as the reference builds
    them (coordinate_model.py:76-94): keys rounded to 2 decimals, on-plane landmarks only, members
    and families in dict order, first label wins a coordinate.

    Returns (y_families, x_families, cross) where each family is a tuple of channels and
    cross[iy][ix] is the channel of the landmark at that crossing or -1.
    """
    coord_to_label, xg, yg = {}, {}, {}
    for ch in REFERENCE_DICT_ORDER:
        x, y, z = WORLD_XYZ[ch]
        if z != 0.0:
            continue
        xr, yr = round(float(x), 2), round(float(y), 2)
        coord_to_label.setdefault((xr, yr), ch)
        xg.setdefault(xr, []).append(ch)
        yg.setdefault(yr, []).append(ch)
    ykeys, xkeys = list(yg), list(xg)
    cross = [[coord_to_label.get((xk, yk), -1) for xk in xkeys] for yk in ykeys]
    return [tuple(yg[k]) for k in ykeys], [tuple(xg[k]) for k in xkeys], cross
