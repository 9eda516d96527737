"""CPU oracle for the per-frame geometry path.  TEST INFRASTRUCTURE ONLY.

A numpy/cv2 (plus one small C file) restatement of what the reference computes on the
decode -> homography -> projection path (eagle/models/keypoint_hrnet.py:575-595 and
eagle/models/coordinate_model.py:188-417), used as the parity checker for the CUDA kernels.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  Nothing under ``eagle_b200/`` imports it, and
the product path raises if its CUDA extension is missing rather than falling back to this code.

Pinning (how this oracle is tied to the real reference) -- see ``oracle/README.md``:
  * ``oracle/ref_harness.py`` executes the reference's own ``CoordinateModel.get_coordinates``
    (from /root/reference, in the authoring container) under stub third-party modules and
    records every ``cv2.findHomography`` / ``cv2.perspectiveTransform`` call; the recordings
    are committed as ``tests/golden/*.npz`` by ``oracle/make_golden.py``.
  * the restated OpenCV internals (``oracle/homography.py``) are checked against the live
    ``cv2`` of the image in ``tests/test_oracle_homography.py``.
"""
