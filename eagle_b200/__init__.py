"""eagle_b200 — B200-native geometry path of nreHieW/Eagle (decode -> homography -> projection).

Importing the package does not load CUDA; `eagle_b200.CoordinateModel` / `GeometryEngine` do, and
raise if libeagle_b200.so has not been built (there is no CPU fallback)."""
__all__ = ["CoordinateModel", "GeometryPath", "GeometryEngine", "KeypointDecoder", "PropagatedPath"]


def __getattr__(name):
    if name in ("CoordinateModel", "GeometryPath"):
        from . import coordinate_model
        return getattr(coordinate_model, name)
    if name == "KeypointDecoder":
        from .keypoints import KeypointDecoder
        return KeypointDecoder
    if name == "PropagatedPath":
        from .propagation import PropagatedPath
        return PropagatedPath
    if name == "GeometryEngine":
        from .engine import GeometryEngine
        return GeometryEngine
    raise AttributeError(name)
