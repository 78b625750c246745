"""geom3d-b200: B200-native (sm_100a) 3D-box geometry hot path of DerekGloudemans/3D-playground.

The directory name is not a Python identifier; import it as `geom3d_b200` (a tiny alias package at the repository
root) or put this directory on sys.path and import the drop-in modules exactly as the reference scripts do:

    sys.path.insert(0, ".../3d-playground_b200")                                   # 2D copy + homography
    from retinanet import losses, utils, model ; import homography
    sys.path.insert(0, ".../3d-playground_b200/pytorch_retinanet_detector_directional")   # 3D directional copy
    from retinanet import losses, utils, model

Layout: csrc/ (CUDA kernels + C ABI, built into libgeom3d.so by build.py), _lib.py (ctypes binding), ops.py
(tensor-level wrappers), losses_impl.py / postprocess.py (shared host logic of the drop-in modules), dist.py
(multi-GPU sharding), and the drop-in module trees named above.
"""
from . import _lib  # noqa: F401
from ._lib import Geom3dError  # noqa: F401

__all__ = ["Geom3dError", "ops", "losses_impl", "postprocess", "homography_impl", "tracker_geometry", "dist"]
__version__ = "0.1"
