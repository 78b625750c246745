"""Drop-in for pytorch_retinanet_detector_directional/retinanet/utils.py: BBoxTransform 12 -> 20 (:82-149) and
ClipBoxes (:152-167)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_pp = __import__("importlib").import_module(_core().__name__ + ".postprocess")
BBoxTransform = _pp.BBoxTransform3D
ClipBoxes = _pp.ClipBoxes
