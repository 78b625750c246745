"""Drop-in for retinanet/utils.py (2D copy): BBoxTransform (:82-126) and ClipBoxes (:129-144).
(BasicBlock / Bottleneck, :6-80, belong to the backbone and are out of scope.)"""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_pp = __import__("importlib").import_module(_core().__name__ + ".postprocess")
BBoxTransform = _pp.BBoxTransform2D
ClipBoxes = _pp.ClipBoxes
