"""Drop-in for pytorch_retinanet_detector_directional/retinanet/anchors.py (same generator as the 2D copy)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_a = __import__("importlib").import_module(_core().__name__ + ".anchors_impl")
Anchors = _a.Anchors
anchors_for_image = _a.anchors_for_image
