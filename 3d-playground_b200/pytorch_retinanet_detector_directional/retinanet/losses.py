"""Drop-in for pytorch_retinanet_detector_directional/retinanet/losses.py: calc_iou (:5-22) and the 3-output
FocalLoss (:24-362: focal classification, 20-d corner smooth-L1, direction-cosine "vp" loss)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_impl = __import__("importlib").import_module(_core().__name__ + ".losses_impl")
calc_iou = _impl.calc_iou
FocalLoss = _impl.FocalLoss
