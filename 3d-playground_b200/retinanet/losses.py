"""Drop-in for retinanet/losses.py (2D copy): calc_iou (:5-22) and FocalLoss (:24-177) on the fused CUDA kernels."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_impl = __import__("importlib").import_module(_core().__name__ + ".losses_impl")
calc_iou = _impl.calc_iou
FocalLoss = _impl.FocalLoss
