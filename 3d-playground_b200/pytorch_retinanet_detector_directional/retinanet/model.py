"""Drop-in for the post-processing of pytorch_retinanet_detector_directional/retinanet/model.py: batched_nms (:19-57),
nms, and PostProcess = everything ResNet.forward does after the heads (:306-397: default, LOCALIZE, MULTI_FRAME)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_pp = __import__("importlib").import_module(_core().__name__ + ".postprocess")
nms = _pp.nms
batched_nms = _pp.batched_nms
PostProcess = _pp.PostProcess3D
detect_per_class = _pp.detect_per_class
detect_multi_frame = _pp.detect_multi_frame
