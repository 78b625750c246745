"""Drop-in for the inference tail of retinanet/model.py (2D copy, :270-311): `nms` and the post-processing module.
The ResNet/FPN backbone (dense convolutions -> cuDNN) is out of scope; see INTEGRATION.md for the three-line change
that makes the reference's ResNet.forward call PostProcess."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_pp = __import__("importlib").import_module(_core().__name__ + ".postprocess")
nms = _pp.nms
batched_nms = _pp.batched_nms
PostProcess = _pp.PostProcess2D
detect_per_class = _pp.detect_per_class
