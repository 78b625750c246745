"""Drop-in for util_track/kf.py: Torch_KF (:14-428) with predict / update on the CUDA kernels of csrc/kf.cu."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

Torch_KF = __import__("importlib").import_module(_core().__name__ + ".kf_impl").Torch_KF
