"""Drop-in for homography.py: Homography (:156-748, transform methods) and Homography_Wrapper (:793-901)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _dropin import core as _core  # noqa: E402
_sys.path.pop(0)

_h = __import__("importlib").import_module(_core().__name__ + ".homography_impl")
Homography = _h.Homography
Homography_Wrapper = _h.Homography_Wrapper
