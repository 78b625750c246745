"""Import alias for the `3d-playground_b200/` package directory (whose name is not a valid Python identifier).

`import geom3d_b200` gives the package living in ../3d-playground_b200: this module only redirects __path__.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "3d-playground_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _fh
