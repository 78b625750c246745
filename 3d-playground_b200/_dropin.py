"""Locates the core package for the drop-in module trees (which are imported under the reference's own module names,
e.g. `retinanet.losses`, with only their directory on sys.path)."""
import os
import sys


def core():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repository root (holds geom3d_b200/)
    if root not in sys.path:
        sys.path.insert(0, root)
    import geom3d_b200
    return geom3d_b200
