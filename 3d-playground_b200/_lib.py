"""ctypes binding of libgeom3d.so — the C ABI declared in include/geom3d.h.

There is NO fallback: if the library is missing or fails to load, importing an op raises.  The library is built
in-tree by build.py (nvcc, sm_100a) so it travels with the repository snapshot.
"""
import ctypes
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgeom3d.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "geom3d.h")

G3D_OK = 0
G3D_ERR_INVALID = -1
G3D_ERR_CUDA = -2
G3D_ERR_UNSUPPORTED = -3
VARIANT_2D = 0
VARIANT_3D = 1
ASSIGN_IGNORE = -2
ASSIGN_NEGATIVE = -1
TUNE_KEYS = {"pdl": 1, "force_anchor_centric": 3}

_c_ptr = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_f64 = ctypes.c_double

# name -> (restype, argtypes); must list every function of include/geom3d.h (tests/test_abi.py checks it)
SIGNATURES = {
    "g3d_version": (ctypes.c_char_p, []),
    "g3d_last_error": (ctypes.c_char_p, []),
    "g3d_sm_count": (_int, [_int]),
    "g3d_calc_iou": (_int, [_c_ptr, _i64, _c_ptr, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_gt_prepare": (_int, [_c_ptr, _i64, _i64, _i64, _int, _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_assign": (_int, [_c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_focal_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "g3d_focal_loss_fwd": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _i64, _i64, _i64, _int,
                                  _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_focal_loss_fwd_bwd": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _i64, _i64, _i64, _int,
                                      _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64,
                                      _c_ptr, _c_ptr, _int, _int, _int, _c_ptr]),
    "g3d_focal_loss_bwd": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _i64, _i64, _i64, _int,
                                  _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr, _c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_set_tuning": (_int, [_int, _i64]),
    "g3d_combine_shard_stats": (_int, [_c_ptr, _i64, _i64, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_exchange_buffer_doubles": (_i64, [_i64]),
    "g3d_exchange_shard_stats": (_int, [_c_ptr, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_decode3d": (_int, [_c_ptr, _c_ptr, _i64, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_decode2d": (_int, [_c_ptr, _i64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _int, _f32, _f32, _c_ptr, _int, _c_ptr]),
    "g3d_clip_boxes": (_int, [_c_ptr, _i64, _i64, _f32, _f32, _int, _c_ptr]),
    "g3d_rowmax": (_int, [_c_ptr, _i64, _i64, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_ladder_workspace_bytes": (_i64, [_i64, _i64]),
    "g3d_threshold_ladder": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _c_ptr,
                                    _c_ptr, _i64, _int, _c_ptr]),
    "g3d_filter_compact": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_gather_candidates": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _i64, _c_ptr,
                                     _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_gather_candidates_decoded": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _c_ptr, _int, _c_ptr, _c_ptr,
                                             _int, _f32, _f32, _c_ptr, _c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _int,
                                             _c_ptr]),
    "g3d_exclusive_scan_i32": (_int, [_c_ptr, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_assemble_detections": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _c_ptr, _i64,
                                       _c_ptr, _int, _c_ptr, _c_ptr, _int, _f32, _f32, _c_ptr, _c_ptr, _c_ptr, _c_ptr,
                                       _i64, _int, _c_ptr]),
    "g3d_nms_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "g3d_nms_segmented": (_int, [_c_ptr, _i64, _i64, _c_ptr, _i64, _c_ptr, _i64, _i64, _f64, _int, _c_ptr, _c_ptr,
                                 _c_ptr, _i64, _int, _c_ptr]),
    "g3d_state_to_space": (_int, [_c_ptr, _i64, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_space_to_im": (_int, [_c_ptr, _int, _i64, _i64, _c_ptr, _i64, _c_ptr, _int, _int, _c_ptr, _int, _c_ptr]),
    "g3d_state_to_im": (_int, [_c_ptr, _i64, _i64, _c_ptr, _i64, _c_ptr, _int, _int, _int, _c_ptr, _int, _int, _c_ptr]),
    "g3d_im_to_space": (_int, [_c_ptr, _c_ptr, _int, _i64, _c_ptr, _i64, _c_ptr, _int, _int, _c_ptr, _int, _c_ptr]),
    "g3d_space_to_state": (_int, [_c_ptr, _int, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_im_to_state": (_int, [_c_ptr, _c_ptr, _int, _i64, _c_ptr, _i64, _c_ptr, _int, _int, _c_ptr, _int, _c_ptr]),
    "g3d_height_from_template": (_int, [_c_ptr, _int, _c_ptr, _int, _c_ptr, _int, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_im_to_state_refined": (_int, [_c_ptr, _c_ptr, _int, _i64, _c_ptr, _c_ptr, _i64, _c_ptr, _int, _int, _c_ptr,
                                       _c_ptr, _int, _c_ptr]),
    "g3d_state_footprint": (_int, [_c_ptr, _i64, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_corners_to_box": (_int, [_c_ptr, _int, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_pairwise_iou_f64": (_int, [_c_ptr, _i64, _c_ptr, _i64, _f64, _int, _c_ptr, _int, _c_ptr]),
    "g3d_md_iou": (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _int, _c_ptr]),
    "g3d_kf_predict": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _f64, _f64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_kf_update": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _c_ptr, _c_ptr, _c_ptr, _int, _c_ptr]),
    "g3d_detect_tail_workspace_bytes": (_i64, [_i64, _i64]),
    "g3d_detect_tail": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _int, _c_ptr, _c_ptr,
                               _int, _f32, _f32, _f64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr,
                               _c_ptr, _i64, _int, _c_ptr]),
    "g3d_detect_tail_short": (_int, [_c_ptr, _i64, _i64, _i64, _i64, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _int, _c_ptr, _c_ptr,
                               _int, _f32, _f32, _f64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr,
                               _c_ptr, _i64, _int, _c_ptr]),
    "g3d_cross_camera_pairs": (_int, [_c_ptr, _c_ptr, _i64, _f64, _c_ptr, _c_ptr, _c_ptr, _i64, _int, _c_ptr]),
    "g3d_generate_anchors": (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _c_ptr, _i64, _int, _c_ptr]),
}

_lock = threading.Lock()
_lib = None


class Geom3dError(RuntimeError):
    """Raised when a libgeom3d entry point reports an error (or the library is missing)."""


def header_symbols(path=HEADER_PATH):
    """Names of all functions declared in include/geom3d.h (used by the ABI tests)."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g3d_[a-z0-9_]+)\s*\(", text)))


def lib():
    """The loaded library (loads on first use).  Raises Geom3dError if it cannot be loaded - no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise Geom3dError(
                f"{LIB_PATH} is missing: build it with `python 3d-playground_b200/build.py` "
                "(or __graft_entry__.build()); there is no CPU/PyTorch fallback for these ops")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise Geom3dError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def check(rc, what):
    """Translate a negative return code into a Python exception carrying the library's message."""
    if rc is not None and rc < 0:
        msg = lib().g3d_last_error()
        msg = msg.decode() if msg else ""
        kind = {G3D_ERR_INVALID: "invalid argument", G3D_ERR_CUDA: "CUDA error", G3D_ERR_UNSUPPORTED: "unsupported"}.get(rc, "error")
        raise Geom3dError(f"{what}: {kind} ({rc}): {msg}")
    return rc
