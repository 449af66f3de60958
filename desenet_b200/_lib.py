"""ctypes binding of the C-ABI library (include/desenet_b200.h).

The product has NO CPU fallback: if the CUDA library is missing or cannot be loaded, importing
callers get a RuntimeError that says how to build it."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

from . import build as _build

_lib = None

# name -> (restype, argtypes); mirrors include/desenet_b200.h one to one
_FLOATPP = POINTER(c_void_p)
SIGNATURES = {
    "dsn_version": (c_int, []),
    "dsn_error_string": (c_char_p, [c_int]),
    "dsn_graph_launch_many": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int]),
    "dsn_detect_decode": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float),
                                  c_int, c_int, c_int, c_int, c_void_p, _FLOATPP, c_void_p]),
    "dsn_filter_levels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int, c_int, c_int, c_float, c_int,
                                  POINTER(c_uint64), c_void_p, c_int64, c_void_p, c_void_p]),
    "dsn_filter_z": (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_float, c_int, POINTER(c_uint64), c_void_p, c_int64,
                             c_void_p, c_void_p]),
    "dsn_nms_levels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float), c_int, c_int,
                               c_int, c_int, c_void_p, c_int64, c_void_p, c_float, c_double, c_int, c_int, c_int, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_nms_levels_cta": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float), c_int, c_int,
                                   c_int, c_int, c_void_p, c_int64, c_void_p, c_float, c_double, c_int, c_int, c_int, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dsn_append_label_keys": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_float, POINTER(c_uint64), c_void_p, c_int64,
                                      c_void_p, c_void_p]),
    "dsn_nms_levels_labels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float), c_int, c_int,
                                      c_int, c_int, c_void_p, c_int64, c_void_p, c_float, c_double, c_int, c_int, c_int, c_float,
                                      c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dsn_conf_cut_levels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int, c_int, c_int, c_float, c_int,
                                    POINTER(c_uint64), c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_filter_levels_cut": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int, c_int, c_int, c_float, c_int,
                                      POINTER(c_uint64), c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "dsn_detect_nms_levels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float), c_int,
                                      c_int, c_int, c_int, c_float, c_int, POINTER(c_uint64), c_double, c_int, c_int,
                                      c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_detect_nms_redo_levels": (c_int, [_FLOATPP, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), POINTER(c_float), c_int,
                                           c_int, c_int, c_int, c_float, c_int, POINTER(c_uint64), c_double, c_int, c_int,
                                           c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_nms_z": (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_int64, c_void_p, c_float, c_double, c_int, c_int,
                          c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "dsn_seg_fused": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                              c_int, c_void_p, c_int, c_int, c_void_p]),
    "dsn_seg_fullres": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_void_p, c_int, c_int, c_void_p]),
    "dsn_upsample_bilinear_ac": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dsn_box_iou": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "dsn_process_batch": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                  c_int, c_void_p, c_void_p]),
    "dsn_labels_to_native": (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "dsn_ap_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "dsn_ap_per_class": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_compute_ap": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dsn_det_confusion_matrix": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                         c_float, c_float, c_int, c_void_p, c_void_p]),
    "dsn_detect_head_conv1x1": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dsn_seg_head_conv1x1": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dsn_peer_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "dsn_peer_free": (c_int, [c_void_p]),
    "dsn_peer_export": (c_int, [c_void_p, c_char_p]),
    "dsn_peer_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "dsn_peer_close": (c_int, [c_void_p]),
    "dsn_peer_signal": (c_int, [c_void_p, c_void_p, c_void_p]),
    "dsn_peer_wait": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "dsn_scale_coords": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
}


class DesenetCudaError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc is present) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DSN_LIB") or _build.LIB
    if os.environ.get("DSN_LIB"):          # tuning builds (e.g. the DSN_TIMING .so)
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib
    try:
        path = _build.build()
    except Exception as e:  # no nvcc on this box: use the prebuilt .so that travelled with the repo
        if not os.path.exists(path):
            raise RuntimeError(
                f"desenet_b200: CUDA library {path} is missing and could not be built ({e}). "
                "Run `python -m desenet_b200.build` on a machine with nvcc; there is no CPU fallback.") from e
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = the .so is stale: rebuild
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dsn_error_string(rc)
        raise DesenetCudaError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
