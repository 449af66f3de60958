"""TEST INFRASTRUCTURE ONLY — ctypes binding of the plain-C oracle (oracle/c/oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_nms_greedy.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, np.float32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    k = lib().orc_nms_greedy(_p(boxes), _p(scores), ctypes.c_int64(n), ctypes.c_double(thr), _p(keep))
    return keep[:k].copy()


def upsample_bilinear_ac(logits: np.ndarray, H: int, W: int) -> np.ndarray:
    v = np.ascontiguousarray(logits, np.float32)
    h, w = v.shape[-2:]
    planes = int(np.prod(v.shape[:-2]))
    out = np.empty(v.shape[:-2] + (H, W), np.float32)
    lib().orc_upsample_bilinear_ac(_p(v), ctypes.c_int64(planes), h, w, H, W, _p(out))
    return out


def argmax_first(logits: np.ndarray) -> np.ndarray:
    v = np.ascontiguousarray(logits, np.float32)
    bs, C = v.shape[:2]
    hw = int(np.prod(v.shape[2:]))
    out = np.empty((bs,) + v.shape[2:], np.int32)
    lib().orc_argmax_first(_p(v), ctypes.c_int64(bs), C, ctypes.c_int64(hw), _p(out))
    return out


def confusion_matrix(pred: np.ndarray, target: np.ndarray, C: int, rows: int) -> np.ndarray:
    p = np.ascontiguousarray(pred, np.int32).reshape(-1)
    t = np.ascontiguousarray(target, np.int64).reshape(-1)
    cm = np.zeros((rows, C), np.int64)
    lib().orc_confusion_matrix(_p(p), _p(t), ctypes.c_int64(p.size), C, rows, _p(cm))
    return cm


def conv1x1_fma(x: np.ndarray, w: np.ndarray, bias) -> np.ndarray:
    """yolo.py:182 / :252,258 as one fp32 FMA per term in channel order (see oracle.c: orc_conv1x1_fma)."""
    x = np.ascontiguousarray(x, np.float32)
    bs, K = x.shape[:2]
    plane = int(np.prod(x.shape[2:]))
    w = np.ascontiguousarray(w, np.float32).reshape(-1, K)
    N = w.shape[0]
    b = None if bias is None else np.ascontiguousarray(bias, np.float32).reshape(N)
    out = np.empty((bs, N) + x.shape[2:], np.float32)
    lib().orc_conv1x1_fma(_p(x), _p(w), _p(b) if b is not None else None, ctypes.c_int64(bs), K, N, ctypes.c_int64(plane), _p(out))
    return out
