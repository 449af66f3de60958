"""Tensor-level wrappers of the C-ABI kernels.  PyTorch is used for device memory and streams only;
all arithmetic happens in the hand-written sm_100a kernels behind ``include/desenet_b200.h``.
Every function enqueues on the current CUDA stream of the input's device and does not synchronise."""
from __future__ import annotations

from ctypes import c_float, c_int, c_uint64, c_void_p
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MAX_WH = 4096.0      # core/utils/general.py:675
MAX_NMS = 30000      # core/utils/general.py:676


def _require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"desenet_b200: {name} must live on a CUDA device (got {t.device}); "
                           "this package has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# element types the kernels load natively (include/desenet_b200.h: DSN_F32 / DSN_F16 / DSN_BF16); they compute in fp32
DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _require_cuda_input(t: torch.Tensor, name: str) -> torch.Tensor:
    """A raw network output: CUDA, fp32 / fp16 / bf16 (other floating types are upcast to fp32), contiguous."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"desenet_b200: {name} must live on a CUDA device (got {t.device}); "
                           "this package has no CPU path")
    if t.dtype not in DTYPE_CODE:
        if not t.is_floating_point():
            raise TypeError(f"{name} must be a floating-point tensor, got {t.dtype}")
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(t.data_ptr() if t is not None else 0)


def class_mask(classes, nc: int):
    """general.py:717-718 `classes` filter as the 4 x uint64 bitmask of the C ABI (None = all)."""
    if classes is None:
        return None
    if isinstance(classes, torch.Tensor):
        classes = classes.tolist()
    m = [0, 0, 0, 0]
    for c in classes:
        c = int(c)
        if 0 <= c < min(nc, 256):
            m[c >> 6] |= 1 << (c & 63)
    return (c_uint64 * 4)(*m)


class HeadSpec:
    """Host copy of a Detect head's geometry buffers (``anchor_grid`` in pixels, ``stride``), built once per head:
    on a real model both are CUDA buffers, and copying them to the host on every forward would cost two blocking
    device syncs per batch."""

    __slots__ = ("nl", "na", "anchors", "strides")

    def __init__(self, anchor_grid, stride):
        st = torch.as_tensor(stride, dtype=torch.float32).detach().cpu().reshape(-1)
        self.nl = int(st.numel())
        ag = torch.as_tensor(anchor_grid, dtype=torch.float32).detach().cpu().reshape(self.nl, -1, 2)
        self.na = int(ag.shape[1])
        self.anchors = ag
        self.strides = st


_SPEC_CACHE = {}


def head_spec(anchor_grid, stride) -> HeadSpec:
    """Cached ``HeadSpec``: keyed on (data_ptr, _version) of tensor arguments (an in-place update such as
    ``check_anchor_order``'s flip, autoanchor.py:11-19, bumps the version) or on the value of plain sequences."""
    if isinstance(anchor_grid, HeadSpec):
        return anchor_grid

    def key_of(v):
        if isinstance(v, torch.Tensor):
            return ("t", v.data_ptr(), v._version, tuple(v.shape), str(v.device))
        try:
            hash(v)
            return ("v", v)
        except TypeError:
            return None

    k1, k2 = key_of(anchor_grid), key_of(stride)
    if k1 is None or k2 is None:
        return HeadSpec(anchor_grid, stride)
    spec = _SPEC_CACHE.get((k1, k2))
    if spec is None:
        if len(_SPEC_CACHE) > 256:
            _SPEC_CACHE.clear()
        spec = _SPEC_CACHE[(k1, k2)] = HeadSpec(anchor_grid, stride)
    return spec


class Head:
    """Raw Detect head outputs + geometry (the arguments of the ``*_levels`` entry points).  ``anchor_grid`` may be a
    ``HeadSpec`` (then ``stride`` is ignored)."""

    def __init__(self, levels: Sequence[torch.Tensor], anchor_grid, stride=None):
        if len(levels) == 0:
            raise ValueError("no detection levels")
        self.levels: List[torch.Tensor] = [_require_cuda_input(l, f"levels[{i}]") for i, l in enumerate(levels)]
        if any(l.dtype != self.levels[0].dtype for l in self.levels):
            self.levels = [l.float() for l in self.levels]
        self.dtype = self.levels[0].dtype
        self.dtype_code = DTYPE_CODE[self.dtype]
        self.device = self.levels[0].device
        self.nl = len(self.levels)
        self.bs = int(self.levels[0].shape[0])
        spec = head_spec(anchor_grid, stride)
        self.spec = spec
        if spec.nl != self.nl:
            raise ValueError("one stride per level expected")
        ag, st = spec.anchors, spec.strides
        self.na = spec.na
        ch = int(self.levels[0].shape[1])
        if ch % self.na:
            raise ValueError(f"channels {ch} not divisible by na={self.na}")
        self.no = ch // self.na
        self.nc = self.no - 5
        for l in self.levels:
            if l.dim() != 4 or int(l.shape[0]) != self.bs or int(l.shape[1]) != ch or l.device != self.device:
                raise ValueError("levels must be (bs, na*no, ny, nx) tensors on one device")
        self.ny = [int(l.shape[2]) for l in self.levels]
        self.nx = [int(l.shape[3]) for l in self.levels]
        self.A = self.na * sum(y * x for y, x in zip(self.ny, self.nx))
        self.anchor_grid = ag
        self.stride = st
        # ctypes views (host arrays of the C ABI)
        self.c_raw = (c_void_p * self.nl)(*[l.data_ptr() for l in self.levels])
        self.c_ny = (c_int * self.nl)(*self.ny)
        self.c_nx = (c_int * self.nl)(*self.nx)
        self.c_stride = (c_float * self.nl)(*st.tolist())
        self.c_anchor = (c_float * (self.nl * self.na * 2))(*ag.reshape(-1).tolist())


def detect_decode(head: Head, want_z: bool = True, want_x: bool = True):
    """Detect.forward eval branch after the convs (core/models/yolo.py:259-277): ``(z, [x_i])``."""
    lib = _lib.load()
    with torch.cuda.device(head.device):
        z = torch.empty((head.bs, head.A, head.no), dtype=torch.float32, device=head.device) if want_z else None
        xs = [torch.empty((head.bs, head.na, ny, nx, head.no), dtype=torch.float32, device=head.device)
              for ny, nx in zip(head.ny, head.nx)] if want_x else None
        c_x = (c_void_p * head.nl)(*[x.data_ptr() for x in xs]) if xs is not None else None
        rc = lib.dsn_detect_decode(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.c_stride, head.c_anchor, head.nl, head.bs,
                                   head.na, head.no, _ptr(z), c_x, _stream(head.device))
    _lib.check(rc, "dsn_detect_decode")
    return z, xs


class ConfCut:
    """Workspace + result of ``conf_cut_levels``: per-image confidence-key bound and the filter's dropped flags."""

    def __init__(self, bs: int, device):
        self.cut_hi = torch.empty((bs,), dtype=torch.int32, device=device)
        self.dropped = torch.zeros((bs,), dtype=torch.int32, device=device)


def cut_target_for(max_det: int) -> int:
    """Candidates the greedy NMS is expected to consume before ``max_det`` survive (it stops there, general.py:735-736):
    an order of magnitude more than the survivors, at least 4096.  The confidence cut keeps 1.5 x as many; an image that
    needs more is redone uncut on the device (``detect_nms_redo_levels``), so the choice affects speed only."""
    return min(MAX_NMS, max(4096, 8 * int(max_det)))


def conf_cut_levels(head: Head, conf_thres: float, multi_label: bool, classes=None, target: int = 4096,
                    cut: Optional[ConfCut] = None) -> Optional[ConfCut]:
    """Sampled per-image confidence cut for huge candidate sets (see include/desenet_b200.h).  Returns None when
    the head layout does not allow it (the filter then simply runs without a cut)."""
    lib = _lib.load()
    ml = bool(multi_label) and head.nc > 1
    with torch.cuda.device(head.device):
        if cut is None:
            cut = ConfCut(head.bs, head.device)
        rc = lib.dsn_conf_cut_levels(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.nl, head.bs, head.na, head.no, float(conf_thres),
                                     int(ml), class_mask(classes, head.nc), int(target), None, _ptr(cut.cut_hi),
                                     _ptr(cut.dropped), _stream(head.device))
    if rc == -2:
        return None
    _lib.check(rc, "dsn_conf_cut_levels")
    return cut


def filter_levels(head: Head, conf_thres: float, multi_label: bool, classes=None, cap: Optional[int] = None,
                  keys: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None, cut: Optional[ConfCut] = None):
    """Fused decode + candidate filter (yolo.py:259-277 + general.py:668-718) -> (keys, counts)."""
    lib = _lib.load()
    ml = bool(multi_label) and head.nc > 1
    if cap is None:
        cap = head.A * (head.nc if ml else 1)
    with torch.cuda.device(head.device):
        if keys is None:
            keys = torch.empty((head.bs, cap), dtype=torch.int64, device=head.device)
        if counts is None:
            counts = torch.empty((head.bs,), dtype=torch.int32, device=head.device)
        rc = lib.dsn_filter_levels_cut(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.nl, head.bs, head.na, head.no,
                                       float(conf_thres), int(ml), class_mask(classes, head.nc), _ptr(keys), cap,
                                       _ptr(counts), _ptr(cut.cut_hi if cut else None), _stream(head.device))
    _lib.check(rc, "dsn_filter_levels_cut")
    return keys, counts


def filter_z(z: torch.Tensor, conf_thres: float, multi_label: bool, classes=None, cap: Optional[int] = None):
    """Candidate filter on a decoded prediction (bs, A, no) (general.py:668-718) -> (keys, counts)."""
    lib = _lib.load()
    z = _require_cuda_input(z, "prediction")
    bs, A, no = (int(v) for v in z.shape)
    nc = no - 5
    ml = bool(multi_label) and nc > 1
    if cap is None:
        cap = max(A * (nc if ml else 1), 1)
    with torch.cuda.device(z.device):
        keys = torch.empty((bs, cap), dtype=torch.int64, device=z.device)
        counts = torch.empty((bs,), dtype=torch.int32, device=z.device)
        rc = lib.dsn_filter_z(_ptr(z), DTYPE_CODE[z.dtype], bs, A, no, float(conf_thres), int(ml), class_mask(classes, nc), _ptr(keys), cap,
                              _ptr(counts), _stream(z.device))
    _lib.check(rc, "dsn_filter_z")
    return keys, counts


def pad_labels(labels, bs: int, device):
    """general.py:690-696 a-priori labels (one (n_i, 5) [cls, x, y, w, h] array per image) as the padded device arrays of
    the C ABI: ``(labels (bs, L, 5) fp32, label_count (bs) int32)``, or None when there is no label at all."""
    L = max((len(l) for l in labels), default=0)
    if L == 0:
        return None
    out = torch.zeros((bs, L, 5), dtype=torch.float32)
    cnt = torch.zeros((bs,), dtype=torch.int32)
    for i, l in enumerate(labels[:bs]):
        if len(l):
            t = torch.as_tensor(l, dtype=torch.float32).detach().cpu().reshape(-1, 5)
            out[i, :t.shape[0]] = t
            cnt[i] = t.shape[0]
    return out.to(device), cnt.to(device)


def append_label_keys(head: Head, padded, conf_thres: float, classes, keys: torch.Tensor, counts: torch.Tensor):
    """Add the label candidates (see ``pad_labels``) to the keys ``filter_levels`` produced (same stream, after it)."""
    lib = _lib.load()
    labels, label_count = padded
    with torch.cuda.device(head.device):
        rc = lib.dsn_append_label_keys(_ptr(labels), _ptr(label_count), head.bs, int(labels.shape[1]), head.A, head.nc,
                                       float(conf_thres), class_mask(classes, head.nc), _ptr(keys), int(keys.shape[1]),
                                       _ptr(counts), _stream(head.device))
    _lib.check(rc, "dsn_append_label_keys")


def nms_levels(head: Head, keys: torch.Tensor, counts: torch.Tensor, conf_thres: float, iou_thres: float,
               agnostic: bool, max_det: int, max_nms: int = MAX_NMS, max_wh: float = MAX_WH,
               det: Optional[torch.Tensor] = None, det_count: Optional[torch.Tensor] = None, cut: Optional[ConfCut] = None,
               labels=None, cta_threads: int = 0):
    """``cta_threads``: 0 = the library's rule (256-thread CTAs without a confidence cut, 512 with one), or 256 / 512."""
    lib = _lib.load()
    with torch.cuda.device(head.device):
        if det is None:
            det = torch.empty((head.bs, max(max_det, 0), 6), dtype=torch.float32, device=head.device)
        if det_count is None:
            det_count = torch.empty((head.bs,), dtype=torch.int32, device=head.device)
        if labels is not None:          # rows >= A are the label rows of ``pad_labels`` (no confidence cut on this path)
            assert cut is None
            rc = lib.dsn_nms_levels_labels(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.c_stride, head.c_anchor, head.nl, head.bs,
                                           head.na, head.no, _ptr(keys), int(keys.shape[1]), _ptr(counts), float(conf_thres),
                                           float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det), float(max_wh),
                                           _ptr(labels[0]), int(labels[0].shape[1]), _ptr(det), _ptr(det_count),
                                           _stream(head.device))
            _lib.check(rc, "dsn_nms_levels_labels")
            return det, det_count
        rc = lib.dsn_nms_levels_cta(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.c_stride, head.c_anchor, head.nl, head.bs,
                                    head.na, head.no, _ptr(keys), int(keys.shape[1]), _ptr(counts), float(conf_thres),
                                    float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det), float(max_wh),
                                    _ptr(cut.cut_hi if cut else None), _ptr(cut.dropped if cut else None), _ptr(det), _ptr(det_count),
                                    int(cta_threads), _stream(head.device))
    _lib.check(rc, "dsn_nms_levels_cta")
    return det, det_count


def detect_nms_levels(head: Head, conf_thres: float, iou_thres: float, classes=None, agnostic: bool = False,
                      multi_label: bool = False, max_det: int = 300, max_nms: int = MAX_NMS, max_wh: float = MAX_WH,
                      keys: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                      det: Optional[torch.Tensor] = None, det_count: Optional[torch.Tensor] = None):
    """Fused decode + filter + NMS, one launch (yolo.py:259-277 + general.py:659-750).
    Returns ``(det (bs, max_det, 6), det_count (bs), counts (bs))``."""
    lib = _lib.load()
    ml = bool(multi_label) and head.nc > 1
    with torch.cuda.device(head.device):
        if keys is None:
            keys = torch.empty((head.bs, max(head.A * (head.nc if ml else 1), 1)), dtype=torch.int64, device=head.device)
        if counts is None:
            counts = torch.empty((head.bs,), dtype=torch.int32, device=head.device)
        if det is None:
            det = torch.empty((head.bs, max(max_det, 0), 6), dtype=torch.float32, device=head.device)
        if det_count is None:
            det_count = torch.empty((head.bs,), dtype=torch.int32, device=head.device)
        rc = lib.dsn_detect_nms_levels(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.c_stride, head.c_anchor, head.nl, head.bs,
                                       head.na, head.no, float(conf_thres), int(ml), class_mask(classes, head.nc),
                                       float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det), float(max_wh),
                                       _ptr(keys), int(keys.shape[1]), _ptr(counts), _ptr(det), _ptr(det_count),
                                       _stream(head.device))
    _lib.check(rc, "dsn_detect_nms_levels")
    return det, det_count, counts


def detect_nms_redo_levels(head: Head, conf_thres: float, iou_thres: float, classes, agnostic: bool, multi_label: bool,
                           max_det: int, keys: torch.Tensor, counts: torch.Tensor, det: torch.Tensor, det_count: torch.Tensor,
                           max_nms: int = MAX_NMS, max_wh: float = MAX_WH):
    """On-device redo of the confidence cut: images whose ``det_count`` is -2 after ``nms_levels`` are redone without the
    cut (one CTA per flagged image, all others exit at once).  No host round trip; capturable in a CUDA graph."""
    lib = _lib.load()
    ml = bool(multi_label) and head.nc > 1
    with torch.cuda.device(head.device):
        rc = lib.dsn_detect_nms_redo_levels(head.c_raw, head.dtype_code, head.c_ny, head.c_nx, head.c_stride, head.c_anchor, head.nl, head.bs,
                                            head.na, head.no, float(conf_thres), int(ml), class_mask(classes, head.nc),
                                            float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det), float(max_wh),
                                            _ptr(keys), int(keys.shape[1]), _ptr(counts), _ptr(det), _ptr(det_count),
                                            _stream(head.device))
    _lib.check(rc, "dsn_detect_nms_redo_levels")
    return det, det_count


def nms_z(z: torch.Tensor, keys: torch.Tensor, counts: torch.Tensor, conf_thres: float, iou_thres: float,
          agnostic: bool, max_det: int, max_nms: int = MAX_NMS, max_wh: float = MAX_WH):
    lib = _lib.load()
    z = _require_cuda_input(z, "prediction")
    bs, A, no = (int(v) for v in z.shape)
    with torch.cuda.device(z.device):
        det = torch.empty((bs, max(max_det, 0), 6), dtype=torch.float32, device=z.device)
        det_count = torch.empty((bs,), dtype=torch.int32, device=z.device)
        rc = lib.dsn_nms_z(_ptr(z), DTYPE_CODE[z.dtype], bs, A, no, _ptr(keys), int(keys.shape[1]), _ptr(counts), float(conf_thres),
                           float(iou_thres), int(bool(agnostic)), int(max_nms), int(max_det), float(max_wh),
                           _ptr(det), _ptr(det_count), _stream(z.device))
    _lib.check(rc, "dsn_nms_z")
    return det, det_count


def cm_rows_for(n_channels: int, nclass: int) -> int:
    """Rows of the target-clamped confusion matrix: exact target values 0..max(nclass, C-1) plus one
    overflow row for everything larger (labeled, but never a class nor equal to a prediction)."""
    return max(int(nclass), int(n_channels) - 1) + 2


def _seg_call(fn_name: str, logits: torch.Tensor, lowres_hw, size, target, mask_dtype, mask_size, cm, cm_rows,
              mask_out=None):
    lib = _lib.load()
    logits = _require_cuda_input(logits, "logits")
    if logits.dim() != 4:
        raise ValueError("logits must be (bs, C, h, w)")
    bs, C = int(logits.shape[0]), int(logits.shape[1])
    H, W = int(size[0]), int(size[1])
    dev = logits.device
    with torch.cuda.device(dev):
        m32 = m8 = None
        mh, mw = (int(mask_size[0]), int(mask_size[1])) if mask_size is not None else (H, W)
        if mask_out is not None:
            if mask_out.dtype == torch.float32:
                m32 = mask_out
            elif mask_out.dtype == torch.uint8:
                m8 = mask_out
            else:
                raise TypeError("mask_out must be float32 or uint8")
            if tuple(mask_out.shape) != (bs, mh, mw) or not mask_out.is_contiguous() or mask_out.device != dev:
                raise ValueError("mask_out must be a contiguous (bs, mask_h, mask_w) tensor on the logits' device")
        elif mask_dtype is not None:
            if mask_dtype == torch.float32:
                m32 = torch.empty((bs, mh, mw), dtype=torch.float32, device=dev)
            elif mask_dtype == torch.uint8:
                m8 = torch.empty((bs, mh, mw), dtype=torch.uint8, device=dev)
            else:
                raise TypeError("mask dtype must be torch.float32 or torch.uint8")
        slots = 0
        if target is not None:
            target = _require_cuda(target, "target", torch.int64)
            if tuple(target.shape) != (bs, H, W):
                raise ValueError(f"target must be {(bs, H, W)}, got {tuple(target.shape)}")
            if cm is None:
                cm = torch.zeros((min(max(bs, 1), 64), cm_rows, C), dtype=torch.int64, device=dev)
            slots = int(cm.shape[0])
            if cm.dtype != torch.int64 or tuple(cm.shape[1:]) != (cm_rows, C) or not cm.is_contiguous():
                raise ValueError("cm must be a contiguous int64 (slots, cm_rows, C) tensor")
        else:
            cm = None
        if fn_name == "dsn_seg_fused":
            h, w = lowres_hw
            rc = lib.dsn_seg_fused(_ptr(logits), DTYPE_CODE[logits.dtype], bs, C, h, w, H, W, _ptr(target), _ptr(m32), _ptr(m8), mh, mw, _ptr(cm),
                                   slots, int(cm_rows), _stream(dev))
        else:
            rc = lib.dsn_seg_fullres(_ptr(logits), DTYPE_CODE[logits.dtype], bs, C, H, W, _ptr(target), _ptr(m32), _ptr(m8), mh, mw, _ptr(cm),
                                     slots, int(cm_rows), _stream(dev))
    _lib.check(rc, fn_name)
    return (m32 if m32 is not None else m8), cm


def seg_fused(logits_low: torch.Tensor, size: Tuple[int, int], target: Optional[torch.Tensor] = None,
              mask_dtype=None, mask_size=None, cm: Optional[torch.Tensor] = None, cm_rows: int = 4, mask_out=None):
    """Fused upsample(align_corners=True) + argmax (+ mask, + confusion matrix) from LOW-RES logits
    (yolo.py:183 + plots.py:222-229 + metrics.py:350-388).  Returns ``(mask | None, cm | None)``; cm is
    (slots, cm_rows, C) int64, accumulate with ``cm.sum(0)``."""
    return _seg_call("dsn_seg_fused", logits_low, (int(logits_low.shape[2]), int(logits_low.shape[3])), size, target,
                     mask_dtype, mask_size, cm, cm_rows, mask_out)


def seg_fullres(logits: torch.Tensor, target: Optional[torch.Tensor] = None, mask_dtype=None, mask_size=None,
                cm: Optional[torch.Tensor] = None, cm_rows: int = 4, mask_out=None):
    """argmax (+ mask, + confusion matrix) from FULL-RES logits (plots.py:222-229, metrics.py:350-388)."""
    return _seg_call("dsn_seg_fullres", logits, None, (int(logits.shape[2]), int(logits.shape[3])), target, mask_dtype,
                     mask_size, cm, cm_rows, mask_out)


def upsample_bilinear_ac(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """nn.Upsample(bilinear, align_corners=True) (yolo.py:183), bit-identical to ATen's CPU kernel."""
    lib = _lib.load()
    x = _require_cuda_input(x, "input")
    H, W = int(size[0]), int(size[1])
    lead = tuple(int(v) for v in x.shape[:-2])
    planes = 1
    for v in lead:
        planes *= v
    with torch.cuda.device(x.device):
        out = torch.empty(lead + (H, W), dtype=torch.float32, device=x.device)
        rc = lib.dsn_upsample_bilinear_ac(_ptr(x), DTYPE_CODE[x.dtype], planes, int(x.shape[-2]), int(x.shape[-1]), H, W, _ptr(out),
                                          _stream(x.device))
    _lib.check(rc, "dsn_upsample_bilinear_ac")
    return out


def seg_head_conv1x1(features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """The seg head's last conv, nn.Conv2d(c_hid, n_segcls, 1) + bias (yolo.py:182): (bs, K, h, w) features in the network's
    precision -> (bs, C, h, w) fp32 low-res logits (what ``seg_fused`` reads)."""
    lib = _lib.load()
    features = _require_cuda_input(features, "features")
    bs, K, h, w = (int(v) for v in features.shape)
    wt = weight.detach().reshape(int(weight.shape[0]), -1).to(device=features.device, dtype=torch.float32).contiguous()
    C = int(wt.shape[0])
    if int(wt.shape[1]) != K:
        raise ValueError(f"weight must be ({C}, {K}[, 1, 1]), got {tuple(weight.shape)}")
    bt = bias.detach().to(device=features.device, dtype=torch.float32).contiguous() if bias is not None else None
    with torch.cuda.device(features.device):
        out = torch.empty((bs, C, h, w), dtype=torch.float32, device=features.device)
        rc = lib.dsn_seg_head_conv1x1(_ptr(features), DTYPE_CODE[features.dtype], _ptr(wt), _ptr(bt), bs, K, C, h, w, _ptr(out),
                                      _stream(features.device))
    _lib.check(rc, "dsn_seg_head_conv1x1")
    return out


_CONV_PARAM_CACHE = {}


def _conv_params(weight: torch.Tensor, bias: Optional[torch.Tensor], dtype, device):
    """(N, K) weight in ``dtype`` and fp32 bias on ``device``; converted once per parameter version."""
    key = (weight.data_ptr(), weight._version, bias.data_ptr() if bias is not None else 0, bias._version if bias is not None else 0,
           dtype, str(device))
    hit = _CONV_PARAM_CACHE.get(id(weight))
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    wt = weight.detach().reshape(int(weight.shape[0]), -1).to(device=device, dtype=dtype).contiguous()
    bt = bias.detach().to(device=device, dtype=torch.float32).contiguous() if bias is not None else None
    if len(_CONV_PARAM_CACHE) > 64:
        _CONV_PARAM_CACHE.clear()
    _CONV_PARAM_CACHE[id(weight)] = (key, wt, bt)
    return wt, bt


def detect_head_conv1x1_supported(features: torch.Tensor, n_out: int) -> bool:
    """Whether ``detect_head_conv1x1`` takes this level (see include/desenet_b200.h: dsn_detect_head_conv1x1)."""
    if not features.is_cuda or features.dim() != 4 or features.dtype not in DTYPE_CODE or n_out > 256:
        return False
    K, plane = int(features.shape[1]), int(features.shape[2]) * int(features.shape[3])
    if features.dtype == torch.float32:
        return 36 * K * 4 <= 160 * 1024
    n_pad = (n_out + 15) // 16 * 16
    return K % 64 == 0 and plane % 8 == 0 and 3 * 16384 + 1024 + n_pad * K * 2 <= 200 * 1024


def detect_head_conv1x1(features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """One Detect output conv, nn.Conv2d(ch[i], na*no, 1) + bias (yolo.py:252,258): (bs, K, ny, nx) features ->
    (bs, na*no, ny, nx) raw level in the features' precision.  fp16 / bf16: tcgen05 tensor cores with fp32 accumulation;
    fp32: exact FMA chain."""
    lib = _lib.load()
    features = _require_cuda_input(features, "features")
    bs, K, ny, nx = (int(v) for v in features.shape)
    N = int(weight.shape[0])
    if int(weight.numel()) != N * K:
        raise ValueError(f"weight must be ({N}, {K}[, 1, 1]), got {tuple(weight.shape)}")
    wt, bt = _conv_params(weight, bias, features.dtype, features.device)
    with torch.cuda.device(features.device):
        out = torch.empty((bs, N, ny, nx), dtype=features.dtype, device=features.device)
        rc = lib.dsn_detect_head_conv1x1(_ptr(features), DTYPE_CODE[features.dtype], _ptr(wt), _ptr(bt), bs, K, N, ny, nx, _ptr(out),
                                         _stream(features.device))
    _lib.check(rc, "dsn_detect_head_conv1x1")
    return out


def letterbox_params(img1_shape, img0_shape, ratio_pad=None):
    """(pad_x, pad_y, gain, w0, h0) exactly as scale_coords derives them in Python (general.py:600-605)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return [float(pad[0]), float(pad[1]), float(gain), float(img0_shape[1]), float(img0_shape[0])]


def scale_coords_(boxes: torch.Tensor, params: torch.Tensor, counts: Optional[torch.Tensor] = None, do_round: bool = False):
    """In-place scale_coords + clip_coords (general.py:598-623) on ``boxes``: (n, >=4) rows of one image or
    (bs, rows, >=4) with per-image ``params`` (bs, 5) fp32 on the device (see ``letterbox_params``)."""
    lib = _lib.load()
    if not boxes.is_cuda or boxes.dtype != torch.float32:
        raise RuntimeError("desenet_b200: boxes must be a float32 CUDA tensor; this package has no CPU path")
    if boxes.stride(-1) != 1:
        raise ValueError("box rows must be contiguous in their last dimension")
    if boxes.dim() == 2:
        bs, rows, row_stride, image_stride = 1, int(boxes.shape[0]), int(boxes.stride(0)), 0
    elif boxes.dim() == 3:
        bs, rows, row_stride, image_stride = int(boxes.shape[0]), int(boxes.shape[1]), int(boxes.stride(1)), int(boxes.stride(0))
    else:
        raise ValueError("boxes must be (n, >=4) or (bs, rows, >=4)")
    params = _require_cuda(params, "params").reshape(-1, 5)
    if int(params.shape[0]) != bs:
        raise ValueError("one (pad_x, pad_y, gain, w0, h0) row per image expected")
    if rows == 0 or bs == 0:
        return boxes
    with torch.cuda.device(boxes.device):
        rc = lib.dsn_scale_coords(_ptr(boxes), row_stride, image_stride, _ptr(counts), bs, rows, _ptr(params),
                                  int(bool(do_round)), _stream(boxes.device))
    _lib.check(rc, "dsn_scale_coords")
    return boxes


def box_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """general.py:626-648: (N, M) IoU matrix of xyxy boxes (row-strided views such as ``labels[:, 1:]`` are fine)."""
    lib = _lib.load()
    for t, name in ((box1, "box1"), (box2, "box2")):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"desenet_b200: {name} must be a float32 CUDA tensor; this package has no CPU path")
        if t.dim() != 2 or t.shape[1] != 4 or (t.shape[0] > 0 and t.stride(1) != 1):
            raise ValueError(f"{name} must be (n, 4) with contiguous rows")
    n, m = int(box1.shape[0]), int(box2.shape[0])
    with torch.cuda.device(box1.device):
        out = torch.empty((n, m), dtype=torch.float32, device=box1.device)
        rc = lib.dsn_box_iou(_ptr(box1), int(box1.stride(0)) if n else 4, n, _ptr(box2), int(box2.stride(0)) if m else 4, m,
                             _ptr(out), _stream(box1.device))
    _lib.check(rc, "dsn_box_iou")
    return out


def process_batch(det: torch.Tensor, det_count: Optional[torch.Tensor], labels: torch.Tensor, label_off: torch.Tensor,
                  iouv: torch.Tensor, max_labels_per_image: int, correct: Optional[torch.Tensor] = None) -> torch.Tensor:
    """val.py:101-122 for a batch: ``det`` (bs, rows, >=6), ``labels`` (L, 5) [cls, xyxy] of all images back to back,
    ``label_off`` (bs+1) int32.  Returns ``correct`` (bs, rows, niou) uint8."""
    lib = _lib.load()
    det = det if det.dim() == 3 else det.unsqueeze(0)
    if not det.is_cuda or det.dtype != torch.float32 or det.stride(2) != 1:
        raise RuntimeError("desenet_b200: det must be a float32 CUDA tensor with contiguous rows; there is no CPU path")
    bs, rows = int(det.shape[0]), int(det.shape[1])
    labels = _require_cuda(labels.reshape(-1, 5), "labels")
    label_off = _require_cuda(label_off, "label_off", torch.int32)
    iouv = _require_cuda(iouv, "iouv")
    niou = int(iouv.numel())
    with torch.cuda.device(det.device):
        if correct is None:
            correct = torch.zeros((bs, rows, niou), dtype=torch.uint8, device=det.device)
        rc = lib.dsn_process_batch(_ptr(det), int(det.stride(0)), int(det.stride(1)), _ptr(det_count), bs, rows, _ptr(labels),
                                   _ptr(label_off), int(max_labels_per_image), _ptr(iouv), niou, _ptr(correct),
                                   _stream(det.device))
    _lib.check(rc, "dsn_process_batch")
    return correct


# ------------------------------------------------------------------------------------------------
# ap_per_class / compute_ap / detection ConfusionMatrix     core/utils/metrics.py:31-170
# ------------------------------------------------------------------------------------------------
_LINSPACE = {}


def _linspace_f64(num: int, device) -> torch.Tensor:
    """np.linspace(0, 1, num) exactly as numpy produces it (the grids metrics.py:56,114 interpolate on)."""
    key = (num, str(device))
    if key not in _LINSPACE:
        _LINSPACE[key] = torch.from_numpy(np.linspace(0, 1, num)).to(device)
    return _LINSPACE[key]


def ap_per_class_device(tp: torch.Tensor, conf: torch.Tensor, pred_cls: torch.Tensor, classes: torch.Tensor,
                        n_labels: torch.Tensor, npx: int = 1000):
    """metrics.py:47-79 on the device -> (ap (ncls, niou), p (ncls, npx), r (ncls, npx)) float64 CUDA tensors."""
    lib = _lib.load()
    dev = conf.device
    if not conf.is_cuda:
        raise RuntimeError("desenet_b200: ap_per_class_device needs CUDA tensors; this package has no CPU path")
    n = int(conf.shape[0])
    tp = tp.reshape(n, int(tp.shape[1]) if tp.dim() == 2 else 1)
    tp = (tp if tp.dtype in (torch.uint8, torch.bool) else (tp != 0)).contiguous().view(torch.uint8)
    niou = int(tp.shape[1])
    conf = _require_cuda(conf.reshape(-1), "conf")
    pred_cls = _require_cuda(pred_cls.reshape(-1), "pred_cls")
    classes = _require_cuda(classes.reshape(-1), "classes")
    n_labels = _require_cuda(n_labels.reshape(-1), "n_labels", torch.int64)
    ncls = int(classes.shape[0])
    with torch.cuda.device(dev):
        ap = torch.zeros((ncls, niou), dtype=torch.float64, device=dev)
        p = torch.zeros((ncls, npx), dtype=torch.float64, device=dev)
        r = torch.zeros((ncls, npx), dtype=torch.float64, device=dev)
        if ncls == 0:
            return ap, p, r
        nbytes = int(lib.dsn_ap_workspace_bytes(n, niou, ncls))
        if nbytes < 0:
            raise ValueError(f"ap_per_class: unsupported sizes n={n} niou={niou} ncls={ncls}")
        ws = torch.empty((nbytes + 255,), dtype=torch.uint8, device=dev)
        base = (ws.data_ptr() + 255) // 256 * 256
        rc = lib.dsn_ap_per_class(_ptr(tp), _ptr(conf), _ptr(pred_cls), n, niou, _ptr(classes), _ptr(n_labels), ncls,
                                  _ptr(_linspace_f64(101, dev)), _ptr(_linspace_f64(npx, dev)), npx, c_void_p(base), nbytes,
                                  _ptr(ap), _ptr(p), _ptr(r), _stream(dev))
        _lib.check(rc, "dsn_ap_per_class")
        ws.record_stream(torch.cuda.current_stream(dev))
    return ap, p, r


def compute_ap_device(recall: torch.Tensor, precision: torch.Tensor) -> torch.Tensor:
    """metrics.py:95-121 for one curve -> (1,) float64 CUDA tensor."""
    lib = _lib.load()
    recall = _require_cuda(recall.reshape(-1), "recall", torch.float64)
    precision = _require_cuda(precision.reshape(-1), "precision", torch.float64)
    dev = recall.device
    n = int(recall.shape[0])
    with torch.cuda.device(dev):
        env = torch.empty((max(n, 1),), dtype=torch.float64, device=dev)
        ap = torch.zeros((1,), dtype=torch.float64, device=dev)
        rc = lib.dsn_compute_ap(_ptr(recall), _ptr(precision), n, _ptr(_linspace_f64(101, dev)), _ptr(env), _ptr(ap), _stream(dev))
    _lib.check(rc, "dsn_compute_ap")
    return ap, env


def det_confusion_matrix_(matrix: torch.Tensor, det: torch.Tensor, det_count: Optional[torch.Tensor], labels: torch.Tensor,
                          label_off: torch.Tensor, max_labels_per_image: int, nc: int, conf: float = 0.25,
                          iou_thres: float = 0.45, skip_empty: bool = False) -> torch.Tensor:
    """metrics.py:132-170 for a batch, accumulating into ``matrix`` ((nc+1, nc+1) int64 CUDA tensor)."""
    lib = _lib.load()
    det = det if det.dim() == 3 else det.unsqueeze(0)
    if not det.is_cuda or det.dtype != torch.float32 or (det.numel() and det.stride(2) != 1):
        raise RuntimeError("desenet_b200: det must be a float32 CUDA tensor with contiguous rows; there is no CPU path")
    bs, rows = int(det.shape[0]), int(det.shape[1])
    labels = _require_cuda(labels.reshape(-1, 5), "labels")
    label_off = _require_cuda(label_off, "label_off", torch.int32)
    matrix = _require_cuda(matrix, "matrix", torch.int64)
    assert tuple(matrix.shape) == (nc + 1, nc + 1)
    with torch.cuda.device(det.device):
        rc = lib.dsn_det_confusion_matrix(_ptr(det), int(det.stride(0)), int(det.stride(1)) if rows else 6, _ptr(det_count), bs,
                                          rows, _ptr(labels), _ptr(label_off), int(max_labels_per_image), int(nc),
                                          float(np.float32(conf)), float(np.float32(iou_thres)), int(bool(skip_empty)), _ptr(matrix),
                                          _stream(det.device))
    _lib.check(rc, "dsn_det_confusion_matrix")
    return matrix


def labels_to_native(targets: torch.Tensor, width: float, height: float, params: torch.Tensor) -> torch.Tensor:
    """val.py:227,243-255 for a whole batch: ``targets`` (n, 6) [image, cls, normalised xywh] -> ``labelsn`` (n, 5)
    [cls, native-space xyxy], with per-image ``params`` (bs, 5) as for ``scale_coords_``."""
    lib = _lib.load()
    targets = _require_cuda(targets.reshape(-1, 6), "targets")
    params = _require_cuda(params, "params").reshape(-1, 5)
    n = int(targets.shape[0])
    with torch.cuda.device(targets.device):
        out = torch.empty((n, 5), dtype=torch.float32, device=targets.device)
        rc = lib.dsn_labels_to_native(_ptr(targets), n, float(width), float(height), _ptr(params), int(params.shape[0]), _ptr(out),
                                      _stream(targets.device))
    _lib.check(rc, "dsn_labels_to_native")
    return out


class GraphBurst:
    """A fixed sequence of captured steps (``torch.cuda.CUDAGraph`` objects) and the streams they run on, launched
    by one native call (``dsn_graph_launch_many``): the serving loop of a rank without per-step Python overhead."""

    def __init__(self, graphs: Sequence[torch.cuda.CUDAGraph], streams: Sequence[torch.cuda.Stream]):
        assert len(graphs) == len(streams)
        self.n = len(graphs)
        self._keep = (list(graphs), list(streams))
        self._execs = (c_void_p * self.n)(*[int(g.raw_cuda_graph_exec()) for g in graphs])
        self._streams = (c_void_p * self.n)(*[int(s.cuda_stream) for s in streams])

    def launch(self, n: Optional[int] = None):
        rc = _lib.load().dsn_graph_launch_many(self._execs, self._streams, self.n if n is None else int(n))
        _lib.check(rc, "dsn_graph_launch_many")
