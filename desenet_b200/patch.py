"""Drop-in installation into an imported copy of the reference (splwany/DeSeNet).

The reference has no plugin layer: callers import the hot-path functions BY NAME
(``from core.utils.general import non_max_suppression`` in scripts/val.py:30-35, scripts/detect.py:27-31,
core/models/common.py:22-24; ``segoutput_to_target`` in val.py:38 / detect.py:32; ``batch_pix_accuracy`` /
``batch_intersection_union`` in val.py:36-37; ``seg_validation`` via ``val.`` in train.py:402), so a
replacement has to be patched into every importing module, not only the defining one.  ``install()`` does
that and returns a handle whose ``uninstall()`` restores the originals.

Two levels:
  * ``install()``                     — function level (strict): same tensors in, same tensors out.
  * ``install(model=model, fused=True)`` — additionally swaps ``Detect.forward``'s post-conv part and the
    seg head's final ``nn.Upsample`` for deferred versions, so NMS / mask / metrics read the RAW head outputs
    (nothing of z, x_i or the full-res logits is materialised unless somebody asks for it).
"""
from __future__ import annotations

import sys
from typing import Iterable, Optional

import torch
import torch.nn as nn

from . import api

_FUNCS = {
    "non_max_suppression": api.non_max_suppression,
    "segoutput_to_target": api.segoutput_to_target,
    "batch_pix_accuracy": api.batch_pix_accuracy,
    "batch_intersection_union": api.batch_intersection_union,
    "seg_validation": api.seg_validation,
}
# scale_coords / clip_coords are also called on CPU tensors and numpy arrays by the reference (plots, dataset
# code), so they are patched with a dispatcher that only takes CUDA float32 tensors

# modules of the reference that define or import those names
_MODULES = ("core.utils.general", "core.utils.plots", "core.utils.metrics", "core.models.common",
            "scripts.val", "scripts.detect", "scripts.show_seg_mIoU", "scripts.train", "val", "detect")


class _LazyPermuted(list):
    """The second element of Detect.forward's eval return value (yolo.py:277): the permuted raw logits,
    materialised on first access (only the loss computation in val.py:223-224 ever reads them)."""

    def __init__(self, head_outputs: api.HeadOutputs):
        super().__init__()
        self._ho = head_outputs
        self._done = False

    def _fill(self):
        if not self._done:
            super().extend(self._ho.x)
            self._done = True

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __getitem__(self, i):
        self._fill()
        return super().__getitem__(i)

    def __len__(self):
        return self._ho.head.nl


class DeferredUpsample(nn.Module):
    """Replacement for the seg head's final ``nn.Upsample(scale_factor=8, mode='bilinear',
    align_corners=True)`` (core/models/yolo.py:183).  eval + fused: returns ``SegLogits`` (the kernels upsample
    on the fly); otherwise materialises the full-res logits with the bit-exact upsample kernel on CUDA inputs
    and falls back to the original module for CPU tensors / training."""

    def __init__(self, original: nn.Module, fused: bool = True):
        super().__init__()
        self.original = original
        self.fused = fused
        sf = original.scale_factor
        self.scale = int(sf if not isinstance(sf, (tuple, list)) else sf[0])

    def forward(self, x):
        if self.training or not x.is_cuda:
            return self.original(x)
        if self.fused:
            return api.SegLogits(x.float() if x.dtype != torch.float32 else x, scale_factor=self.scale)
        return api.SegLogits(x.float(), scale_factor=self.scale).full().to(x.dtype)


def _detect_forward_factory(original_forward, fused: bool):
    def forward(self, x):
        if self.training or not x[0].is_cuda:
            return original_forward(self, x)
        raw = [self.m[i](x[i]) for i in range(self.nl)]                      # yolo.py:258
        raw = [r.float() if r.dtype != torch.float32 else r for r in raw]
        ho = api.HeadOutputs(raw, self.anchor_grid.reshape(self.nl, -1, 2), self.stride)
        if fused:
            return ho, _LazyPermuted(ho)
        return ho.z, ho.x                                                    # yolo.py:277
    return forward


class Installed:
    def __init__(self):
        self._undo = []

    def _set(self, obj, name, value):
        old = getattr(obj, name)
        setattr(obj, name, value)
        self._undo.append((obj, name, old))

    def uninstall(self):
        for obj, name, old in reversed(self._undo):
            setattr(obj, name, old)
        self._undo.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.uninstall()


def install(model: Optional[nn.Module] = None, fused: bool = True, modules: Optional[Iterable[str]] = None) -> Installed:
    """Patch every already-imported reference module that holds one of the hot-path callables; with
    ``model`` also defer the Detect decode and the seg head's final Upsample of that model instance."""
    h = Installed()
    for name in (modules or _MODULES):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for fn, repl in _FUNCS.items():
            if hasattr(mod, fn):
                h._set(mod, fn, repl)
    # scale_coords: CUDA float32 tensors go to the kernel, everything else (CPU tensors, numpy) to the original
    for name in (modules or _MODULES):
        mod = sys.modules.get(name)
        orig = getattr(mod, "scale_coords", None) if mod is not None else None
        if orig is None or getattr(orig, "_dsn_dispatch", False):
            continue

        def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, _orig=orig):
            if isinstance(coords, torch.Tensor) and coords.is_cuda and coords.dtype == torch.float32:
                return api.scale_coords(img1_shape, coords, img0_shape, ratio_pad)
            return _orig(img1_shape, coords, img0_shape, ratio_pad)

        scale_coords._dsn_dispatch = True
        h._set(mod, "scale_coords", scale_coords)
    # detection metrics: CUDA tensors go to the kernels; ap_per_class takes the device path whenever
    # a GPU is present and no plot is requested (plotting stays with the reference)
    def _cuda_f32(*ts):
        return all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in ts)

    for name in (modules or _MODULES):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for fn in ("box_iou", "process_batch", "ap_per_class"):
            orig = getattr(mod, fn, None)
            if orig is None or getattr(orig, "_dsn_dispatch", False):
                continue
            if fn == "box_iou":
                def repl(box1, box2, _orig=orig):
                    return api.box_iou(box1, box2) if _cuda_f32(box1, box2) else _orig(box1, box2)
            elif fn == "process_batch":
                def repl(detections, labels, iouv, _orig=orig):
                    return api.process_batch(detections, labels, iouv) if _cuda_f32(detections, labels, iouv) else _orig(detections, labels, iouv)
            else:
                def repl(tp, conf, pred_cls, target_cls, plot=False, save_dir='.', names=(), _orig=orig):
                    if plot or not torch.cuda.is_available():
                        return _orig(tp, conf, pred_cls, target_cls, plot=plot, save_dir=save_dir, names=names)
                    return api.ap_per_class(tp, conf, pred_cls, target_cls)
            repl._dsn_dispatch = True
            repl.__name__ = fn
            h._set(mod, fn, repl)
    if model is not None:
        patched_classes = set()
        for m in model.modules():
            cls = type(m)
            if cls.__name__ == "Detect" and hasattr(m, "anchor_grid") and hasattr(m, "m") and cls not in patched_classes:
                h._set(cls, "forward", _detect_forward_factory(cls.forward, fused))
                patched_classes.add(cls)
            out = getattr(m, "out", None)
            if (cls.__name__.startswith("SegMask") and isinstance(out, nn.Sequential) and len(out)
                    and isinstance(out[-1], nn.Upsample)):
                idx = len(out) - 1
                h._undo.append((_SeqSlot(out, idx), "module", out[idx]))
                out[idx] = DeferredUpsample(out[idx], fused)
    return h


class _SeqSlot:
    """``setattr(slot, 'module', m)`` puts ``m`` back at ``seq[idx]`` (used by Installed.uninstall)."""

    def __init__(self, seq, idx):
        object.__setattr__(self, "_seq", seq)
        object.__setattr__(self, "_idx", idx)

    def __setattr__(self, name, value):
        self._seq[self._idx] = value
