"""Drop-in installation into an imported copy of the reference (splwany/DeSeNet).

The reference has no plugin layer: callers import the hot-path functions BY NAME
(``from core.utils.general import non_max_suppression`` in scripts/val.py:30-35, scripts/detect.py:27-31,
core/models/common.py:22-24; ``segoutput_to_target`` in val.py:38 / detect.py:32; ``batch_pix_accuracy`` /
``batch_intersection_union`` in val.py:36-37; ``seg_validation`` via ``val.`` in train.py:402), so a
replacement has to be patched into every importing module, not only the defining one.  ``install()`` does
that and returns a handle whose ``uninstall()`` restores the originals.

Every installed callable is a DISPATCHER: CUDA tensors (fp32, or fp16 / bf16 as ``--half`` produces them,
val.py:144,178-179) and the deferred head outputs (``HeadOutputs`` / ``SegLogits``) go to the sm_100a kernels;
anything else — CPU tensors, numpy arrays, a model living on the CPU — goes to the reference's ORIGINAL function, exactly
as before the patch (scripts/detect.py:198 hands ``seg_pred.cpu()`` to ``segoutput_to_target``; plots and dataset code
call ``scale_coords`` on numpy arrays).  The kernels themselves have no CPU path; the reference keeps its own.

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

from . import api, ops

# names of the hot-path callables and the reference module that DEFINES each; every other module
# holding the same function object imported it by name and is patched as well
_DEFINED_IN = {
    "non_max_suppression": ("core.utils.general",),
    "scale_coords": ("core.utils.general",),
    "box_iou": ("core.utils.general",),
    "segoutput_to_target": ("core.utils.plots",),
    "batch_pix_accuracy": ("core.utils.metrics",),
    "batch_intersection_union": ("core.utils.metrics",),
    "ap_per_class": ("core.utils.metrics",),
    "seg_validation": ("scripts.val", "val"),
    "process_batch": ("scripts.val", "val"),
}


def _is_cuda_tensor(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda and t.is_floating_point()


def _ours(x) -> bool:
    """True when the kernels take ``x``: a deferred head output or a floating-point CUDA tensor."""
    return isinstance(x, (api.HeadOutputs, api.SegLogits)) or _is_cuda_tensor(x)


def _make_dispatchers(orig):
    """name -> dispatcher; ``orig[name]`` is the reference's own function (the fall-through for non-CUDA inputs)."""

    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300):
        fn = api.non_max_suppression if _ours(prediction) else orig["non_max_suppression"]
        return fn(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)

    def segoutput_to_target(output, size=None):
        return (api.segoutput_to_target if _ours(output) else orig["segoutput_to_target"])(output, size)

    def batch_pix_accuracy(output, target):
        return (api.batch_pix_accuracy if _ours(output) else orig["batch_pix_accuracy"])(output, target)

    def batch_intersection_union(output, target, nclass):
        return (api.batch_intersection_union if _ours(output) else orig["batch_intersection_union"])(output, target, nclass)

    def seg_validation(model, n_segcls, valloader, half_precision=True):
        p = next(iter(model.parameters()), None) if hasattr(model, "parameters") else None
        on_cuda = p is not None and p.is_cuda
        return (api.seg_validation if on_cuda else orig["seg_validation"])(model, n_segcls, valloader, half_precision)

    def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
        if _is_cuda_tensor(coords) and coords.dtype == torch.float32:
            return api.scale_coords(img1_shape, coords, img0_shape, ratio_pad)
        return orig["scale_coords"](img1_shape, coords, img0_shape, ratio_pad)

    def _cuda_f32(*ts):
        return all(_is_cuda_tensor(t) and t.dtype == torch.float32 for t in ts)

    def box_iou(box1, box2):
        return api.box_iou(box1, box2) if _cuda_f32(box1, box2) else orig["box_iou"](box1, box2)

    def process_batch(detections, labels, iouv):
        if _cuda_f32(detections, labels, iouv):
            return api.process_batch(detections, labels, iouv)
        return orig["process_batch"](detections, labels, iouv)

    def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, save_dir='.', names=()):
        # plotting stays with the reference; without a GPU there is nothing to dispatch to
        if plot or not torch.cuda.is_available():
            return orig["ap_per_class"](tp, conf, pred_cls, target_cls, plot=plot, save_dir=save_dir, names=names)
        return api.ap_per_class(tp, conf, pred_cls, target_cls)

    out = dict(non_max_suppression=non_max_suppression, segoutput_to_target=segoutput_to_target,
               batch_pix_accuracy=batch_pix_accuracy, batch_intersection_union=batch_intersection_union,
               seg_validation=seg_validation, scale_coords=scale_coords, box_iou=box_iou, process_batch=process_batch,
               ap_per_class=ap_per_class)
    for name, fn in out.items():
        fn._dsn_dispatch = True
        fn._dsn_original = orig.get(name)
        fn.__doc__ = f"desenet_b200 dispatcher for the reference's `{name}`: CUDA inputs -> sm_100a kernels, anything else -> the original."
    return out


class _LazyPermuted(list):
    """The second element of Detect.forward's eval return value (yolo.py:277): the permuted raw logits,
    materialised on first access (only the loss computation in val.py:223-224 ever reads them)."""

    def __init__(self, head_outputs: api.HeadOutputs):
        super().__init__()
        self._ho = head_outputs
        self._done = False

    def _fill(self):
        if not self._done:
            xs = self._ho.x
            if self._ho.dtype != torch.float32:
                xs = [x.to(self._ho.dtype) for x in xs]
            super().extend(xs)
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
        self.training = original.training      # a module built after model.eval() would otherwise start in training mode
        self.out_dtype_of = None               # head_convs: a parameter whose dtype the strict mode's logits take
        sf = original.scale_factor
        self.scale = int(sf if not isinstance(sf, (tuple, list)) else sf[0])

    def forward(self, x):
        if self.training or not x.is_cuda:
            return self.original(x)
        if self.fused:
            return api.SegLogits(x, scale_factor=self.scale)
        return api.SegLogits(x, scale_factor=self.scale).full().to(x.dtype if self.out_dtype_of is None else self.out_dtype_of.dtype)


class FusedHeadConv(nn.Module):
    """Opt-in replacement (``install(..., head_convs=True)``) for the seg head's last ``nn.Conv2d(c_hid, n_segcls, 1)``
    (core/models/yolo.py:182): eval + CUDA -> the streaming 1x1 kernel, which reads the features in the network's
    precision and hands fp32 low-res logits to the deferred Upsample; anything else -> the original module.  SURVEY.md
    §8f rank 4 — it changes the boundary: the logits are no longer cuDNN's (summation order; under ``--half`` no fp16
    rounding), so parity with the reference's conv is a tolerance."""

    def __init__(self, original: nn.Conv2d):
        super().__init__()
        self.original = original
        self.training = original.training

    @staticmethod
    def eligible(m) -> bool:
        return (isinstance(m, nn.Conv2d) and tuple(m.kernel_size) == (1, 1) and tuple(m.stride) == (1, 1) and tuple(m.padding) == (0, 0)
                and tuple(m.dilation) == (1, 1) and m.groups == 1 and m.out_channels <= 4 and m.padding_mode == "zeros")

    def forward(self, x):
        if self.training or not x.is_cuda:
            return self.original(x)
        return ops.seg_head_conv1x1(x, self.original.weight, self.original.bias)


def _detect_head_conv(conv: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``Detect.m[i](x[i])`` (yolo.py:258) through dsn_detect_head_conv1x1 when the level qualifies (plain 1x1 conv, shapes
    the kernel takes), else the module itself."""
    if (isinstance(conv, nn.Conv2d) and tuple(conv.kernel_size) == (1, 1) and tuple(conv.stride) == (1, 1)
            and tuple(conv.padding) == (0, 0) and tuple(conv.dilation) == (1, 1) and conv.groups == 1
            and x.dtype == conv.weight.dtype and ops.detect_head_conv1x1_supported(x, conv.out_channels)
            # fp32: the kernel is an exact FMA chain; it stands in for the framework's convolution only when that one is
            # exact fp32 as well (cuDNN's default for fp32 convolutions is TF32, a different precision class and faster)
            and (x.dtype != torch.float32 or not torch.backends.cudnn.allow_tf32)):
        return ops.detect_head_conv1x1(x, conv.weight, conv.bias)
    return conv(x)


def _detect_forward_factory(original_forward, fused: bool, head_convs: bool = False):
    def forward(self, x):
        if self.training or not x[0].is_cuda:
            return original_forward(self, x)
        if head_convs:                                                       # the convs as producers of the tail
            raw = [_detect_head_conv(self.m[i], x[i]) for i in range(self.nl)]
        else:
            raw = [self.m[i](x[i]) for i in range(self.nl)]                  # yolo.py:258
        # host copy of anchor_grid / stride: once per head (re-read only when a buffer is updated in place)
        ver = (self.anchor_grid.data_ptr(), self.anchor_grid._version, self.stride.data_ptr(), self.stride._version)
        cached = getattr(self, "_dsn_spec", None)
        if cached is None or cached[0] != ver:
            cached = (ver, ops.HeadSpec(self.anchor_grid.reshape(self.nl, -1, 2), self.stride))
            object.__setattr__(self, "_dsn_spec", cached)
        ho = api.HeadOutputs(raw, cached[1], out_dtype=raw[0].dtype)
        if fused:
            return ho, _LazyPermuted(ho)
        z, xs = ho.z, ho.x                                                   # yolo.py:277
        if raw[0].dtype != torch.float32:
            z, xs = z.to(raw[0].dtype), [v.to(raw[0].dtype) for v in xs]
        return z, xs
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


def _find_originals(modules):
    """The reference's own function objects: from the defining modules, or (explicit ``modules``) from wherever
    the name is found first."""
    orig = {}
    if modules is None:
        for name, mods in _DEFINED_IN.items():
            for m in mods:
                mod = sys.modules.get(m)
                fn = getattr(mod, name, None) if mod is not None else None
                if fn is not None:
                    if getattr(fn, "_dsn_dispatch", False):        # installed before: the original is remembered on it
                        fn = fn._dsn_original
                    if fn is not None:
                        orig[name] = fn
                        break
    else:
        for m in modules:
            mod = sys.modules.get(m)
            if mod is None:
                continue
            for name in _DEFINED_IN:
                fn = getattr(mod, name, None)
                if fn is not None and name not in orig and not getattr(fn, "_dsn_dispatch", False):
                    orig[name] = fn
    return orig


def install(model: Optional[nn.Module] = None, fused: bool = True, modules: Optional[Iterable[str]] = None,
            head_convs: bool = False) -> Installed:
    """Patch every already-imported module that holds one of the reference's hot-path callables (wherever it was
    imported by name — including ``__main__`` when scripts/detect.py or scripts/val.py is the entry script); with
    ``model`` also defer the Detect decode and the seg head's final Upsample of that model instance.
    ``modules``: restrict the patch to these module names (their current attributes are taken as the originals).
    ``head_convs`` (opt-in, needs ``model``): also run the heads' last 1x1 convolutions in this package's kernels - the seg
    head's ``nn.Conv2d(c_hid, n_segcls, 1)`` (``FusedHeadConv``, streaming kernel) and the ``Detect.m[i]`` output convs
    (tcgen05 tensor cores for fp16 / bf16 models, exact fp32 FMA otherwise)."""
    h = Installed()
    modules = list(modules) if modules is not None else None
    orig = _find_originals(modules)
    disp = _make_dispatchers(orig)
    if modules is None:
        targets = [m for m in list(sys.modules.values()) if m is not None and getattr(m, "__name__", "").split(".")[0] != "desenet_b200"]
    else:
        targets = [sys.modules[m] for m in modules if m in sys.modules]
    for mod in targets:
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict):
            continue
        for name, fn in orig.items():
            cur = d.get(name)
            if cur is not None and cur is fn:                     # the very function object of the reference
                h._set(mod, name, disp[name])
    if model is not None:
        patched_classes = set()
        for m in model.modules():
            cls = type(m)
            if cls.__name__ == "Detect" and hasattr(m, "anchor_grid") and hasattr(m, "m") and cls not in patched_classes:
                h._set(cls, "forward", _detect_forward_factory(cls.forward, fused, head_convs))
                patched_classes.add(cls)
            if cls.__name__.startswith("SegMask"):
                # every seg head of the reference ends in nn.Sequential(..., nn.Conv2d(c, n_segcls, k), nn.Upsample(x8,
                # bilinear, align_corners=True)) — `out` (PSP yolo.py:178-184, PSP2 :227-230, BiSe :67-73), `decoder`
                # (Lab :118-124), `m` (Base :141-150); the aux heads (training only) have the same tail
                for seq in m.children():
                    if _is_seg_tail(seq):
                        idx = len(seq) - 1
                        h._undo.append((_SeqSlot(seq, idx), "module", seq[idx]))
                        seq[idx] = DeferredUpsample(seq[idx], fused)
                        if head_convs and FusedHeadConv.eligible(seq[idx - 1]):
                            seq[idx].out_dtype_of = seq[idx - 1].weight       # strict mode returns logits in the model's precision
                            h._undo.append((_SeqSlot(seq, idx - 1), "module", seq[idx - 1]))
                            seq[idx - 1] = FusedHeadConv(seq[idx - 1])
    return h


def _is_seg_tail(seq) -> bool:
    if not isinstance(seq, nn.Sequential) or len(seq) < 2:
        return False
    up, conv = seq[-1], seq[-2]
    if not isinstance(up, nn.Upsample) or up.mode != "bilinear" or not up.align_corners or up.scale_factor is None:
        return False
    sf = up.scale_factor
    sf = sf if not isinstance(sf, (tuple, list)) else (sf[0] if len(set(sf)) == 1 else None)
    return sf is not None and float(sf) == int(sf) and isinstance(conv, (nn.Conv2d, nn.Identity))


class _SeqSlot:
    """``setattr(slot, 'module', m)`` puts ``m`` back at ``seq[idx]`` (used by Installed.uninstall)."""

    def __init__(self, seq, idx):
        object.__setattr__(self, "_seq", seq)
        object.__setattr__(self, "_idx", idx)

    def __setattr__(self, name, value):
        self._seq[self._idx] = value
