"""Host-side mirror of the reference's interface for the post-network path: same names, argument
meaning and error behaviour as splwany/DeSeNet, so these callables drop into the reference's
``scripts/val.py`` / ``scripts/detect.py`` (see INTEGRATION.md and ``desenet_b200.patch``).

Two ways in:
  * strict — the reference's own tensors: ``non_max_suppression(prediction (bs,A,no), ...)``,
    ``segoutput_to_target(full-res logits, size)``, ``batch_*(full-res logits, target)``;
  * fused — hand the kernels what the network actually produced, before the reference materialises
    anything: ``HeadOutputs`` (raw Detect conv outputs) instead of ``prediction`` and ``SegLogits``
    (low-res logits before the final x8 Upsample) instead of full-res logits.  Same results, none of
    the z / x_i / full-res-logit traffic.
All arithmetic runs in the sm_100a kernels; there is no CPU fallback."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

__all__ = [
    "HeadOutputs", "SegLogits", "detect_forward", "non_max_suppression", "segoutput_to_target",
    "batch_pix_accuracy", "batch_intersection_union", "seg_confusion_matrix", "seg_stats_from_cm",
    "seg_validation", "xywh2xyxy", "scale_coords", "clip_coords", "box_iou", "process_batch",
    "ap_per_class", "compute_ap", "ConfusionMatrix",
]


def _kernel_input(t: torch.Tensor) -> torch.Tensor:
    """The kernels compute in fp32, like the reference's fp32 path, and load fp32 / fp16 / bf16 natively: half-precision
    head outputs (``--half``, val.py:144,178-179) are converted as they are read — the results are those of the fp32 path
    on ``t.float()`` without a conversion pass.  Any other floating type is upcast here."""
    return t if t.dtype in ops.DTYPE_CODE else t.float()


class _TensorLike:
    """Shared behaviour of ``HeadOutputs`` / ``SegLogits``: anything the kernels do not take over falls back to the
    materialised reference tensor (``torch.*`` functions through ``__torch_function__``, indexing, unknown attributes),
    so code that treats the object as the tensor the reference would have returned keeps working."""

    def _materialise(self) -> torch.Tensor:          # pragma: no cover - overridden
        raise NotImplementedError

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, _TensorLike):
                return a._materialise()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(v) for v in a)
            return a
        return func(*conv(tuple(args)), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getitem__(self, idx):
        return self._materialise()[idx]

    def __setitem__(self, idx, value):              # Model._descale_pred (yolo.py:360-366) writes into the prediction
        self._materialise()[idx] = value

    def __len__(self):
        return int(self.shape[0])

    def __getattr__(self, name):
        # only reached for attributes the class does not define: tensor methods such as .argmax / .softmax / .numpy
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._materialise(), name)

    def dim(self):
        return len(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    def size(self, dim=None):
        shp = torch.Size(self.shape)
        return shp if dim is None else shp[dim]

    def detach(self):
        return self

    def contiguous(self):
        return self

    def float(self):
        return self.to(torch.float32)

    def half(self):
        return self.to(torch.float16)

    def cpu(self):
        return self.to(torch.device("cpu"))

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device()) if device is None else device)

    def type_as(self, other):
        return self.to(other.dtype)


def _parse_to(args, kwargs, device, dtype):
    """(device, dtype) after ``Tensor.to(*args, **kwargs)``."""
    for a in list(args) + [kwargs.get("device"), kwargs.get("dtype"), kwargs.get("other")]:
        if a is None:
            continue
        if isinstance(a, torch.dtype):
            dtype = a
        elif isinstance(a, (torch.device, str, int)):
            device = torch.device("cuda", a) if isinstance(a, int) else torch.device(a)
        elif isinstance(a, torch.Tensor):
            device, dtype = a.device, a.dtype
    return device, dtype


class HeadOutputs(_TensorLike):
    """Raw outputs of the three ``Detect.m[i]`` convs (core/models/yolo.py:258) + the head's
    ``anchor_grid`` / ``stride`` buffers.  ``.z`` / ``.x`` materialise what ``Detect.forward`` returns
    (yolo.py:277) on demand; ``non_max_suppression`` accepts the object directly (fused path).
    ``out_dtype`` is the dtype the reference's ``z`` would have had (fp16 under ``--half``); everything is computed in
    fp32 and only a materialised ``z`` is cast to it."""

    def __init__(self, levels: Sequence[torch.Tensor], anchor_grid, stride=None, out_dtype=None):
        self._out_dtype = out_dtype if out_dtype is not None else levels[0].dtype
        self.head = ops.Head([_kernel_input(l) for l in levels], anchor_grid, stride)
        self._z = None
        self._zc = None
        self._x = None

    @property
    def shape(self):
        return (self.head.bs, self.head.A, self.head.no)

    @property
    def device(self):
        return self.head.device

    @property
    def dtype(self):
        return self._out_dtype

    @property
    def z(self) -> torch.Tensor:
        if self._z is None:
            self._z, _ = ops.detect_decode(self.head, want_z=True, want_x=False)
        return self._z

    @property
    def x(self) -> List[torch.Tensor]:
        if self._x is None:
            _, self._x = ops.detect_decode(self.head, want_z=False, want_x=True)
        return self._x

    def _materialise(self):
        # cached, so that in-place edits (augmented inference de-scales the prediction in place) are kept
        if self._out_dtype == torch.float32:
            return self.z
        if self._zc is None:
            self._zc = self.z.to(self._out_dtype)
        return self._zc

    def to(self, *args, **kwargs):
        device, dtype = _parse_to(args, kwargs, self.device, self._out_dtype)
        if device.type != "cuda":
            return self.z.to(device=device, dtype=dtype)       # leaves the device: hand over the reference tensor
        if device != self.device:
            return self.z.to(device=device, dtype=dtype)
        ho = HeadOutputs.__new__(HeadOutputs)
        ho.head, ho._z, ho._zc, ho._x, ho._out_dtype = self.head, self._z, None, self._x, dtype
        return ho


class SegLogits(_TensorLike):
    """Low-res seg logits (bs, C, h, w) — the input of the head's final
    ``nn.Upsample(scale_factor=8, mode='bilinear', align_corners=True)`` (core/models/yolo.py:183) —
    plus the size that Upsample would produce.  ``.full()`` materialises the reference tensor.

    The object stands in for the full-resolution logits the reference model returns, so it answers ``.shape`` /
    ``.device`` / ``.dtype`` as that tensor would, and ``.cpu()`` / ``.to()`` / ``.float()`` / ``.half()`` return a
    ``SegLogits`` that remembers where the caller wants results (scripts/detect.py:198 calls
    ``segoutput_to_target(seg_pred.cpu(), size=...)``): the low-res logits stay on the GPU, the kernels run there,
    and only the RESULT (the mask) is moved to the requested device — 1.6 MB per image instead of a 13 MB logit D2H
    plus a CPU argmax."""

    def __init__(self, low: torch.Tensor, scale_factor: int = 8, size: Optional[Tuple[int, int]] = None,
                 out_device=None, out_dtype=None):
        self.low = low
        self.full_hw = (int(low.shape[2]) * scale_factor, int(low.shape[3]) * scale_factor) if size is None else \
            (int(size[0]), int(size[1]))
        self.out_device = torch.device(out_device) if out_device is not None else low.device
        self.out_dtype = out_dtype if out_dtype is not None else low.dtype

    @property
    def shape(self):
        return (int(self.low.shape[0]), int(self.low.shape[1])) + tuple(self.full_hw)

    @property
    def device(self):
        return self.out_device

    @property
    def dtype(self):
        return self.out_dtype

    @property
    def low_in(self) -> torch.Tensor:
        return _kernel_input(self.low)

    def full(self) -> torch.Tensor:
        """The full-resolution fp32 logits on the GPU (bit-identical to ATen's CPU upsample of the fp32 logits)."""
        return ops.upsample_bilinear_ac(self.low_in, self.full_hw)

    def _materialise(self):
        return self.full().to(device=self.out_device, dtype=self.out_dtype)

    def to(self, *args, **kwargs):
        device, dtype = _parse_to(args, kwargs, self.out_device, self.out_dtype)
        return SegLogits(self.low, size=self.full_hw, out_device=device, out_dtype=dtype)


# ------------------------------------------------------------------------------------------------
# a1  Detect.forward, eval branch after the convs                      core/models/yolo.py:259-277
# ------------------------------------------------------------------------------------------------
def detect_forward(x: Sequence[torch.Tensor], anchor_grid, stride) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """``x[i]`` = output of ``Detect.m[i]`` (bs, na*no, ny, nx).  Returns ``(cat(z,1), x)`` exactly
    as yolo.py:277: z (bs, A, no) decoded, x[i] (bs, na, ny, nx, no) permuted raw logits."""
    head = ops.Head(x, anchor_grid, stride)
    z, xs = ops.detect_decode(head, want_z=True, want_x=True)
    return z, xs


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """core/utils/general.py:523-530 (tiny host-side helper, kept for API completeness)."""
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


# ------------------------------------------------------------------------------------------------
# a2  non_max_suppression                                          core/utils/general.py:659-750
# ------------------------------------------------------------------------------------------------
def _append_label_rows(z: torch.Tensor, labels) -> torch.Tensor:
    """general.py:690-696: a-priori label rows [cls, x, y, w, h] become extra prediction rows with
    obj = 1 and a one-hot class; appended after the A anchor rows so candidate order is unchanged."""
    bs, A, no = z.shape
    L = max(len(l) for l in labels)
    if L == 0:
        return z
    extra = torch.zeros((bs, L, no), dtype=z.dtype, device=z.device)
    for i, l in enumerate(labels):
        if len(l):
            l = torch.as_tensor(l, dtype=z.dtype, device=z.device)
            n = l.shape[0]
            extra[i, :n, :4] = l[:, 1:5]
            extra[i, :n, 4] = 1.0
            extra[i, torch.arange(n, device=z.device), l[:, 0].long() + 5] = 1.0
    return torch.cat((z, extra), 1)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300) -> List[torch.Tensor]:
    """Same signature and result as the reference (general.py:659-750): list of (n_i, 6) tensors
    [x1, y1, x2, y2, conf, cls] per image, sorted by confidence, on ``prediction.device``.
    ``prediction`` is the decoded (bs, A, no) tensor, or a ``HeadOutputs`` (fused path).
    Differences, both documented in DESIGN.md: no 10 s wall-clock ``break`` (general.py:746-748), and
    among equal confidences beyond 30 000 candidates the order is canonical (stable) where the
    reference's non-stable argsort (general.py:729) leaves it implementation-defined."""
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    has_labels = bool(labels) and any(len(l) for l in labels)
    if isinstance(prediction, HeadOutputs):
        head = prediction.head
        ml = bool(multi_label) and head.nc > 1
        if has_labels:
            # a-priori labels (general.py:690-696) join the candidates as rows A + j of confidence 1.0; filter + NMS as
            # two launches, without the confidence cut (autolabelling runs are not the throughput case)
            padded = ops.pad_labels(labels, head.bs, head.device)
            cap = head.A * (head.nc if ml else 1) + int(padded[0].shape[1])
            keys, counts = ops.filter_levels(head, conf_thres, ml, classes, cap=cap)
            ops.append_label_keys(head, padded, conf_thres, classes, keys, counts)
            det, det_count = ops.nms_levels(head, keys, counts, conf_thres, iou_thres, agnostic, int(max_det), labels=padded)
        else:
            target = ops.cut_target_for(max_det)
            cut = None
            if ml and head.A * head.nc >= 8 * target:      # eval setting: sampled confidence cut (exact, see ops.conf_cut_levels)
                cut = ops.conf_cut_levels(head, conf_thres, ml, classes, target=target)
            if cut is not None:
                keys, counts = ops.filter_levels(head, conf_thres, ml, classes, cut=cut)
                det, det_count = ops.nms_levels(head, keys, counts, conf_thres, iou_thres, agnostic, int(max_det), cut=cut)
                # images whose NMS ran out of emitted candidates (det_count == -2) are redone uncut on the device
                ops.detect_nms_redo_levels(head, conf_thres, iou_thres, classes, agnostic, ml, int(max_det), keys, counts, det, det_count)
            else:
                det, det_count, _ = ops.detect_nms_levels(head, conf_thres, iou_thres, classes, agnostic, multi_label, int(max_det))
    else:
        z = _kernel_input(prediction)
        if has_labels:
            z = _append_label_rows(z, labels)
        if z.shape[0] == 0:
            return []
        keys, counts = ops.filter_z(z, conf_thres, multi_label, classes)
        det, det_count = ops.nms_z(z, keys, counts, conf_thres, iou_thres, agnostic, int(max_det))
    n = det_count.cpu().tolist()                       # the single host sync of the call
    if any(v < 0 for v in n):
        raise RuntimeError("desenet_b200: candidate buffer overflow (internal capacity error)")
    return [det[i, :n[i]] for i in range(len(n))]      # views of a per-call buffer: caller-owned


# ------------------------------------------------------------------------------------------------
# f1  scale_coords / clip_coords        core/utils/general.py:598-623
# ------------------------------------------------------------------------------------------------
_PARAM_CACHE = {}


def _letterbox_params_device(img1_shape, img0_shape, ratio_pad, device) -> torch.Tensor:
    """(1, 5) fp32 device tensor of ``ops.letterbox_params``, cached per (shapes, ratio_pad, device): callers run
    scale_coords once per image (val.py:250, detect.py:218), almost always with the same few shapes, and building a
    device tensor from Python floats costs a pageable H2D copy each time."""
    try:
        key = (tuple(int(v) for v in img1_shape[:2]), tuple(int(v) for v in img0_shape[:2]),
               None if ratio_pad is None else (float(ratio_pad[0][0]), float(ratio_pad[1][0]), float(ratio_pad[1][1])), str(device))
    except (TypeError, ValueError, IndexError):
        key = None
    t = _PARAM_CACHE.get(key) if key is not None else None
    if t is None:
        t = torch.tensor([ops.letterbox_params(img1_shape, img0_shape, ratio_pad)], dtype=torch.float32, device=device)
        if key is not None:
            if len(_PARAM_CACHE) > 4096:
                _PARAM_CACHE.clear()
            _PARAM_CACHE[key] = t
    return t


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """Same signature and in-place behaviour as the reference: rescale ``coords`` (n, 4) xyxy (any row-strided
    view, e.g. ``det[:, :4]``) from ``img1_shape`` to ``img0_shape`` and clip; returns ``coords``."""
    return ops.scale_coords_(coords, _letterbox_params_device(img1_shape, img0_shape, ratio_pad, coords.device))


def clip_coords(boxes, shape):
    """general.py:614-623 on a CUDA tensor (n, 4), in place: clamp to the image ``shape`` (h, w)."""
    ops.scale_coords_(boxes, _letterbox_params_device((int(shape[0]), int(shape[1])), (int(shape[0]), int(shape[1])),
                                                      ((1.0,), (0.0, 0.0)), boxes.device))


# ------------------------------------------------------------------------------------------------
# f2  box_iou / process_batch   general.py:626-648, scripts/val.py:101-122
# ------------------------------------------------------------------------------------------------
def box_iou(box1, box2):
    """general.py:626-648: (N, M) IoU matrix, fp32, on the device."""
    return ops.box_iou(box1, box2)


def process_batch(detections, labels, iouv):
    """Same signature and result as scripts/val.py:101-122: ``detections`` (N, 6) [xyxy, conf, cls], ``labels`` (M, 5)
    [cls, xyxy] -> ``correct`` (N, len(iouv)) bool on ``iouv.device``.  Equal IoUs of one detection with two
    labels resolve to the lowest label index (the reference's order there depends on numpy's unstable argsort)."""
    n, m = int(detections.shape[0]), int(labels.shape[0])
    if n == 0 or m == 0:
        return torch.zeros(n, iouv.shape[0], dtype=torch.bool, device=iouv.device)
    off = torch.tensor([0, m], dtype=torch.int32, device=detections.device)
    c = ops.process_batch(detections, None, labels.float().contiguous(), off, iouv.float().contiguous(), m)
    return c[0].bool()


# ------------------------------------------------------------------------------------------------
# f3  ap_per_class / compute_ap / ConfusionMatrix   core/utils/metrics.py:31-170
# ------------------------------------------------------------------------------------------------
def _to_device(a, dtype, device):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=device, dtype=dtype, non_blocking=True)


def _stats_device(*arrays):
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
    if not torch.cuda.is_available():
        raise RuntimeError("desenet_b200: ap_per_class needs a CUDA device; this package has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, save_dir='.', names=()):
    """Same signature and result as the reference (metrics.py:31-92): ``(p, r, ap, f1, unique_classes)`` as numpy
    arrays.  Inputs may be numpy arrays (what val.py:282 hands over) or torch tensors; the sort, the cumulative
    counts, the AP integration and the 1000-point precision / recall curves run on the device, bit-identical to the
    reference's numpy float64 arithmetic.  The (ncls, 1000) F1 tail (metrics.py:85-91) is evaluated here from the
    device curves.  ``plot`` is not supported (plotting is out of scope); equal confidences keep their input order."""
    if plot:
        raise NotImplementedError("desenet_b200.ap_per_class: plotting (metrics.py:82-88) is out of scope")
    dev = _stats_device(tp, conf, pred_cls, target_cls)
    conf_d = _to_device(conf, torch.float32, dev).reshape(-1)
    n = int(conf_d.shape[0])
    tp_d = tp if isinstance(tp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(tp))
    tp_d = (tp_d != 0).to(dev).reshape(n, int(tp_d.shape[1]) if tp_d.dim() == 2 else 1)
    pred_d = _to_device(pred_cls, torch.float32, dev).reshape(-1)
    tgt_d = _to_device(target_cls, torch.float32, dev).reshape(-1)
    classes, n_labels = torch.unique(tgt_d, return_counts=True)                    # metrics.py:52,59 (sorted ascending)
    ap, p, r = ops.ap_per_class_device(tp_d, conf_d, pred_d, classes, n_labels)
    ap, p, r = ap.cpu().numpy(), p.cpu().numpy(), r.cpu().numpy()
    f1 = 2 * p * r / (p + r + 1e-16)                                               # metrics.py:85
    i = f1.mean(0).argmax() if f1.shape[0] else 0                                  # metrics.py:90
    return p[:, i], r[:, i], ap, f1[:, i], classes.cpu().numpy().astype('int32')


def compute_ap(recall, precision):
    """metrics.py:95-121: ``(ap, mpre, mrec)`` for one recall / precision curve (numpy in, numpy out; the envelope,
    the 101-point interpolation and the trapezoid sum run on the device)."""
    dev = _stats_device(recall, precision)
    rec = _to_device(recall, torch.float64, dev).reshape(-1)
    pre = _to_device(precision, torch.float64, dev).reshape(-1)
    ap, env = ops.compute_ap_device(rec, pre)
    n = int(rec.shape[0])
    e = env[:n].cpu().numpy()
    mpre = np.concatenate(([max(1.0, float(e[0])) if n else 1.0], e, [0.0]))
    mrec = np.concatenate(([0.0], rec.cpu().numpy(), [1.0]))
    return float(ap.item()), mpre, mrec


class ConfusionMatrix:
    """Detection confusion matrix with the reference's interface (metrics.py:124-172; plot / print out of scope).
    ``process_batch(detections, labels)`` takes one image like the reference; ``process_images`` takes the batched
    NMS output buffers so that a whole batch is one launch and nothing is copied to the host."""

    def __init__(self, nc, conf=0.25, iou_thres=0.45, device=None):
        self.nc, self.conf, self.iou_thres = int(nc), conf, iou_thres
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._m = torch.zeros((self.nc + 1, self.nc + 1), dtype=torch.int64, device=self.device)

    def process_batch(self, detections, labels):
        det = detections.to(self.device, torch.float32)
        lab = labels.to(self.device, torch.float32).reshape(-1, 5)
        off = torch.tensor([0, lab.shape[0]], dtype=torch.int32, device=self.device)
        det = det.contiguous() if det.numel() else det.new_zeros((0, 6))
        ops.det_confusion_matrix_(self._m, det, None, lab, off, int(lab.shape[0]), self.nc, self.conf, self.iou_thres)

    def process_images(self, det, det_count, labels, label_off, max_labels_per_image, skip_empty=False):
        """``skip_empty``: images without detections contribute nothing (what val.run's loop does, val.py:241-244)."""
        ops.det_confusion_matrix_(self._m, det, det_count, labels, label_off, max_labels_per_image, self.nc, self.conf,
                                  self.iou_thres, skip_empty=skip_empty)

    def matrix(self):
        """(nc+1, nc+1) float64 numpy array like the reference's ``self.matrix`` (metrics.py:127)."""
        return self._m.cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------------------------------------
# a6  segoutput_to_target                                              core/utils/plots.py:222-229
# ------------------------------------------------------------------------------------------------
def segoutput_to_target(output, size=None) -> torch.Tensor:
    """argmax over classes -> float -> nearest resize to ``size`` -> (bs, size[0], size[1]) float32.
    ``output``: full-res logits (bs, C, H, W) as in the reference (fp16 / bf16 are upcast), or ``SegLogits`` (fused
    path; the mask is returned on the device the ``SegLogits`` stands for, e.g. the CPU after ``.cpu()``)."""
    if isinstance(output, SegLogits):
        full = output.full_hw
        msize = full if size is None else tuple(int(v) for v in size)[-2:]
        mask, _ = ops.seg_fused(output.low_in, full, mask_dtype=torch.float32, mask_size=msize)
        return mask if output.out_device == mask.device else mask.to(output.out_device)
    output = _kernel_input(output)
    H, W = int(output.shape[2]), int(output.shape[3])
    msize = (H, W) if size is None else tuple(int(v) for v in size)[-2:]
    mask, _ = ops.seg_fullres(output, mask_dtype=torch.float32, mask_size=msize)
    return mask


# ------------------------------------------------------------------------------------------------
# a7 / a8  seg metrics                                               core/utils/metrics.py:350-388
# ------------------------------------------------------------------------------------------------
def seg_confusion_matrix(output, target: torch.Tensor, nclass: int, cm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Target-clamped confusion matrix CM[r][p] (device int64, shape (slots, rows, C); sum over
    slots) of argmax(output) against ``target`` — the sufficient statistic of metrics.py:350-388.
    ``target`` may have any integer / float dtype and live on any device (the reference takes it through
    ``.cpu().numpy().astype('int32')``, metrics.py:359,380); fp16 / bf16 logits are upcast."""
    C = int(output.shape[1])
    rows = ops.cm_rows_for(C, nclass)
    dev = output.low.device if isinstance(output, SegLogits) else output.device
    if not isinstance(target, torch.Tensor):
        target = torch.as_tensor(target)
    if target.device != dev or target.dtype != torch.int64:
        target = target.to(device=dev, dtype=torch.int64)
    if isinstance(output, SegLogits):
        _, cm = ops.seg_fused(output.low_in, output.full_hw, target=target, cm=cm, cm_rows=rows)
    else:
        _, cm = ops.seg_fullres(_kernel_input(output), target=target, cm=cm, cm_rows=rows)
    return cm


def seg_stats_from_cm(cm: np.ndarray, nclass: int):
    """(pixel_correct, pixel_labeled, area_inter[nclass-1], area_union[nclass-1]) from CM[r][p].

    Closed form of metrics.py:350-388: labeled = target > 0 (metrics.py:361); correct = predict ==
    target among labeled (:362); the three np.histogram(bins=nclass-1, range=(1, nclass)) calls
    (:383-385) put value k in bin min(k, nclass-1)-1 for 1 <= k <= nclass (last bin right-closed) and
    drop everything else; predictions are < C, exact target rows are 0..rows-2, the last row holds
    every larger target (labeled, never a class, never equal to a prediction)."""
    cm = np.asarray(cm, dtype=np.int64)
    if cm.ndim == 3:
        cm = cm.sum(0)
    rows, C = cm.shape
    exact = rows - 1                      # rows 0..exact-1 are exact target values
    labeled = np.int64(cm[1:].sum())
    correct = np.int64(sum(cm[k, k] for k in range(1, min(exact, C))))
    nb = nclass - 1
    inter = np.zeros(nb, np.int64)
    pred = np.zeros(nb, np.int64)
    lab = np.zeros(nb, np.int64)
    for k in range(1, nclass + 1):
        b = min(k, nclass - 1) - 1
        if k < C:
            pred[b] += cm[:, k].sum()
            if k < exact:
                inter[b] += cm[k, k]
        if k < exact:
            lab[b] += cm[k].sum()
    return correct, labeled, inter, pred + lab - inter


def batch_pix_accuracy(output, target):
    """metrics.py:350-365: returns (pixel_correct, pixel_labeled) as numpy int64 scalars."""
    C = int(output.shape[1])
    cm = seg_confusion_matrix(output, target, max(C - 1, 2)).sum(0).cpu().numpy()
    correct, labeled, _, _ = seg_stats_from_cm(cm, max(C - 1, 2))
    assert correct <= labeled, "Correct area should be smaller than Labeled"
    return correct, labeled


def batch_intersection_union(output, target, nclass):
    """metrics.py:368-388: returns (area_inter, area_union), int64 arrays of length nclass-1."""
    cm = seg_confusion_matrix(output, target, nclass).sum(0).cpu().numpy()
    _, _, inter, union = seg_stats_from_cm(cm, nclass)
    assert (inter <= union).all(), "Intersection area should be smaller than Union area"
    return inter, union


# ------------------------------------------------------------------------------------------------
# a9  seg_validation                                                      scripts/val.py:42-76
# ------------------------------------------------------------------------------------------------
def seg_validation(model, n_segcls, valloader, half_precision=True, device=None, progress=None):
    """Same contract as scripts/val.py:42-76: loops the loader, accumulates pixel / IoU statistics,
    returns mIoU.  The statistics are accumulated ON DEVICE in one int64 confusion matrix (one
    D2H at the end instead of four full-resolution maps per batch); pixAcc / IoU / mIoU are then
    formed exactly as val.py:71-73 (float64, ``np.spacing(1)`` eps).  ``model(image)[1]`` may return
    full-res logits (reference model) or ``SegLogits`` (patched head).  fp16 logits
    (``half_precision``) are upcast to fp32 on load."""
    import torch.nn.functional as F

    first = next(iter(model.parameters()), None) if hasattr(model, "parameters") else None
    if device is None:
        device = first.device if first is not None else torch.device("cuda")
    half = half_precision and device.type != "cpu" and first is not None
    if half:
        model.half()                                                              # val.py:54-56
    if hasattr(model, "eval"):
        model.eval()
    cm = None
    rows = None
    it = valloader if progress is None else progress(valloader)
    with torch.no_grad():
        for image, _, target, *_ in it:                                            # val.py:61
            image = image.to(device, non_blocking=True)
            image = image.half() if half else image.float()
            image = image / 255.0                                                  # val.py:64
            target = target.to(device, non_blocking=True)
            pred = model(image)[1]                                                 # val.py:45
            if isinstance(pred, SegLogits) and tuple(pred.full_hw) != tuple(target.shape[1:]):
                pred = pred.full()
            if not isinstance(pred, SegLogits):
                pred = pred.float()
                if tuple(pred.shape[2:]) != tuple(target.shape[1:]):               # val.py:47
                    pred = F.interpolate(pred, tuple(target.shape[1:]), mode="bilinear", align_corners=False)
            C = int(pred.shape[1])
            if cm is None:
                rows = ops.cm_rows_for(C, n_segcls)
                cm_dev = pred.low.device if isinstance(pred, SegLogits) else pred.device
                cm = torch.zeros((64, rows, C), dtype=torch.int64, device=cm_dev)
            seg_confusion_matrix(pred, target, n_segcls, cm=cm)                    # val.py:48-49
    if cm is None:
        return 0.0
    correct, labeled, inter, union = seg_stats_from_cm(cm.sum(0).cpu().numpy(), n_segcls)
    pix_acc = 1.0 * correct / (np.spacing(1) + labeled)                            # val.py:71
    iou = 1.0 * inter / (np.spacing(1) + union)                                    # val.py:72
    miou = iou.mean()                                                              # val.py:73
    seg_validation.last = dict(pixAcc=float(pix_acc), mIoU=float(miou), cm=cm.sum(0))
    return float(miou)
