"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of DeSeNet's post-network hot path.

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package ``desenet_b200`` never imports anything under ``oracle/`` and has no CPU fallback.

What is restated (all ``file:line`` are into the reference tree, splwany/DeSeNet):

  decode_levels            core/models/yolo.py:255-277   Detect.forward, eval branch (post-conv)
  xywh2xyxy                core/utils/general.py:523-530
  non_max_suppression      core/utils/general.py:659-750 (per image, no wall-clock break)
  nms_greedy_np            torchvision.ops.nms CPU kernel (third-party, call site general.py:734;
                           requirement ``torchvision>=0.9.1`` requirements.txt:11, installed
                           0.26.0+cu128).  Published algorithm restated, see docstring.
  upsample_x8_aten         core/models/yolo.py:183       nn.Upsample(8, bilinear, align_corners=True)
  upsample_bilinear_ac_np  the same op as explicit fp32 arithmetic incl. ATen's FMA contraction
                           (bit-identical to ATen CPU; what the CUDA kernel matches bit-for-bit)
  segoutput_to_target      core/utils/plots.py:222-229
  nearest_resize_np        ATen legacy 'nearest' index rule used by plots.py:229
  batch_pix_accuracy       core/utils/metrics.py:350-365
  batch_intersection_union core/utils/metrics.py:368-388
  confusion_matrix / seg_stats_from_cm   closed form of the two functions above
  seg_validation_accumulate scripts/val.py:58-76

Parity pinning: the reference ships no tests or golden vectors for this path.  This
port is pinned (a) against the unmodified reference imported in place (``oracle/ref_shim.py``,
``tests/test_oracle_vs_reference.py`` — runs wherever ``/root/reference`` exists) and (b) against
fixtures generated from the real reference and committed under ``tests/golden/`` (generator:
``tests/golden/make_golden.py``), which travel to the GPU box.

Third-party arithmetic (``torch.sigmoid``, ``torchvision.ops.nms``, ``F.interpolate``,
``np.histogram``) is called through the same installed wheels the reference calls, because the
reference's results are *defined* by them; independent restatements (numpy / C under
``oracle/c``) are provided for the pieces the CUDA kernels must match bit-for-bit.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MAX_WH = 4096          # general.py:675
MAX_NMS = 30000        # general.py:676

# yolov5s(_seg).yaml anchors in pixels, P3/P4/P5 (core/models/yolov5s_seg.yaml:8-11)
DEFAULT_ANCHORS = ((10, 13, 16, 30, 33, 23), (30, 61, 62, 45, 59, 119), (116, 90, 156, 198, 373, 326))
DEFAULT_STRIDES = (8.0, 16.0, 32.0)


# --------------------------------------------------------------------------------------------
# a1  Detect.forward (eval branch, everything after the 1x1 convs)        yolo.py:255-277
# --------------------------------------------------------------------------------------------
def decode_levels(raw_levels: Sequence[torch.Tensor], anchor_grid: torch.Tensor,
                  strides: Sequence[float], nc: int) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """raw_levels[i]: (bs, na*no, ny, nx) fp32 conv outputs.  anchor_grid: (nl, na, 2) pixel
    anchors (the ``anchor_grid`` buffer of yolo.py:251).  Returns ``(z, [x_i])`` exactly as
    yolo.py:277: z (bs, A, no), x_i (bs, na, ny, nx, no) = permuted raw logits (yolo.py:260)."""
    no = nc + 5
    anchor_grid = torch.as_tensor(anchor_grid, dtype=torch.float32).reshape(len(raw_levels), -1, 2)
    na = anchor_grid.shape[1]
    zs, xs = [], []
    for lvl, raw in enumerate(raw_levels):
        bs, ch, ny, nx = raw.shape
        assert ch == na * no
        # channel = a*no + o  ->  (bs, a, gy, gx, o)                                   yolo.py:260
        xp = raw.reshape(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        y = torch.sigmoid(xp)                                                        # yolo.py:266
        gx = torch.arange(nx, dtype=torch.float32).view(1, 1, 1, nx)
        gy = torch.arange(ny, dtype=torch.float32).view(1, 1, ny, 1)                 # yolo.py:279-282
        s = float(strides[lvl])
        out = y.clone()
        out[..., 0] = (y[..., 0] * 2.0 - 0.5 + gx) * s                               # yolo.py:268
        out[..., 1] = (y[..., 1] * 2.0 - 0.5 + gy) * s
        tw = y[..., 2] * 2.0
        th = y[..., 3] * 2.0
        aw = anchor_grid[lvl, :, 0].view(1, na, 1, 1)
        ah = anchor_grid[lvl, :, 1].view(1, na, 1, 1)
        out[..., 2] = tw * tw * aw                                                   # yolo.py:269
        out[..., 3] = th * th * ah
        zs.append(out.reshape(bs, na * ny * nx, no))                                 # yolo.py:275
        xs.append(xp)
    return torch.cat(zs, 1), xs


# --------------------------------------------------------------------------------------------
# a3  xywh2xyxy                                                          general.py:523-530
# --------------------------------------------------------------------------------------------
def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    hw, hh = x[:, 2] / 2, x[:, 3] / 2
    return torch.stack((x[:, 0] - hw, x[:, 1] - hh, x[:, 0] + hw, x[:, 1] + hh), 1)


# --------------------------------------------------------------------------------------------
# a4  torchvision.ops.nms, CPU kernel — restated
# --------------------------------------------------------------------------------------------
def nms_greedy_np(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Greedy NMS exactly as torchvision's CPU kernel computes it (probed on 0.26):
    stable descending score order (ties: lower index first); fp32 ``area=(x2-x1)*(y2-y1)``;
    ``w=max(0,min(x2)-max(x1))``, ``h`` likewise; ``inter=w*h``; ``ovr=inter/(a_i+a_j-inter)`` in
    fp32; suppress iff ``double(ovr) > iou_threshold`` (strict; NaN never suppresses).  Returns
    kept indices (int64) in descending-score order."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable")   # exact negate, stable
    b = boxes[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    thr = np.float64(iou_threshold)
    dead = np.zeros(n, dtype=bool)
    keep = []
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        for i in range(n):
            if dead[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            zero = np.float32(0)
            w = np.maximum(zero, np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
            h = np.maximum(zero, np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
            inter = w * h
            ovr = inter / (area[i] + area[i + 1:] - inter)
            dead[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def nms_torchvision(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    import torchvision
    return torchvision.ops.nms(boxes, scores, iou_threshold)                        # general.py:734


# --------------------------------------------------------------------------------------------
# a2  non_max_suppression                                               general.py:659-750
# --------------------------------------------------------------------------------------------
def nms_candidates(x: torch.Tensor, conf_thres: float, classes, multi_label: bool,
                   labels_i: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Steps (1)-(6) of general.py:684-729 for ONE image: ``x`` (A, no) decoded rows ->
    candidate matrix (n, 6) = [x1,y1,x2,y2,conf,cls] in reference order, already capped to the
    top ``MAX_NMS`` by confidence.  Tie order inside the cap is *canonical* here (stable: lower
    candidate index first); the reference's non-stable ``argsort`` (general.py:729) leaves it
    implementation-defined."""
    nc = x.shape[1] - 5
    multi_label = bool(multi_label) and nc > 1                                      # general.py:679
    x = x[x[:, 4] > conf_thres]                                                      # :668,:687
    if labels_i is not None and len(labels_i):                                       # :690-696
        lb = labels_i
        v = torch.zeros((len(lb), nc + 5), dtype=x.dtype)
        v[:, :4] = lb[:, 1:5]
        v[:, 4] = 1.0
        v[torch.arange(len(lb)), lb[:, 0].long() + 5] = 1.0
        x = torch.cat((x, v), 0)
    if x.shape[0] == 0:
        return torch.zeros((0, 6))
    cls_conf = x[:, 5:] * x[:, 4:5]                                                  # :703
    box = xywh2xyxy(x[:, :4])                                                        # :706
    if multi_label:
        i, j = (cls_conf > conf_thres).nonzero(as_tuple=False).T                     # :710 (row-major)
        det = torch.cat((box[i], cls_conf[i, j, None], j[:, None].float()), 1)
    else:
        conf, j = cls_conf.max(1, keepdim=True)                                      # :713 first max
        det = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
    if classes is not None:                                                          # :717-718
        det = det[(det[:, 5:6] == torch.tensor(classes, dtype=det.dtype)).any(1)]
    if det.shape[0] > MAX_NMS:                                                       # :728-729
        order = torch.sort(det[:, 4], descending=True, stable=True).indices[:MAX_NMS]
        det = det[order]
    return det


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes=None, agnostic: bool = False, multi_label: bool = False,
                        labels=(), max_det: int = 300, nms_impl: str = "torchvision") -> List[torch.Tensor]:
    """Same signature and result as general.py:659-750, minus the 10 s wall-clock ``break``
    (:746-748), i.e. identical to calling the reference one image at a time."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}"       # :671
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}"                           # :672
    out: List[torch.Tensor] = []
    for xi in range(prediction.shape[0]):
        lab = labels[xi] if (labels and len(labels[xi])) else None
        det = nms_candidates(prediction[xi], conf_thres, classes, multi_label, lab)
        if det.shape[0] == 0:
            out.append(torch.zeros((0, 6)))
            continue
        off = det[:, 5:6] * (0 if agnostic else MAX_WH)                              # :732
        boxes, scores = det[:, :4] + off, det[:, 4]                                  # :733
        if nms_impl == "torchvision":
            keep = nms_torchvision(boxes, scores, iou_thres)
        elif nms_impl == "numpy":
            keep = torch.from_numpy(nms_greedy_np(boxes.numpy(), scores.numpy(), iou_thres))
        else:
            from oracle import cport
            keep = torch.from_numpy(cport.nms_greedy(boxes.numpy(), scores.numpy(), iou_thres))
        keep = keep[:max_det]                                                        # :735-736
        out.append(det[keep])                                                        # :745
    return out


# --------------------------------------------------------------------------------------------
# a5  final x8 bilinear upsample of the seg head                              yolo.py:183
# --------------------------------------------------------------------------------------------
def upsample_x8_aten(logits: torch.Tensor, scale: int = 8) -> torch.Tensor:
    """What the reference executes: ATen ``upsample_bilinear2d`` (align_corners=True)."""
    return F.interpolate(logits, scale_factor=scale, mode="bilinear", align_corners=True)


def bilinear_ac_coeffs(n_in: int, n_out: int):
    """ATen's align_corners=True source index rule, in fp32: ``scale=(in-1)/(out-1)`` (0 when
    out==1); ``src=scale*dst``; ``i0=int(src)``; ``i1=i0+(i0<in-1)``; ``l1=src-i0``; ``l0=1-l1``."""
    scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
    dst = np.arange(n_out, dtype=np.float32)
    src = (scale * dst).astype(np.float32)
    i0 = src.astype(np.int64)
    i0 = np.minimum(i0, n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def _fma32(a, b, c):
    """fp32 fma emulated in float64 (a*b is exact in float64; the single extra rounding of the
    sum can differ from a true fma only in double-rounding corner cases — the C oracle
    ``oracle/c/oracle.c`` uses a real ``fmaf`` and is the authority)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def upsample_bilinear_ac_np(logits: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Explicit fp32 restatement of ATen's CPU kernel: every lerp ``w0*a + w1*b`` is
    ``fma(w0, a, w1*b)`` (probed: 100 % bit-identical to ``F.interpolate`` on torch 2.11 CPU);
    horizontal lerps first, then vertical: ``out = fma(l0y, T, l1y*B)``,
    ``T = fma(l0x, v00, l1x*v01)``, ``B = fma(l0x, v10, l1x*v11)``.  The CUDA kernels
    reproduce THIS bit-for-bit."""
    v = np.ascontiguousarray(logits, dtype=np.float32)
    y0, y1, ly0, ly1 = bilinear_ac_coeffs(v.shape[-2], out_h)
    x0, x1, lx0, lx1 = bilinear_ac_coeffs(v.shape[-1], out_w)
    top = v[..., y0, :]
    bot = v[..., y1, :]
    h_top = _fma32(lx0, top[..., x0], (lx1 * top[..., x1]).astype(np.float32))
    h_bot = _fma32(lx0, bot[..., x0], (lx1 * bot[..., x1]).astype(np.float32))
    ly0 = ly0[:, None]
    ly1 = ly1[:, None]
    return _fma32(ly0, h_top, (ly1 * h_bot).astype(np.float32))


# --------------------------------------------------------------------------------------------
# a6  segoutput_to_target                                                 plots.py:222-229
# --------------------------------------------------------------------------------------------
def segoutput_to_target(output: torch.Tensor, size=None) -> torch.Tensor:
    size = tuple(output.shape[-2:]) if size is None else tuple(size)
    am = output.argmax(dim=1, keepdim=True).type(torch.float)                        # plots.py:228
    return F.interpolate(am, size, mode="nearest").squeeze(dim=1)                    # plots.py:229


def nearest_src_index(n_in: int, n_out: int) -> np.ndarray:
    """ATen legacy 'nearest': ``src=min(floor(dst*float32(in/out)), in-1)``."""
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    return np.minimum(np.floor(dst * scale).astype(np.int64), n_in - 1)


def argmax_first_np(logits: np.ndarray) -> np.ndarray:
    """argmax over axis 1, lowest index on ties (torch semantics; NaN counts as maximal)."""
    v = np.asarray(logits, dtype=np.float32)
    best = v[:, 0]
    idx = np.zeros(best.shape, np.int64)
    for c in range(1, v.shape[1]):
        cur = v[:, c]
        better = (cur > best) | (np.isnan(cur) & ~np.isnan(best))
        idx = np.where(better, c, idx)
        best = np.where(better, cur, best)
    return idx


def segoutput_to_target_np(logits: np.ndarray, size=None) -> np.ndarray:
    am = argmax_first_np(logits).astype(np.float32)
    if size is None or tuple(size) == tuple(am.shape[-2:]):
        return am
    iy = nearest_src_index(am.shape[-2], size[0])
    ix = nearest_src_index(am.shape[-1], size[1])
    return am[:, iy][:, :, ix]


# --------------------------------------------------------------------------------------------
# a7 / a8  seg metrics                                                  metrics.py:350-388
# --------------------------------------------------------------------------------------------
def batch_pix_accuracy(output: torch.Tensor, target: torch.Tensor):
    predict = torch.max(output, 1)[1].cpu().numpy().astype("int32")                  # metrics.py:356-358
    tgt = target.cpu().numpy().astype("int32")
    labeled = np.sum(tgt > 0)                                                        # :361
    correct = np.sum((predict == tgt) * (tgt > 0))                                   # :362
    assert correct <= labeled
    return correct, labeled


def batch_intersection_union(output: torch.Tensor, target: torch.Tensor, nclass: int):
    predict = torch.max(output, 1)[1].cpu().numpy().astype("int32")                  # metrics.py:375-379
    tgt = target.cpu().numpy().astype("int32")
    inter_img = predict * (predict == tgt)                                           # :381
    rng = (1, nclass)
    nb = nclass - 1
    a_inter = np.histogram(inter_img, bins=nb, range=rng)[0]                         # :383
    a_pred = np.histogram(predict, bins=nb, range=rng)[0]                            # :384
    a_lab = np.histogram(tgt, bins=nb, range=rng)[0]                                 # :385
    a_union = a_pred + a_lab - a_inter
    assert (a_inter <= a_union).all()
    return a_inter, a_union


def cm_rows(n_channels: int, nclass: int) -> int:
    """Row count of the target-clamped confusion matrix used by the CUDA path: rows 0..R-2 are
    exact target values 0..max(nclass, C-1) (negatives fold into row 0: neither labeled nor a
    class), row R-1 collects every larger target (labeled, never a class nor a prediction)."""
    return max(nclass, n_channels - 1) + 2


def confusion_matrix(predict: np.ndarray, target: np.ndarray, n_channels: int, rows: int) -> np.ndarray:
    """CM[row(t)][p] over all pixels, int64; ``target`` is first wrapped to int32 exactly as
    metrics.py:359,380 do (``astype('int32')``), then clamped into [0, rows-1]."""
    t = np.asarray(target).astype("int32").astype(np.int64).reshape(-1)
    p = np.asarray(predict).astype(np.int64).reshape(-1)
    r = np.clip(t, 0, rows - 1)
    cm = np.zeros((rows, n_channels), np.int64)
    np.add.at(cm, (r, p), 1)
    return cm


def seg_stats_from_cm(cm: np.ndarray, nclass: int):
    """(correct, labeled, area_inter[nclass-1], area_union[nclass-1]) from CM — the closed form
    of metrics.py:350-388 incl. np.histogram's right-closed last bin (value==nclass folds into
    the last bin; class 0 and values > nclass are never counted)."""
    rows, C = cm.shape
    exact = rows - 1
    labeled = int(cm[1:].sum())
    correct = int(sum(cm[k, k] for k in range(1, min(exact, C))))
    nb = nclass - 1
    inter = np.zeros(nb, np.int64)
    pred = np.zeros(nb, np.int64)
    lab = np.zeros(nb, np.int64)
    for k in range(1, nclass + 1):
        b = min(k, nclass - 1) - 1
        if b < 0:
            continue
        if k < C:
            pred[b] += cm[:, k].sum()
            if k < exact:
                inter[b] += cm[k, k]
        if k < exact:
            lab[b] += cm[k].sum()
    return correct, labeled, inter, pred + lab - inter


# --------------------------------------------------------------------------------------------
# a9  seg_validation accumulation                                         val.py:58-76
# --------------------------------------------------------------------------------------------
def seg_validation_accumulate(batches, n_segcls: int) -> Tuple[float, float]:
    """``batches`` yields (full-res logits (B,C,H,W), target (B,H,W) int64).  Returns
    (pixAcc, mIoU) with the reference's float64 arithmetic and eps (val.py:71-73)."""
    tot_i = np.zeros(n_segcls - 1, np.int64)
    tot_u = np.zeros(n_segcls - 1, np.int64)
    tot_c = 0
    tot_l = 0
    for pred, target in batches:
        if tuple(pred.shape[-2:]) != tuple(target.shape[-2:]):
            pred = F.interpolate(pred, tuple(target.shape[-2:]), mode="bilinear", align_corners=False)  # val.py:47
        c, l = batch_pix_accuracy(pred, target)
        i, u = batch_intersection_union(pred, target, n_segcls)
        tot_c += c
        tot_l += l
        tot_i += i
        tot_u += u
    pix_acc = 1.0 * tot_c / (np.spacing(1) + tot_l)
    iou = 1.0 * tot_i / (np.spacing(1) + tot_u)
    return float(pix_acc), float(iou.mean())


# --------------------------------------------------------------------------------------------
# f1  scale_coords / clip_coords            general.py:598-623
# --------------------------------------------------------------------------------------------
def scale_coords(img1_shape, coords: torch.Tensor, img0_shape, ratio_pad=None, do_round: bool = False) -> torch.Tensor:
    """Restatement with the reference's own tensor ops (fp32 tensor - Python float, / Python float, clamp);
    returns a new tensor.  ``do_round`` adds detect.py:218's ``.round()``."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    c = coords.clone()
    c[:, 0] -= pad[0]
    c[:, 2] -= pad[0]
    c[:, 1] -= pad[1]
    c[:, 3] -= pad[1]
    c[:, :4] /= gain
    c[:, 0].clamp_(0, img0_shape[1])
    c[:, 1].clamp_(0, img0_shape[0])
    c[:, 2].clamp_(0, img0_shape[1])
    c[:, 3].clamp_(0, img0_shape[0])
    return c.round() if do_round else c


# --------------------------------------------------------------------------------------------
# f2  box_iou / process_batch     general.py:626-648, val.py:101-122
# --------------------------------------------------------------------------------------------
def box_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    def box_area(box):
        return (box[2] - box[0]) * (box[3] - box[1])
    area1 = box_area(box1.T)
    area2 = box_area(box2.T)
    inter = (torch.min(box1[:, None, 2:], box2[:, 2:]) - torch.max(box1[:, None, :2], box2[:, :2])).clamp(0).prod(2)
    return inter / (area1[:, None] + area2 - inter)


def process_batch(detections: torch.Tensor, labels: torch.Tensor, iouv: torch.Tensor) -> torch.Tensor:
    """Canonical restatement of val.py:101-122 without the numpy sort: per detection its best class-matching
    label with IoU >= iouv[0] (equal IoUs: lowest label index), then per label the first such detection in
    detection order, correct[d] = iou >= iouv."""
    n, m = detections.shape[0], labels.shape[0]
    correct = torch.zeros(n, iouv.shape[0], dtype=torch.bool)
    if n == 0 or m == 0:
        return correct
    iou = box_iou(labels[:, 1:], detections[:, :4])                     # (m, n)
    ok = (iou >= iouv[0]) & (labels[:, 0:1] == detections[:, 5])
    taken = set()
    for d in range(n):
        ls = torch.nonzero(ok[:, d]).flatten().tolist()
        if not ls:
            continue
        best = max(ls, key=lambda l: (float(iou[l, d]), -l))
        if best in taken:
            continue
        taken.add(best)
        correct[d] = iou[best, d] >= iouv
    return correct


# --------------------------------------------------------------------------------------------
# ap_per_class / compute_ap (metrics.py:31-121) and the detection ConfusionMatrix
# (metrics.py:124-170).  The float64 arithmetic of np.interp / np.trapz is restated operation by operation (plain
# Python floats = IEEE binary64, no FMA) so that the CUDA kernels have a bit-level specification; the restatement is
# pinned against fixtures generated by the reference itself (tests/golden/make_golden_stats.py).
# --------------------------------------------------------------------------------------------
def interp_scalar(x: float, xp, fp, left: float, right: float) -> float:
    """numpy's arr_interp for one x: j = last index with xp[j] <= x; slope form with the NaN retries."""
    n = len(xp)
    if x != x:
        return x
    if x > xp[n - 1]:
        return right
    if x < xp[0]:
        return left
    j = int(np.searchsorted(xp, x, side="right")) - 1
    if j == n - 1:
        return float(fp[j])
    if xp[j] == x:
        return float(fp[j])
    slope = (float(fp[j + 1]) - float(fp[j])) / (float(xp[j + 1]) - float(xp[j]))
    res = slope * (x - float(xp[j])) + float(fp[j])
    if res != res:
        res = slope * (x - float(xp[j + 1])) + float(fp[j + 1])
        if res != res and fp[j] == fp[j + 1]:
            res = float(fp[j])
    return res


def pairwise_sum_small(a) -> float:
    """numpy's pairwise add.reduce for a contiguous float64 vector of fewer than 128 elements."""
    n = len(a)
    a = [float(v) for v in a]
    if n < 8:
        res = 0.0 if n == 0 else a[0]          # numpy starts from the first element
        for v in a[1:]:
            res += v
        return res
    assert n <= 128
    r = a[:8]
    i = 8
    while i < n - (n % 8):
        for k in range(8):
            r[k] += a[i + k]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res


def compute_ap(recall, precision) -> float:
    """metrics.py:95-121 ('interp' method): sentinels, precision envelope, 101-point interpolation, trapezoid sum."""
    mrec = np.concatenate(([0.0], np.asarray(recall, dtype=np.float64), [1.0]))
    mpre = np.concatenate(([1.0], np.asarray(precision, dtype=np.float64), [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    x = np.linspace(0, 1, 101)
    y = [interp_scalar(float(v), mrec, mpre, float(mpre[0]), float(mpre[-1])) for v in x]
    terms = [(float(x[k + 1]) - float(x[k])) * (y[k + 1] + y[k]) / 2.0 for k in range(100)]
    return pairwise_sum_small(terms)


def ap_per_class(tp, conf, pred_cls, target_cls):
    """metrics.py:31-92 without the plots.  Equal confidences keep their input order (stable sort) where the
    reference's default quicksort leaves the order implementation-defined.  Returns (p, r, ap, f1, classes)."""
    tp = np.asarray(tp)
    tp = tp.reshape(len(conf), tp.shape[1] if tp.ndim == 2 else 1)
    order = np.argsort(-np.asarray(conf), kind="stable")
    tp, conf, pred_cls = tp[order], np.asarray(conf)[order], np.asarray(pred_cls)[order]
    classes = np.unique(target_cls)
    nc, niou = classes.shape[0], tp.shape[1]
    px = np.linspace(0, 1, 1000)
    ap, p, r = np.zeros((nc, niou)), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_l, n_p = int((np.asarray(target_cls) == c).sum()), int(sel.sum())
        if n_p == 0 or n_l == 0:
            continue
        tpc = tp[sel].astype(np.int64).cumsum(0)
        cnt = np.arange(1, n_p + 1, dtype=np.int64)[:, None]            # tpc + fpc
        recall = tpc / (n_l + 1e-16)
        precision = tpc / cnt
        xp = -conf[sel].astype(np.float64)
        r[ci] = [interp_scalar(-float(v), xp, recall[:, 0], 0.0, float(recall[-1, 0])) for v in px]
        p[ci] = [interp_scalar(-float(v), xp, precision[:, 0], 1.0, float(precision[-1, 0])) for v in px]
        for j in range(niou):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])
    f1 = 2 * p * r / (p + r + 1e-16)
    i = f1.mean(0).argmax()
    return p[:, i], r[:, i], ap, f1[:, i], classes.astype("int32")


def det_confusion_matrix(matrix: np.ndarray, detections: torch.Tensor, labels: torch.Tensor, nc: int, conf: float = 0.25,
                         iou_thres: float = 0.45) -> np.ndarray:
    """Canonical restatement of ConfusionMatrix.process_batch (metrics.py:132-170) without the numpy sorts: every
    detection above `conf` picks its best-IoU label (IoU > iou_thres; ties: lowest label), every label keeps its
    best-IoU such detection (ties: lowest detection); matched label -> matrix[det class, gt class], unmatched label
    -> matrix[nc, gt class]; detections left over -> matrix[det class, nc], but only if the image had any match
    (metrics.py:165-168).  Updates and returns `matrix` ((nc+1, nc+1), rows = predicted)."""
    det = detections[detections[:, 4] > conf]
    gt = labels[:, 0].int().tolist()
    dc = det[:, 5].int().tolist()
    n, m = det.shape[0], labels.shape[0]
    match_of_label = {}
    if n and m:
        iou = box_iou(labels[:, 1:], det[:, :4])
        thr = torch.tensor(iou_thres, dtype=torch.float32)
        best = {}
        for d in range(n):
            cand = [(float(iou[l, d]), -l) for l in range(m) if iou[l, d] > thr]
            if cand:
                v, negl = max(cand)
                best[d] = (-negl, v)
        for d, (l, v) in best.items():
            if l not in match_of_label or (v, -d) > (match_of_label[l][1], -match_of_label[l][0]):
                match_of_label[l] = (d, v)
    any_match = len(match_of_label) > 0
    for l, gc in enumerate(gt):
        if l in match_of_label:
            matrix[dc[match_of_label[l][0]], gc] += 1
        else:
            matrix[nc, gc] += 1
    if any_match:
        used = {d for d, _ in match_of_label.values()}
        for d, c in enumerate(dc):
            if d not in used:
                matrix[c, nc] += 1
    return matrix


# --------------------------------------------------------------------------------------------
# Synthetic inputs: family T (throughput) — seeded, device-independent (CPU RNG)
# --------------------------------------------------------------------------------------------
def synth_head_outputs(bs: int, S: int, nc: int, C: int = 2, seed: int = 0, obj_mean: float = -4.0,
                       cls_mean: float = -2.0, na: int = 3, strides=(8, 16, 32), with_target: bool = True):
    """Raw head logits per level (bs, na*no, S/s, S/s): box ch ~N(0,1), obj ~N(obj_mean,1.5^2),
    cls ~N(cls_mean,1.5^2); seg logits (bs,C,S/8,S/8) ~N(0,1); targets randint(0,C) int64."""
    g = torch.Generator().manual_seed(seed)
    no = nc + 5
    levels = []
    for s in strides:
        n = S // s
        t = torch.randn(bs, na, no, n, n, generator=g)
        t[:, :, 4] = t[:, :, 4] * 1.5 + obj_mean
        t[:, :, 5:] = t[:, :, 5:] * 1.5 + cls_mean
        levels.append(t.reshape(bs, na * no, n, n).contiguous())
    seg = torch.randn(bs, C, S // 8, S // 8, generator=g)
    target = torch.randint(0, C, (bs, S, S), generator=g, dtype=torch.int64) if with_target else None
    return levels, seg, target
