"""DetEvaluator: the detection-metric half of ``val.run`` (scripts/val.py:227-300) with everything on the device.

The reference walks the images of a batch in Python, copies every image's detections to the host for
``process_batch`` and collects numpy ``stats``.  Here one batch is four launches on the NMS output buffers
(``PostProcessor.det`` / ``det_count``) and nothing is synchronised until ``compute()``:

    labels_to_native   val.py:227,243-255   all labels of the batch: pixels, xywh2xyxy, native-space rescale + clip
    scale_coords_      val.py:250-251       detections -> native space (on a copy: ``pred`` itself stays as it is)
    process_batch      val.py:257           (bs, rows, niou) `correct` flags
    det_confusion      val.py:258-259       optional (nc+1, nc+1) confusion matrix (metrics.py:132-170)

``compute()`` hands the concatenated statistics to ``ap_per_class`` (device sort / scans / float64 integration)
and derives mp, mr, map50, map and the per-class table exactly as val.py:281-287 does."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import api, ops


class DetEvaluator:
    def __init__(self, nc: int, iouv: Optional[torch.Tensor] = None, device=None, confusion_matrix: bool = False,
                 conf: float = 0.25, iou_thres: float = 0.45, single_cls: bool = False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nc = int(nc)
        self.iouv = (torch.linspace(0.5, 0.95, 10) if iouv is None else iouv).to(self.device, torch.float32)   # val.py:186
        self.niou = int(self.iouv.numel())
        self.single_cls = single_cls
        self.cm = api.ConfusionMatrix(nc, conf=conf, iou_thres=iou_thres, device=self.device) if confusion_matrix else None
        self.seen = 0
        self._chunks = []          # (correct (bs, rows, niou) u8, det (bs, rows, 6), det_count (bs))
        self._tcls = []

    def update(self, det: torch.Tensor, det_count: torch.Tensor, targets: torch.Tensor, img_hw, params: torch.Tensor,
               max_labels_per_image: Optional[int] = None):
        """One batch.  ``det`` (bs, rows, 6) / ``det_count`` (bs): NMS output in network-input pixels; ``targets``
        (n, 6) [image, cls, normalised xywh] sorted by image (as the collate function builds them, val.py:213);
        ``img_hw`` = (height, width) of the network input; ``params`` (bs, 5) from ``ops.letterbox_params`` per image
        (shapes[si] in val.py:239,250).  ``max_labels_per_image``: an upper bound on the labels of one image (sizes the
        kernels' shared memory); by default the true maximum of this batch — computed on the host for free when
        ``targets`` arrives as a CPU tensor (the dataloader's collate output), with one sync when it is already on the
        device."""
        bs, rows = int(det.shape[0]), int(det.shape[1])
        height, width = float(img_hw[0]), float(img_hw[1])
        targets = targets.reshape(-1, 6)
        if max_labels_per_image is None:
            max_labels_per_image = int(torch.bincount(targets[:, 0].long(), minlength=1).max()) if targets.shape[0] else 0
        targets = targets.to(self.device, torch.float32)
        n_lab = max(int(max_labels_per_image), 1)
        labelsn = ops.labels_to_native(targets, width, height, params)
        counts = torch.bincount(targets[:, 0].long(), minlength=bs)[:bs]
        label_off = torch.zeros(bs + 1, dtype=torch.int32, device=self.device)
        label_off[1:] = torch.cumsum(counts, 0)
        pred = det.clone() if not self.single_cls else det.clone().index_fill_(2, torch.tensor([5], device=det.device), 0.0)
        predn = pred.clone()                                                   # val.py:249
        ops.scale_coords_(predn, params, det_count)                            # val.py:250-251
        correct = ops.process_batch(predn, det_count, labelsn, label_off, self.iouv, n_lab)   # val.py:257
        if self.cm is not None:
            self.cm.process_images(predn, det_count, labelsn, label_off, n_lab, skip_empty=True)   # val.py:259
        self._chunks.append((correct, pred, det_count.clone()))
        self._tcls.append(targets[:, 1].clone())
        self.seen += bs
        return correct

    def stats(self):
        """The four arrays val.py:282 concatenates: (tp (n, niou) bool, conf (n), pred_cls (n), target_cls (m)) as
        CUDA tensors (this is where the padded per-batch buffers are compacted; the only host sync)."""
        tp, conf, pcls = [], [], []
        for correct, det, cnt in self._chunks:
            rows = det.shape[1]
            valid = torch.arange(rows, device=self.device)[None, :] < cnt[:, None]
            tp.append(correct[valid].bool())
            conf.append(det[..., 4][valid])
            pcls.append(det[..., 5][valid])
        cat = lambda xs, shape, dt: torch.cat(xs) if xs else torch.zeros(shape, dtype=dt, device=self.device)
        return (cat(tp, (0, self.niou), torch.bool), cat(conf, (0,), torch.float32), cat(pcls, (0,), torch.float32),
                cat(self._tcls, (0,), torch.float32))

    def compute(self):
        """val.py:281-300: ``(mp, mr, map50, map), maps, (p, r, ap50, ap, ap_class, nt)``."""
        tp, conf, pcls, tcls = self.stats()
        mp = mr = map50 = map_ = 0.0
        p = r = ap50 = ap = np.zeros(0)
        ap_class = np.zeros(0, dtype="int32")
        if tp.numel() and bool(tp.any()):                                      # val.py:283
            p, r, ap, f1, ap_class = api.ap_per_class(tp, conf, pcls, tcls)
            ap50, ap = ap[:, 0], ap.mean(1)                                    # val.py:285
            mp, mr, map50, map_ = p.mean(), r.mean(), ap50.mean(), ap.mean()   # val.py:286
        nt = np.bincount(tcls.cpu().numpy().astype(np.int64), minlength=self.nc)   # val.py:287
        maps = np.zeros(self.nc) + map_                                        # val.py:342-344
        for i, c in enumerate(ap_class):
            maps[c] = ap[i]
        return (mp, mr, map50, map_), maps, (p, r, ap50, ap, ap_class, nt)
