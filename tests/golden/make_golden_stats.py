"""Generates tests/golden/det_stats.npz by running the UNMODIFIED reference's ap_per_class / compute_ap and
ConfusionMatrix.process_batch (core/utils/metrics.py:31-121, 124-170), imported in place.  Build container only:
    python tests/golden/make_golden_stats.py
Confidences and IoUs are unique random fp32 values, so the reference's unstable numpy sorts never see ties."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
from oracle import ref_shim  # noqa: E402
from make_golden_match import make_case  # noqa: E402


def make_stats(rng, n, n_targets, nc, niou=10, present=None):
    """(tp, conf, pred_cls, target_cls) shaped like val.py:263,282: bool (n, niou), f32 (n,), f32 (n,), f32 (m,)."""
    present = list(range(nc)) if present is None else present
    conf = rng.permutation(np.linspace(0.001, 0.999, n, dtype=np.float64)).astype(np.float32)
    assert np.unique(conf).size == n
    pred_cls = rng.choice(np.arange(nc + 1), size=n).astype(np.float32)         # class nc never appears in the targets
    target_cls = rng.choice(np.asarray(present), size=n_targets).astype(np.float32)
    # true positives: likelier at high confidence, monotone over the IoU thresholds (correct[:, j] implies [:, j-1]);
    # per class and threshold never more TPs than labels (one label matches at most one detection)
    level = (rng.random(n) < conf * 0.9) * rng.integers(1, niou + 1, size=n)
    tp = np.arange(niou)[None, :] < level[:, None]
    for c in np.unique(pred_cls):
        n_l = int((target_cls == c).sum())
        idx = np.nonzero(pred_cls == c)[0]
        idx = idx[np.argsort(-conf[idx])]
        for j in range(niou):
            on = idx[tp[idx, j]]
            tp[on[n_l:], j:] = False
    return tp, conf, pred_cls, target_cls


def main():
    ref_shim.load()
    from core.utils import metrics as ref_metrics
    rng = np.random.default_rng(11)
    arrs = {}
    cases = [(400, 60, 3, None), (5000, 900, 6, None), (37, 10, 2, [1]), (20000, 2500, 80, list(range(0, 80, 3))),
             (3, 2, 2, None), (64, 200, 4, [0, 1, 2, 3])]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, (n, m, nc, present) in enumerate(cases):
            tp, conf, pred_cls, target_cls = make_stats(rng, n, m, nc, present=present)
            p, r, ap, f1, cls = ref_metrics.ap_per_class(tp, conf, pred_cls, target_cls, plot=False, names={})
            arrs.update({f"tp{k}": tp, f"conf{k}": conf, f"pred_cls{k}": pred_cls, f"target_cls{k}": target_cls,
                         f"p{k}": p, f"r{k}": r, f"ap{k}": ap, f"f1{k}": f1, f"cls{k}": cls})
    arrs["n_cases"] = np.int64(len(cases))
    # compute_ap alone on one curve
    rec = np.sort(rng.random(50))
    pre = np.sort(rng.random(50))[::-1].copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a, mpre, mrec = ref_metrics.compute_ap(rec, pre)
    arrs.update({"ca_recall": rec, "ca_precision": pre, "ca_ap": np.float64(a), "ca_mpre": mpre, "ca_mrec": mrec})
    # detection confusion matrix over a sequence of images
    g = torch.Generator().manual_seed(13)
    cm_cases = [(6, [(12, 40), (1, 5), (30, 300), (7, 1), (0, 9), (9, 0)]), (80, [(50, 120), (20, 300)])]
    for k, (nc, imgs) in enumerate(cm_cases):
        cm = ref_metrics.ConfusionMatrix(nc=nc, conf=0.25, iou_thres=0.45)
        for q, (m, n) in enumerate(imgs):
            if m == 0:
                det, _ = make_case(g, 3, n, nc)
                labels = torch.zeros(0, 5)
            else:
                det, labels = make_case(g, m, n, nc)
            arrs[f"cm{k}_det{q}"] = det.numpy()
            arrs[f"cm{k}_labels{q}"] = labels.numpy()
            cm.process_batch(det, labels)
        arrs[f"cm{k}_matrix"] = cm.matrix().copy()
        arrs[f"cm{k}_nc"] = np.int64(nc)
        arrs[f"cm{k}_n"] = np.int64(len(imgs))
    arrs["n_cm"] = np.int64(len(cm_cases))
    np.savez_compressed(os.path.join(HERE, "det_stats.npz"), **arrs)
    print("det_stats.npz written; mAP@.5 per case:", [float(arrs[f"ap{k}"][:, 0].mean()) for k in range(len(cases))],
          "cm sums:", [float(arrs[f"cm{k}_matrix"].sum()) for k in range(len(cm_cases))])


if __name__ == "__main__":
    main()
