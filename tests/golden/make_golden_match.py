"""Generates tests/golden/process_batch.npz by running the UNMODIFIED reference's process_batch (scripts/val.py:
101-122) and box_iou (core/utils/general.py:626-648), imported in place.  Build container only:
    python tests/golden/make_golden_match.py
Inputs have no exactly tied IoUs (random fp32 boxes), so the reference's numpy-sort tie order never matters."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_shim  # noqa: E402


def make_case(g, n_lab, n_det, nc, S=640.0):
    lab_xy = torch.rand(n_lab, 2, generator=g) * S
    lab_wh = 30 + torch.rand(n_lab, 2, generator=g) * 120
    labels = torch.cat((torch.randint(0, nc, (n_lab, 1), generator=g).float(), lab_xy - lab_wh / 2, lab_xy + lab_wh / 2), 1)
    # detections: jittered copies of labels (several per label, some with the wrong class) + random boxes
    src = torch.randint(0, max(n_lab, 1), (n_det,), generator=g)
    jit = torch.randn(n_det, 4, generator=g) * 12
    boxes = (labels[src, 1:] + jit) if n_lab else torch.rand(n_det, 4, generator=g) * S
    cls = labels[src, 0].clone() if n_lab else torch.zeros(n_det)
    flip = torch.rand(n_det, generator=g) < 0.2
    cls[flip] = torch.randint(0, nc, (int(flip.sum()),), generator=g).float()
    conf = torch.sort(torch.rand(n_det, generator=g), descending=True).values
    det = torch.cat((boxes, conf[:, None], cls[:, None]), 1)
    return det.float(), labels.float()


def main():
    ref_shim.load()
    with ref_shim.in_reference_cwd():
        sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "scripts"))
        import scripts.val as ref_val
    from core.utils.general import box_iou
    g = torch.Generator().manual_seed(9)
    iouv = torch.linspace(0.5, 0.95, 10)
    arrs = {"iouv": iouv.numpy()}
    cases = [(12, 40, 3), (1, 5, 2), (30, 300, 6), (7, 1, 2), (50, 120, 80)]
    for k, (m, n, nc) in enumerate(cases):
        det, labels = make_case(g, m, n, nc)
        arrs[f"det{k}"] = det.numpy()
        arrs[f"labels{k}"] = labels.numpy()
        arrs[f"correct{k}"] = ref_val.process_batch(det, labels, iouv).numpy()
        arrs[f"iou{k}"] = box_iou(labels[:, 1:], det[:, :4]).numpy()
    arrs["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "process_batch.npz"), **arrs)
    print("process_batch.npz written", [int(arrs[f"correct{k}"].sum()) for k in range(len(cases))])


if __name__ == "__main__":
    main()
