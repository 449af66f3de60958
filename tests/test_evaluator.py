"""GPU (-m gpu): DetEvaluator — the detection-metric half of val.run (scripts/val.py:227-300) on the device —
against the same loop composed from the oracle's per-image functions (scale_coords, process_batch, ap_per_class,
detection confusion matrix): `correct`, the statistics and mp / mr / mAP@.5 / mAP@.5:.95 are bit-identical."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _batch(g, bs, rows, nc, S):
    """NMS-style padded detections in network-input pixels + collated normalised targets."""
    det = torch.zeros(bs, rows, 6)
    cnt = torch.zeros(bs, dtype=torch.int32)
    targets = []
    for i in range(bs):
        m = int(torch.randint(0, 9, (1,), generator=g))
        n = int(torch.randint(0, rows + 1, (1,), generator=g)) if i != 1 else 0            # image 1: no detections
        xy = 60 + torch.rand(m, 2, generator=g) * (S - 120)
        wh = 30 + torch.rand(m, 2, generator=g) * 90
        cls = torch.randint(0, nc, (m, 1), generator=g).float()
        if m:
            targets.append(torch.cat((torch.full((m, 1), float(i)), cls, xy / S, wh / S), 1))
        if n:
            if m:
                src = torch.randint(0, m, (n,), generator=g)
                box = torch.cat((xy[src] - wh[src] / 2, xy[src] + wh[src] / 2), 1) + torch.randn(n, 4, generator=g) * 6
                c = cls[src, 0].clone()
                flip = torch.rand(n, generator=g) < 0.15
                c[flip] = torch.randint(0, nc, (int(flip.sum()),), generator=g).float()
            else:
                p = torch.rand(n, 2, generator=g) * (S - 80)
                box = torch.cat((p, p + 20 + torch.rand(n, 2, generator=g) * 60), 1)
                c = torch.randint(0, nc, (n,), generator=g).float()
            conf = torch.sort(0.05 + 0.9 * torch.rand(n, generator=g), descending=True).values
            det[i, :n] = torch.cat((box, conf[:, None], c[:, None]), 1)
        cnt[i] = n
    targets = torch.cat(targets) if targets else torch.zeros(0, 6)
    return det, cnt, targets


def test_det_evaluator_matches_oracle_loop():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from desenet_b200 import ops
    from desenet_b200.evaluator import DetEvaluator
    g = torch.Generator().manual_seed(77)
    S, nc, bs, rows = 320, 4, 5, 40
    iouv = torch.linspace(0.5, 0.95, 10)
    ev = DetEvaluator(nc, confusion_matrix=True)
    stats, cm_want = [], np.zeros((nc + 1, nc + 1))
    shapes = [((240, 320), None), ((333, 517), None), ((640, 480), ((0.5, 0.5), (40.0, 0.0))), ((320, 320), None), ((200, 300), None)]
    for _ in range(3):
        det, cnt, targets = _batch(g, bs, rows, nc, S)
        params = torch.tensor([ops.letterbox_params((S, S), s0, rp) for s0, rp in shapes], dtype=torch.float32).cuda()
        correct = ev.update(det.cuda(), cnt.cuda(), targets.cuda(), (S, S), params).cpu()
        # ---- the reference loop (val.py:232-263) from the oracle's per-image functions ----
        tpix = targets.clone()
        tpix[:, 2:] *= torch.tensor([S, S, S, S], dtype=torch.float32)
        for si in range(bs):
            labels = tpix[tpix[:, 0] == si, 1:]
            n = int(cnt[si])
            pred = det[si, :n]
            tcls = labels[:, 0].tolist()
            if n == 0:
                if len(labels):
                    stats.append((np.zeros((0, 10), dtype=bool), np.zeros(0, np.float32), np.zeros(0, np.float32), tcls))
                continue
            predn = pred.clone()
            predn[:, :4] = port.scale_coords((S, S), predn[:, :4], shapes[si][0], shapes[si][1])
            if len(labels):
                tbox = port.scale_coords((S, S), port.xywh2xyxy(labels[:, 1:5]), shapes[si][0], shapes[si][1])
                labelsn = torch.cat((labels[:, 0:1], tbox), 1)
                want = port.process_batch(predn, labelsn, iouv)
                port.det_confusion_matrix(cm_want, predn, labelsn, nc)
            else:
                want = torch.zeros(n, 10, dtype=torch.bool)
            assert torch.equal(correct[si, :n].bool(), want), si
            stats.append((want.numpy(), pred[:, 4].numpy(), pred[:, 5].numpy(), tcls))
    tp, conf, pcls, tcls = [np.concatenate(x, 0) for x in zip(*stats)]
    got = [t.cpu().numpy() for t in ev.stats()]
    assert np.array_equal(got[0], tp) and np.array_equal(got[1], conf) and np.array_equal(got[2], pcls)
    assert np.array_equal(np.sort(got[3]), np.sort(np.asarray(tcls, dtype=np.float32)))
    p, r, ap, f1, ap_class = port.ap_per_class(tp, conf, pcls, np.asarray(tcls))
    (mp, mr, map50, map_), maps, (p2, r2, ap50, apm, cls2, nt) = ev.compute()
    assert (mp, mr, map50, map_) == (p.mean(), r.mean(), ap[:, 0].mean(), ap.mean(1).mean())
    assert np.array_equal(p2, p) and np.array_equal(r2, r) and np.array_equal(apm, ap.mean(1)) and np.array_equal(cls2, ap_class)
    assert np.array_equal(nt, np.bincount(np.asarray(tcls).astype(np.int64), minlength=nc))
    assert map50 > 0.2                                     # the synthetic detections really do match labels
    assert ev.seen == 15
    # detection confusion matrix: val.py:258-259 feeds it the images that have both labels and detections
    assert np.array_equal(ev.cm.matrix(), cm_want) and cm_want.sum() > 0
