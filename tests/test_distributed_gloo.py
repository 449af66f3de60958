"""CPU (-m "not gpu"): the multi-GPU host logic on 2 gloo ranks — contiguous image sharding, the int64
confusion-matrix all-reduce and the detection gather restore exactly the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from desenet_b200 import distributed as dd
from oracle import port


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 3, 8):
            spans = [dd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port_no, n_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)            # every rank builds the same global batch
        max_det, C, rows = 5, 2, 4
        det = torch.rand(n_images, max_det, 6, generator=g)
        cnt = torch.randint(0, max_det + 1, (n_images,), generator=g, dtype=torch.int32)
        pred = torch.randint(0, C, (n_images, 8, 8), generator=g)
        tgt = torch.randint(-1, 4, (n_images, 8, 8), generator=g)
        lo, hi = dd.shard_range(n_images, rank, world)
        sizes = [b - a for a, b in (dd.shard_range(n_images, r, world) for r in range(world))]
        # confusion matrix of the local shard (slots, rows, C) -> all-reduce
        cm_local = torch.from_numpy(np.stack([port.confusion_matrix(pred[i].numpy(), tgt[i].numpy(), C, rows)
                                              for i in range(lo, hi)])) if hi > lo else torch.zeros(1, rows, C, dtype=torch.int64)
        total = dd.allreduce_confusion_matrix(cm_local)
        dets = dd.gather_detections(det[lo:hi].clone(), cnt[lo:hi].clone(), images_per_rank=sizes)
        want_cm = port.confusion_matrix(pred.numpy(), tgt.numpy(), C, rows)
        ok = np.array_equal(total.numpy(), want_cm) and len(dets) == n_images
        for i in range(n_images):
            ok = ok and torch.equal(dets[i], det[i, :int(cnt[i])])
        ok = ok and dd.max_over_ranks(float(rank)) == float(world - 1)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_allreduce():
    ctx = mp.get_context("spawn")
    for n_images in (8, 7):                              # even and uneven shards
        q = ctx.Queue()
        port_no = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port_no, n_images, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert res == [(0, True), (1, True)], res


def test_single_process_paths():
    det = torch.rand(3, 4, 6)
    cnt = torch.tensor([0, 4, 2], dtype=torch.int32)
    out = dd.gather_detections(det, cnt)
    assert [o.shape[0] for o in out] == [0, 4, 2]
    cm = torch.ones(5, 4, 2, dtype=torch.int64)
    assert torch.equal(dd.allreduce_confusion_matrix(cm), torch.full((4, 2), 5))
    assert dd.max_over_ranks(3.5) == 3.5
