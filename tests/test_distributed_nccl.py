"""GPU, needs >= 2 devices (skipped otherwise; run with `gpurun --gpus 2`): the sharded CUDA path on 2 NCCL ranks returns
exactly the unsharded result — detections gathered over NCCL (`distributed.gather_detections`) and over peer memory
(`peer.PeerGather`: the NMS kernels write into rank 0's buffer), confusion matrix all-reduced (val.py:67-73 totals) —
and both equal the oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S, NC, C, N_IMAGES = 160, 6, 2, 10
KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_shard(dev, lo, hi, det_out=None, cnt_out=None, gather=None):
    from desenet_b200.pipeline import PostProcessor
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_images
    levels, seg, tgt = synth_images(lo, hi - lo, S, NC, C, seed=5)
    pp = PostProcessor(hi - lo, S, NC, C, DEFAULT_ANCHORS, DEFAULT_STRIDES, mode="eval", nclass=2, device=dev, det_out=det_out,
                       det_count_out=cnt_out, gather=gather, **KW)
    pp.run([l.to(dev) for l in levels], seg.to(dev), tgt.to(dev))
    torch.cuda.synchronize(dev)
    return pp


def _worker(rank, world, port_no, q):
    import torch.distributed as dist

    from desenet_b200 import distributed as dd
    from desenet_b200 import peer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = dd.shard_range(N_IMAGES, rank, world)
        sizes = [b - a for a, b in (dd.shard_range(N_IMAGES, r, world) for r in range(world))]
        # (1) local buffers + NCCL gather / all-reduce
        pp = _run_shard(dev, lo, hi)
        dets = dd.gather_detections(pp.det, pp.det_count, images_per_rank=sizes)
        cm = dd.allreduce_confusion_matrix(pp.cm)
        # (2) the same shard, the NMS kernel writing straight into rank 0's gathered buffer
        g = peer.PeerGather(N_IMAGES, KW["max_det"], 1, dev)
        _run_shard(dev, lo, hi, det_out=g.det_view(0, lo, hi), cnt_out=g.cnt_view(0, lo, hi), gather=(g, 0))
        ok, msg = True, ""
        if rank == 0:
            g.wait(0, torch.ones(world, dtype=torch.int32, device=dev))
            torch.cuda.synchronize(dev)
            g.check()
            pdets = [d.clone() for d in g.detections(0)]
            # unsharded run on this GPU alone
            full = _run_shard(dev, 0, N_IMAGES)
            n = full.det_count.cpu().tolist()
            for i in range(N_IMAGES):
                want = full.det[i, :n[i]]
                if not (torch.equal(dets[i], want) and torch.equal(pdets[i], want)):
                    ok, msg = False, f"image {i}: sharded detections differ from the unsharded run"
            if not torch.equal(cm, full.cm.sum(0)):
                ok, msg = False, "all-reduced confusion matrix differs from the unsharded run"
            q.put((n, [d.cpu().numpy() for d in dets], cm.cpu().numpy()))
        g.close()
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_equals_unsharded_and_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    import desenet_b200 as dsn
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_images
    from oracle import cport, port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(3)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    payload = [g for g in got if len(g) == 3 and isinstance(g[0], list)][0]
    status = [g for g in got if not isinstance(g[0], list)]
    assert all(s[1] for s in status), [s[2] for s in status]
    counts, dets, cm = payload
    # oracle on the unsharded batch: NMS on the kernel's own decoded prediction, confusion matrix from the C restatement
    levels, seg, tgt = synth_images(0, N_IMAGES, S, NC, C, seed=5)
    z = dsn.HeadOutputs([l.cuda() for l in levels], DEFAULT_ANCHORS, DEFAULT_STRIDES).z.cpu()
    want = port.non_max_suppression(z, KW["conf_thres"], KW["iou_thres"], multi_label=True, max_det=KW["max_det"], nms_impl="c")
    for i in range(N_IMAGES):
        assert np.array_equal(dets[i], want[i].numpy()), f"image {i}"
    full = cport.upsample_bilinear_ac(seg.numpy(), S, S)
    ref_cm = port.confusion_matrix(cport.argmax_first(full), tgt.numpy(), C, dsn.ops.cm_rows_for(C, 2))
    assert np.array_equal(cm, ref_cm)
