"""GPU (-m gpu): the exact objects bench.py drives — PostProcessor (two-kernel and fused detection paths, direct
launches and CUDA-graph replay, side-stream overlap) and HostPipeline (pinned host buffers in and out) — give the
oracle's results."""
import numpy as np
import pytest
import torch

from oracle import cport, port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from desenet_b200 import api, ops, pipeline
    return api, ops, pipeline


def _oracle(levels, seg, tgt, kw, S, nclass):
    z, _ = port.decode_levels(levels, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, levels[0].shape[1] // 3 - 5)
    full = cport.upsample_bilinear_ac(seg.numpy(), S, S)
    pred = cport.argmax_first(full)
    cm = cport.confusion_matrix(pred, tgt.numpy(), seg.shape[1], port.cm_rows(seg.shape[1], nclass)) if tgt is not None else None
    return z, pred, cm


@pytest.mark.parametrize("fused_det", [False, True])
@pytest.mark.parametrize("mode", ["detect", "eval"])
def test_postprocessor_matches_oracle(env, fused_det, mode):
    api, ops, pipeline = env
    bs, S, nc, C = 6, 320, 6, 2
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=300) if mode == "detect" else \
        dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
    levels, seg, tgt = port.synth_head_outputs(bs, S, nc, C, seed=21)
    dl = [l.cuda() for l in levels]
    pp = pipeline.PostProcessor(bs, S, nc, C, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode=mode, nclass=2,
                                fused_det=fused_det, **kw)
    pp.run(dl, seg.cuda(), tgt.cuda() if mode == "eval" else None)
    torch.cuda.synchronize()
    # detections: bit-exact vs the oracle on the kernel's own decode
    ho = api.HeadOutputs(dl, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
    want = port.non_max_suppression(ho.z.cpu(), nms_impl="c", **kw)
    got = pp.detections()
    for g, w in zip(got, want):
        assert torch.equal(g.cpu(), w)
    _, pred, cm = _oracle(levels, seg, tgt if mode == "eval" else None, kw, S, 2)
    if mode == "detect":
        assert np.array_equal(pp.mask.cpu().numpy(), pred.astype(np.float32))
    else:
        assert np.array_equal(pp.cm.sum(0).cpu().numpy(), cm)
    # CUDA-graph replay of the same step (fresh outputs) gives the same results; eval accumulates
    first_det = pp.det.clone()
    if mode == "detect":
        pp.mask.zero_()
    pp.det.zero_()
    pp.capture(dl, seg.cuda(), tgt.cuda() if mode == "eval" else None)
    cm_before = pp.cm.sum(0).clone() if mode == "eval" else None
    pp.replay()
    torch.cuda.synchronize()
    n = pp.det_count.cpu().tolist()
    for i in range(bs):
        assert torch.equal(pp.det[i, :n[i]], first_det[i, :n[i]])
    if mode == "detect":
        assert np.array_equal(pp.mask.cpu().numpy(), pred.astype(np.float32))
    else:
        assert np.array_equal((pp.cm.sum(0) - cm_before).cpu().numpy(), cm)


def test_host_pipeline_round_trip(env):
    api, ops, pipeline = env
    bs, S, nc, C = 4, 160, 6, 2
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=100)
    sets = [port.synth_head_outputs(bs, S, nc, C, seed=30 + k, with_target=False) for k in range(3)]

    def make_pp():
        return pipeline.PostProcessor(bs, S, nc, C, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode="detect", **kw)

    hp = pipeline.HostPipeline(make_pp, [tuple(l.shape) for l in sets[0][0]], tuple(sets[0][1].shape), None, depth=2)
    assert hp.h2d_bytes == sum(l.numel() * 4 for l in sets[0][0]) + sets[0][1].numel() * 4
    for levels, seg, _ in sets:                     # 3 steps over 2 slots: slot 0 is reused
        slot = hp.submit([l.pin_memory() for l in levels], seg.pin_memory())
        slot["done"].synchronize()
        z, pred, _ = _oracle(levels, seg, None, kw, S, 2)
        ho = api.HeadOutputs([l.cuda() for l in levels], port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
        want = port.non_max_suppression(ho.z.cpu(), nms_impl="c", **kw)
        cnt = slot["h_cnt"].tolist()
        for i in range(bs):
            assert torch.equal(slot["h_det"][i, :cnt[i]], want[i])
        assert np.array_equal(slot["h_mask"].numpy(), pred.astype(np.float32))
    hp.synchronize()


def test_postprocessor_scale_option(env):
    api, ops, pipeline = env
    bs, S, nc, C = 2, 160, 6, 2
    levels, seg, _ = port.synth_head_outputs(bs, S, nc, C, seed=40, with_target=False)
    params = torch.tensor([ops.letterbox_params((S, S), (333, 517))] * bs, dtype=torch.float32).cuda()
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=100)
    pp = pipeline.PostProcessor(bs, S, nc, C, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode="det_only",
                                scale_params=params, round_boxes=True, **kw)
    dl = [l.cuda() for l in levels]
    pp.run(dl)
    ho = api.HeadOutputs(dl, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
    want = port.non_max_suppression(ho.z.cpu(), nms_impl="c", **kw)
    for g, w in zip(pp.detections(), want):
        w = w.clone()
        w[:, :4] = port.scale_coords((S, S), w[:, :4], (333, 517), do_round=True)
        assert torch.equal(g.cpu(), w)


def test_conf_cut_is_exact(env):
    """Sampled confidence cut (dsn_conf_cut_levels + dsn_filter_levels_cut): far fewer keys are emitted, the
    detections are identical to the run without a cut and to the oracle (> max_nms candidates per image)."""
    api, ops, pipeline = env
    bs, S, nc, C = 2, 256, 80, 2
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
    levels, _, _ = port.synth_head_outputs(bs, S, nc, C, seed=50, with_target=False)
    dl = [l.cuda() for l in levels]
    mk = lambda cut: pipeline.PostProcessor(bs, S, nc, C, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode="det_only",
                                            conf_cut=cut, **kw)
    a, b = mk(False), mk(None)
    assert b.conf_cut and not a.conf_cut            # auto: multi-label and cap = A*nc >= 8 * cut_target
    a.run(dl)
    b.run(dl)
    torch.cuda.synchronize()
    ca, cb = a.counts.cpu(), b.counts.cpu()
    assert (ca > 2 * ops.MAX_NMS).all() and (cb >= b.cut_target).all() and (cb < 4 * b.cut_target).all()
    assert (b.det_count.cpu() >= 0).all()
    for x, y in zip(a.detections(), b.detections()):
        assert torch.equal(x, y)
    ho = api.HeadOutputs(dl, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
    want = port.non_max_suppression(ho.z[:1].cpu(), nms_impl="c", **kw)
    assert torch.equal(b.detections()[0].cpu(), want[0])
    got = api.non_max_suppression(ho, **kw)               # the reference-facing call takes the cut path as well
    for x, y in zip(got, a.detections()):
        assert torch.equal(x, y)
    # graph replay with the cut
    first = b.det.clone()
    b.det.zero_()
    b.capture(dl)
    b.replay()
    torch.cuda.synchronize()
    assert torch.equal(b.det, first)


def test_conf_cut_fallback_when_sample_misleads(env):
    """Adversarial input: only the SAMPLED cells (every 64th unit of 4 cells) are confident, so the histogram
    over-estimates 64x and the cut lands above almost everything; with agnostic NMS the few emitted candidates
    (same box for all classes of a cell) cannot give max_det survivors, the NMS runs out of emitted candidates
    -> det_count -2 after dsn_nms_levels -> the on-device redo (dsn_detect_nms_redo_levels, part of the same enqueue and
    of a captured graph) reruns those images without the cut: no negative count ever reaches a consumer."""
    api, ops, pipeline = env
    bs, S, nc = 2, 256, 80
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300, agnostic=True)
    g = torch.Generator().manual_seed(7)
    levels = []
    for s in (8, 16, 32):
        n = S // s
        x = torch.randn((bs, 3, nc + 5, n * n), generator=g)
        x[:, :, 4] = -5.0 + 0.1 * torch.randn((bs, 3, n * n), generator=g)          # passes 0.001, low
        x[:, :, 5:] = -1.0 + 0.5 * torch.randn((bs, 3, nc, n * n), generator=g)
        levels.append(x)
    # sampled units: unit index u = 64*k over the concatenated (level, anchor, cell/4) units of one image
    off = 0
    for x in levels:
        plane = x.shape[-1]
        for a in range(3):
            for u in range(plane // 4):
                if (off + u) % 64 == 0:
                    x[:, a, 4, 4 * u:4 * u + 4] = 6.0
                    x[:, a, 5:, 4 * u:4 * u + 4] += 4.0
            off += plane // 4
    levels = [x.reshape(bs, 3 * (nc + 5), int(x.shape[-1] ** 0.5), -1).contiguous() for x in levels]
    dl = [l.cuda() for l in levels]
    mk = lambda cut: pipeline.PostProcessor(bs, S, nc, 2, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode="det_only",
                                            conf_cut=cut, **kw)
    a, b = mk(False), mk(True)
    a.run(dl)
    # the chain by hand up to the NMS: the misled cut starves every image
    head = ops.Head(dl, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
    cut = ops.conf_cut_levels(head, kw["conf_thres"], True, None, target=b.cut_target, cut=b.cut)
    ops.filter_levels(head, kw["conf_thres"], True, None, cap=int(b.keys.shape[1]), keys=b.keys, counts=b.counts, cut=cut)
    ops.nms_levels(head, b.keys, b.counts, kw["conf_thres"], kw["iou_thres"], True, kw["max_det"], det=b.det, det_count=b.det_count, cut=cut)
    assert (b.det_count.cpu() == -2).all()
    ops.detect_nms_redo_levels(head, kw["conf_thres"], kw["iou_thres"], None, True, True, kw["max_det"], keys=b.keys, counts=b.counts,
                               det=b.det, det_count=b.det_count)
    for x, y in zip(a.detections(), b.detections()):
        assert torch.equal(x, y)
    # the whole chain in one enqueue, and replayed from a CUDA graph
    c = mk(True)
    c.run(dl)
    torch.cuda.synchronize()
    assert (c.det_count.cpu() >= 0).all()
    for x, y in zip(a.detections(), c.detections()):
        assert torch.equal(x, y)
    c.capture(dl)
    c.det_count.fill_(-7)
    c.replay()
    for x, y in zip(a.detections(), c.detections()):
        assert torch.equal(x, y)
    ho = api.HeadOutputs(dl, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES)
    for x, y in zip(api.non_max_suppression(ho, **kw), a.detections()):      # same fallback inside the API call
        assert torch.equal(x, y)


def test_graph_burst_native_serving_loop(env):
    """ops.GraphBurst / dsn_graph_launch_many: several captured steps on two streams enqueued by one native call give the
    same results as replaying them one by one."""
    api, ops, pipeline = env
    bs, S, nc, C = 3, 160, 6, 2
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=100)
    sets = [port.synth_head_outputs(bs, S, nc, C, seed=60 + k, with_target=False) for k in range(2)]
    pps = [pipeline.PostProcessor(bs, S, nc, C, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, mode="detect", **kw) for _ in range(2)]
    dev_sets = [([l.cuda() for l in lv], seg.cuda()) for lv, seg, _ in sets]
    for q, (lv, seg) in zip(pps, dev_sets):
        q.capture(lv, seg)
    want = []
    for q in pps:
        q.replay()
        torch.cuda.synchronize()
        want.append((q.det.clone(), q.det_count.clone(), q.mask.clone()))
        q.det.zero_(); q.det_count.zero_(); q.mask.zero_()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    burst = ops.GraphBurst([pps[i % 2].graph() for i in range(6)], [streams[i % 2] for i in range(6)])
    burst.launch()
    torch.cuda.synchronize()
    for q, (d, c, m) in zip(pps, want):
        assert torch.equal(q.det_count, c) and torch.equal(q.mask, m)
        for i in range(bs):
            assert torch.equal(q.det[i, :int(c[i])], d[i, :int(c[i])])
