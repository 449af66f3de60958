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
