"""CPU (-m "not gpu"): the oracle port / C restatement reproduce the fixtures generated from the real
reference (tests/golden/make_golden.py).  This is what pins the oracle on the GPU box, where
/root/reference does not exist."""
import numpy as np
import pytest
import torch

from oracle import cport, port
from tests.nms_param_sets import NMS_PARAM_SETS


def _levels(g):
    return [torch.from_numpy(g[f"raw{i}"]) for i in range(3)]


@pytest.mark.parametrize("name", ["synthT_s160_nc6.npz", "synthT_s64_nc80.npz"])
def test_decode_and_nms_match_reference_fixtures(golden, name):
    g = golden(name)
    nc = int(g["nc"])
    z, _ = port.decode_levels(_levels(g), port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, nc)
    assert np.array_equal(z.numpy(), g["z"])          # same wheels, same ops -> bit-exact on CPU
    for k, kw in enumerate(NMS_PARAM_SETS):
        for impl in ("torchvision", "c"):
            out = port.non_max_suppression(z, nms_impl=impl, **kw)
            for i, o in enumerate(out):
                assert np.array_equal(o.numpy(), g[f"nms{k}_{i}"]), (name, k, i, impl)


def test_real_model_fixture(golden):
    g = golden("real_desenet_s128.npz")
    z, xs = port.decode_levels(_levels(g), g["anchor_grid"], g["stride"], 6)
    assert np.array_equal(z.numpy(), g["z"])
    for i, x in enumerate(xs):
        assert np.array_equal(x.numpy(), g[f"xperm{i}"])
    out = port.non_max_suppression(z, 0.001, 0.6, multi_label=True, max_det=300)
    for i, o in enumerate(out):
        assert np.array_equal(o.numpy(), g[f"nms_val_{i}"])
        assert 0 < o.shape[0] <= 300
    out = port.non_max_suppression(z, 0.25, 0.45)
    for i, o in enumerate(out):
        assert o.shape == (0, 6) and g[f"nms_det_{i}"].shape == (0, 6)   # KAT (8): 0 dets at conf .25
    obj = g["z"][..., 4]
    assert 0.001 < obj.min() and obj.max() < 0.03
    # seg: upsample bit-exact (C restatement uses a real fmaf), mask, metrics
    full = cport.upsample_bilinear_ac(g["seg_low"], 128, 128)
    assert np.array_equal(full, g["seg_full"])
    assert np.array_equal(port.upsample_bilinear_ac_np(g["seg_low"], 128, 128), g["seg_full"])
    assert np.array_equal(port.segoutput_to_target_np(full), g["mask"])
    assert np.array_equal(port.segoutput_to_target_np(full, (96, 170)), g["mask_96x170"])
    pred = cport.argmax_first(full)
    for nclass in (2, 3):
        cm = cport.confusion_matrix(pred, g["target"], 2, port.cm_rows(2, nclass))
        c, l, inter, union = port.seg_stats_from_cm(cm, nclass)
        assert [c, l] == list(g[f"pix_{nclass}"])
        assert np.array_equal(inter, g[f"inter_{nclass}"]) and np.array_equal(union, g[f"union_{nclass}"])


def test_over_30000_candidates_fixture(golden):
    g = golden("over30k_s160_nc80.npz")
    assert int(g["n_pairs"]) > port.MAX_NMS
    z, _ = port.decode_levels(_levels(g), port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, 80)
    out = port.non_max_suppression(z, 0.001, 0.6, multi_label=True, max_det=300, nms_impl="c")[0]
    assert np.array_equal(out.numpy(), g["nms0_0"])


def test_seg_cases_fixture(golden):
    g = golden("seg_cases.npz")
    for k in range(int(g["n_cases"])):
        low, full, tgt = g[f"low{k}"], g[f"full{k}"], g[f"target{k}"]
        C = low.shape[1]
        H, W = full.shape[-2:]
        up = cport.upsample_bilinear_ac(low, H, W)
        if H * W > 4096:
            assert np.array_equal(up, full)
        else:
            # ATen switches to a differently-contracted CPU code path when out_h*out_w <= 4096
            # (probed on torch 2.11); there the restatement agrees to 1 ulp-ish, not bit-for-bit.
            assert np.abs(up - full).max() <= 5e-7
        assert np.array_equal(port.segoutput_to_target_np(full), g[f"mask{k}"])
        size = tuple(int(v) for v in g[f"resized_size{k}"])
        assert np.array_equal(port.segoutput_to_target_np(full, size), g[f"mask{k}_resized"])
        pred = cport.argmax_first(full)
        assert np.array_equal(pred, port.argmax_first_np(full))
        for nclass in (2, 3, 4):
            rows = port.cm_rows(C, nclass)
            cm = cport.confusion_matrix(pred, tgt, C, rows)
            assert np.array_equal(cm, port.confusion_matrix(pred, tgt, C, rows))
            c, l, inter, union = port.seg_stats_from_cm(cm, nclass)
            assert [c, l] == list(g[f"pix{k}"])
            assert np.array_equal(inter, g[f"inter{k}_{nclass}"]), (k, nclass)
            assert np.array_equal(union, g[f"union{k}_{nclass}"]), (k, nclass)
            # and the port's direct restatement of metrics.py
            c2, l2 = port.batch_pix_accuracy(torch.from_numpy(full), torch.from_numpy(tgt))
            i2, u2 = port.batch_intersection_union(torch.from_numpy(full), torch.from_numpy(tgt), nclass)
            assert (c2, l2) == (c, l) and np.array_equal(i2, inter) and np.array_equal(u2, union)


# ---- known-answer cases established against torchvision 0.26 / torch 2.11 CPU ----
@pytest.mark.parametrize("impl", ["torchvision", "numpy", "c"])
def test_nms_known_answers(impl):
    def run(boxes, scores, thr):
        b = np.asarray(boxes, np.float32)
        s = np.asarray(scores, np.float32)
        if impl == "torchvision":
            return port.nms_torchvision(torch.from_numpy(b), torch.from_numpy(s), thr).tolist()
        if impl == "numpy":
            return port.nms_greedy_np(b, s, thr).tolist()
        return cport.nms_greedy(b, s, thr).tolist()

    assert run([[0, 0, 10, 10]] * 2, [0.5, 0.5], 0.5) == [0]                         # (1) tie -> lower index
    assert run([[0, 0, 1, 1], [0, 0, 1, 0.5]], [0.9, 0.8], 0.5) == [0, 1]            # (2) IoU == thr kept
    assert run([[5, 5, 5, 5]] * 2, [0.9, 0.8], 0.5) == [0, 1]                        # (3) zero-area NaN
    assert run([[0, 0, 1, 2.7293658], [0, 0, 1, 1.6376196]], [0.9, 0.8], 0.6) == [0]  # (4) double compare
    assert run([[0, 0, 1, 1], [2, 2, 3, 3], [0, 0, 1, 1]], [0.1, 0.9, 0.1], 0.5) == [1, 0]


def test_misc_known_answers():
    assert np.float32(123.4567) + np.float32(79 * 4096) - np.float32(79 * 4096) == np.float32(123.46875)  # (5)
    assert port.argmax_first_np(np.zeros((1, 4, 2, 2), np.float32)).max() == 0       # (6) all-equal -> 0
    for nclass, want in ((2, [4]), (3, [2, 3]), (4, [2, 2, 1])):                     # (7) histogram rule
        got = np.histogram(np.array([0, 1, 2, 3, 1, 2]), bins=nclass - 1, range=(1, nclass))[0]
        assert got.tolist() == want


def test_seg_validation_accumulate():
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(2, 2, 16, 16, generator=g), torch.randint(0, 2, (2, 16, 16), generator=g)) for _ in range(3)]
    pa, miou = port.seg_validation_accumulate(batches, 2)
    assert 0.3 < pa < 0.7 and 0.2 < miou < 0.5


def test_conv1x1_oracle_against_exact_convolution_and_reference_fixture(golden):
    """oracle.c: orc_conv1x1_fma (the summation order the CUDA head-conv producers use) stays within 1e-6 of the exact
    float64 convolution relative to the sum of |terms|, and within 1e-5 of the reference's own conv output on the real
    model's tensors (fixture generated by tests/golden/make_golden_headconv.py from core/models/yolo.py:182)."""
    import torch
    from oracle import cport
    g = golden("real_desenet_s128_headconv.npz")
    got = cport.conv1x1_fma(g["feat"], g["weight"], g["bias"])
    exact = torch.nn.functional.conv2d(torch.from_numpy(g["feat"]).double(), torch.from_numpy(g["weight"]).double(),
                                       torch.from_numpy(g["bias"]).double()).numpy()
    scale = torch.nn.functional.conv2d(torch.from_numpy(np.abs(g["feat"])).double(), torch.from_numpy(np.abs(g["weight"])).double(),
                                       torch.from_numpy(np.abs(g["bias"])).double()).numpy()
    assert (np.abs(got - exact) <= 1e-6 * scale).all()
    assert (np.abs(got - g["seg_low"]) <= 1e-5 * scale).all()
    rng = np.random.default_rng(3)
    x, w, b = rng.standard_normal((2, 37, 5, 7), np.float32), rng.standard_normal((33, 37), np.float32), rng.standard_normal(33, np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double().view(33, 37, 1, 1), torch.from_numpy(b).double()).numpy()
    assert np.allclose(cport.conv1x1_fma(x, w, b), ref, rtol=0, atol=1e-4)
    assert np.array_equal(cport.conv1x1_fma(x, w, None), cport.conv1x1_fma(x, w, np.zeros(33, np.float32)))
