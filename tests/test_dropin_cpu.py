"""CPU: the drop-in installed over a reference-shaped module layout (tests/standin.py) leaves every CPU call path on the
reference's own functions — the call sequences of scripts/val.py:211-234 and scripts/detect.py:144-218 give the
fixture's reference results with the patch installed, in both modes — and the deferred head outputs behave like the
tensors they stand for (``.cpu()`` / ``.to()`` / ``.float()`` / ``.half()``, shape / device / dtype)."""
import numpy as np
import pytest
import torch

from tests import standin


@pytest.fixture()
def fx(golden):
    return golden("real_desenet_s128.npz")


@pytest.mark.parametrize("fused", [True, False])
def test_cpu_call_sequences_unchanged_by_install(fx, fused):
    from desenet_b200 import patch
    ns, names = standin.fake_reference_modules(f"dsn_fake_cpu_{int(fused)}")
    orig_nms, orig_seg = ns.general.non_max_suppression, ns.plots.segoutput_to_target
    model = standin.Model(fx).eval()
    img = torch.zeros(2, 3, 128, 128)
    tgt = torch.from_numpy(fx["target"])
    h = patch.install(model=model, fused=fused, modules=names)
    try:
        # every module that imported the names holds the dispatcher now (val.py:30-38, detect.py:27-32)
        for m in (ns.general, ns.val, ns.detect):
            assert getattr(m.non_max_suppression, "_dsn_dispatch", False) and m.non_max_suppression._dsn_original is orig_nms
        assert ns.detect.segoutput_to_target._dsn_original is orig_seg
        assert isinstance(model.seg.out[-1], patch.DeferredUpsample)
        out, seg_out, train_out = standin.val_sequence(ns, model, img, tgt)
        for i in range(2):
            assert np.array_equal(out[i].numpy(), fx[f"nms_val_{i}"])
            assert np.array_equal(train_out[i].numpy(), fx[f"xperm{i}"])
        assert np.array_equal(seg_out.numpy(), fx["mask"])
        pred, seg_pred = standin.detect_sequence(ns, model, img, (96, 170, 3))
        assert all(p.shape == (0, 6) for p in pred)                     # random-init objectness: nothing at conf .25
        assert np.array_equal(seg_pred.numpy(), fx["mask_96x170"])
        c, l = ns.val.batch_pix_accuracy(torch.from_numpy(fx["seg_full"]), tgt)
        assert [int(c), int(l)] == fx["pix_2"].tolist()
        i_, u_ = ns.val.batch_intersection_union(torch.from_numpy(fx["seg_full"]), tgt, 3)
        assert np.array_equal(i_, fx["inter_3"]) and np.array_equal(u_, fx["union_3"])
    finally:
        h.uninstall()
    assert ns.val.non_max_suppression is orig_nms and ns.detect.segoutput_to_target is orig_seg
    assert isinstance(model.seg.out[-1], torch.nn.Upsample)


def test_install_twice_keeps_the_true_original(fx):
    from desenet_b200 import patch
    ns, names = standin.fake_reference_modules("dsn_fake_twice")
    orig = ns.general.non_max_suppression
    h1 = patch.install(modules=names)
    h2 = patch.install(modules=names)          # nothing left to patch: attributes are dispatchers already
    assert ns.val.non_max_suppression._dsn_original is orig
    h2.uninstall()
    h1.uninstall()
    assert ns.val.non_max_suppression is orig


def test_seglogits_stands_in_for_the_full_res_tensor():
    from desenet_b200.api import SegLogits
    low = torch.randn(2, 3, 5, 7)
    sl = SegLogits(low)
    assert sl.shape == (2, 3, 40, 56) and tuple(sl.size()) == (2, 3, 40, 56) and sl.size(1) == 3 and sl.dim() == 4
    assert sl.dtype == torch.float32 and sl.device == low.device and len(sl) == 2
    moved = sl.cpu()
    assert isinstance(moved, SegLogits) and moved.low is low and moved.device.type == "cpu"
    h = sl.half()
    assert isinstance(h, SegLogits) and h.dtype == torch.float16 and h.low is low       # data stays fp32 for the kernels
    assert sl.to("cpu", torch.float64).dtype == torch.float64
    assert sl.to(torch.zeros(1, dtype=torch.bfloat16)).dtype == torch.bfloat16
    assert sl.detach() is sl and sl.float().dtype == torch.float32
    assert SegLogits(low, size=(33, 41)).shape == (2, 3, 33, 41)
