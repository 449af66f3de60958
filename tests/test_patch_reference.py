"""CPU, build container only (needs /root/reference): installing the drop-in into the imported reference
replaces the hot-path callables in every module that imported them by name, swaps Detect.forward and the seg
head's final Upsample of a model instance, leaves CPU / training behaviour untouched, and uninstalls cleanly."""
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    return ref_shim.load()


def test_install_and_uninstall(ref):
    from desenet_b200 import api, patch
    import core.models.common as common
    orig_nms = ref.general.non_max_suppression
    orig_box_iou = ref.general.box_iou
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    det, seg = model.model[-1], model.model[-2]
    orig_forward = type(det).forward
    orig_up = seg.out[4]
    img = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        (z0, x0), s0 = model(img)
    h = patch.install(model=model, fused=True)
    try:
        assert ref.general.non_max_suppression is api.non_max_suppression
        assert common.non_max_suppression is api.non_max_suppression          # imported by name (common.py:22-24)
        assert ref.plots.segoutput_to_target is api.segoutput_to_target
        assert ref.metrics.batch_pix_accuracy is api.batch_pix_accuracy
        assert ref.metrics.batch_intersection_union is api.batch_intersection_union
        assert isinstance(seg.out[4], patch.DeferredUpsample)
        assert type(det).forward is not orig_forward
        # CPU inputs keep the reference behaviour bit-for-bit (the CUDA path has no CPU fallback to take over)
        with torch.no_grad():
            (z1, x1), s1 = model(img)
        assert torch.equal(z0, z1) and torch.equal(s0, s1) and all(torch.equal(a, b) for a, b in zip(x0, x1))
        with pytest.raises(RuntimeError, match="CUDA"):
            ref.general.non_max_suppression(z1)
        # detection-metric dispatchers: CPU tensors still reach the reference's own code
        assert getattr(ref.metrics.ap_per_class, "_dsn_dispatch", False) and getattr(ref.general.box_iou, "_dsn_dispatch", False)
        b = torch.tensor([[0., 0., 10., 10.], [5., 5., 15., 15.]])
        assert torch.equal(ref.general.box_iou(b, b), orig_box_iou(b, b))
    finally:
        h.uninstall()
    assert ref.general.non_max_suppression is orig_nms and common.non_max_suppression is orig_nms
    assert ref.general.box_iou is orig_box_iou
    assert type(det).forward is orig_forward and seg.out[4] is orig_up
