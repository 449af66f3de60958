"""CPU, build container only (marker `reference`, needs /root/reference): pins the oracle port LIVE against the unmodified
reference imported in place (oracle/ref_shim.py) — what oracle/port.py's header promises.  The committed fixtures under
tests/golden/ pin the same functions on the GPU box, where the reference does not exist."""
import numpy as np
import pytest
import torch

from oracle import cport, port, ref_shim
from tests.nms_param_sets import NMS_PARAM_SETS

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    return ref_shim.load()


def _identity_detect(ref, nc):
    det = ref.Detect(nc, port.DEFAULT_ANCHORS, (8, 8, 8)).eval()
    det.stride = torch.tensor(port.DEFAULT_STRIDES)
    det.m = torch.nn.ModuleList([torch.nn.Identity()] * 3)
    return det


@pytest.mark.parametrize("nc,S,seed", [(6, 160, 21), (80, 64, 22), (1, 96, 23)])
def test_decode_bit_equal(ref, nc, S, seed):
    """core/models/yolo.py:255-277 vs port.decode_levels: same ATen ops in the same order -> bit-equal."""
    levels, _, _ = port.synth_head_outputs(2, S, nc, seed=seed, with_target=False)
    with torch.no_grad():
        z, xs = _identity_detect(ref, nc)([l.clone() for l in levels])
    z2, xs2 = port.decode_levels(levels, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, nc)
    assert torch.equal(z, z2) and all(torch.equal(a, b) for a, b in zip(xs, xs2))


@pytest.mark.parametrize("k", range(len(NMS_PARAM_SETS)))
def test_nms_bit_equal_all_parameter_sets(ref, k):
    """core/utils/general.py:659-750 (one image per call: no 10 s break) vs the port through torchvision, numpy and C."""
    kw = NMS_PARAM_SETS[k]
    levels, _, _ = port.synth_head_outputs(2, 160, 6, seed=30 + k, with_target=False)
    z, _ = port.decode_levels(levels, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, 6)
    for i in range(2):
        want = ref.non_max_suppression(z[i:i + 1].clone(), **kw)[0]
        for impl in ("torchvision", "numpy", "c"):
            got = port.non_max_suppression(z[i:i + 1], nms_impl=impl, **kw)[0]
            assert torch.equal(got, want), f"{impl} {kw} image {i}"


def test_nms_with_apriori_labels(ref):
    """general.py:690-696 (save_hybrid, val.py:228)."""
    levels, _, _ = port.synth_head_outputs(2, 160, 6, seed=41, with_target=False)
    z, _ = port.decode_levels(levels, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, 6)
    labels = [torch.tensor([[1.0, 60, 60, 30, 40], [3.0, 100, 90, 20, 20]]), torch.zeros(0, 5)]
    kw = dict(conf_thres=0.2, iou_thres=0.5, multi_label=True)
    want = ref.non_max_suppression(z.clone(), labels=labels, **kw)
    got = port.non_max_suppression(z, labels=labels, nms_impl="c", **kw)
    assert all(torch.equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("C,h,w", [(2, 20, 20), (3, 9, 13), (4, 80, 80)])
def test_seg_tail_bit_equal(ref, C, h, w):
    """yolo.py:183, plots.py:222-229, metrics.py:350-388 vs the port, the numpy restatement and the C restatement."""
    g = torch.Generator().manual_seed(50 + C)
    low = torch.randn(2, C, h, w, generator=g)
    tgt = torch.randint(-1, C + 2, (2, 8 * h, 8 * w), generator=g, dtype=torch.int64)
    tgt[0, 0, :4] = 255
    full = torch.nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True)(low)
    assert torch.equal(port.upsample_x8_aten(low), full)
    if 64 * h * w > 4096:                      # ATen takes a differently-rounded path for tiny outputs
        assert np.array_equal(port.upsample_bilinear_ac_np(low.numpy(), 8 * h, 8 * w), full.numpy())
        assert np.array_equal(cport.upsample_bilinear_ac(low.numpy(), 8 * h, 8 * w), full.numpy())
    for size in (None, (3 * h + 1, 11 * w - 2)):
        want = ref.segoutput_to_target(full, size)
        assert torch.equal(port.segoutput_to_target(full, size), want)
        assert np.array_equal(port.segoutput_to_target_np(full.numpy(), size), want.numpy())
    c, l = ref.batch_pix_accuracy(full, tgt)
    assert (c, l) == tuple(port.batch_pix_accuracy(full, tgt))
    for nclass in (2, 3, 4):
        i1, u1 = ref.batch_intersection_union(full, tgt, nclass)
        i2, u2 = port.batch_intersection_union(full, tgt, nclass)
        assert np.array_equal(i1, i2) and np.array_equal(u1, u2)
        cm = port.confusion_matrix(full.argmax(1).numpy(), tgt.numpy(), C, port.cm_rows(C, nclass))
        c3, l3, i3, u3 = port.seg_stats_from_cm(cm, nclass)
        assert (c3, l3) == (c, l) and np.array_equal(i3, i1) and np.array_equal(u3, u1)


def test_scale_coords_and_box_iou(ref):
    g = torch.Generator().manual_seed(60)
    boxes = torch.rand(50, 4, generator=g) * 640
    boxes[:, 2:] += boxes[:, :2]
    for img0, rp in (((480, 640, 3), None), ((1080, 1920, 3), None), ((375, 500), ((0.78, 0.78), (12.0, 70.0)))):
        want = ref.general.scale_coords((640, 640), boxes.clone(), img0, rp)
        assert torch.equal(port.scale_coords((640, 640), boxes, img0, rp), want)
    assert torch.equal(port.box_iou(boxes[:20], boxes[20:]), ref.general.box_iou(boxes[:20], boxes[20:]))
