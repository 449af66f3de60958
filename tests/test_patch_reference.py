"""CPU, build container only (needs /root/reference): the drop-in installed into the REAL imported reference.

  * every module that imported a hot-path callable by name holds the dispatcher (scripts.val, scripts.detect,
    core.models.common, ...), ``Detect.forward`` and the seg head's final Upsample of a model instance are swapped, and
    ``uninstall()`` restores everything;
  * with the patch installed, the reference's OWN callers — ``scripts.val.run`` and ``scripts.val.seg_validation``,
    unmodified — run end to end on a CPU model and return exactly what they return without the patch (CPU inputs fall
    through to the reference's original functions; the CUDA side of the same dispatchers is exercised on the B200 by
    tests/test_dropin_gpu.py);
  * the detect.py:158-198 call sequence (``segoutput_to_target(seg_pred.cpu(), size=...)``, positional NMS arguments)
    works with the patch installed in both modes."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    return ref_shim.load()


@pytest.fixture(scope="module")
def scripts(ref):
    with ref_shim.in_reference_cwd():
        import scripts.detect as detect
        import scripts.val as val
    return val, detect


def _loader(n_batches=2, bs=2, S=64, nc=6, seed=0):
    """What create_mixed_dataloader yields (val.py:207): (img uint8, targets (n,6), seg_targets int64, paths, shapes)."""
    g = torch.Generator().manual_seed(seed)
    batches = []
    for b in range(n_batches):
        img = torch.randint(0, 256, (bs, 3, S, S), generator=g, dtype=torch.uint8)
        t = []
        for i in range(bs):
            n = 3
            xy = torch.rand(n, 2, generator=g) * 0.6 + 0.2
            wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
            cls = torch.randint(0, nc, (n, 1), generator=g).float()
            t.append(torch.cat((torch.full((n, 1), float(i)), cls, xy, wh), 1))
        seg = torch.randint(0, 2, (bs, S, S), generator=g, dtype=torch.int64)
        shapes = [((S, S), ((1.0, 1.0), (0.0, 0.0))) for _ in range(bs)]
        batches.append((img, torch.cat(t, 0), seg, [f"im{b}_{i}.jpg" for i in range(bs)], shapes))
    return batches


def test_install_and_uninstall(ref, scripts):
    from desenet_b200 import patch
    import core.models.common as common
    val, detect = scripts
    orig_nms = ref.general.non_max_suppression
    orig_box_iou = ref.general.box_iou
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    det, seg = model.model[-1], model.model[-2]
    orig_forward = type(det).forward
    orig_up = seg.out[4]
    h = patch.install(model=model, fused=True)
    try:
        for mod in (ref.general, common, val, detect):                # imported by name (common.py:22-24, val.py:30-35, detect.py:27-31)
            assert getattr(mod.non_max_suppression, "_dsn_dispatch", False), mod.__name__
            assert mod.non_max_suppression._dsn_original is orig_nms
        for mod in (ref.plots, val, detect):
            assert getattr(mod.segoutput_to_target, "_dsn_dispatch", False), mod.__name__
        assert getattr(ref.metrics.batch_pix_accuracy, "_dsn_dispatch", False) and getattr(val.batch_intersection_union, "_dsn_dispatch", False)
        assert getattr(val.seg_validation, "_dsn_dispatch", False) and getattr(val.process_batch, "_dsn_dispatch", False)
        assert getattr(ref.metrics.ap_per_class, "_dsn_dispatch", False) and getattr(ref.general.box_iou, "_dsn_dispatch", False)
        assert isinstance(seg.out[4], patch.DeferredUpsample)
        assert type(det).forward is not orig_forward
        b = torch.tensor([[0., 0., 10., 10.], [5., 5., 15., 15.]])
        assert torch.equal(ref.general.box_iou(b, b), orig_box_iou(b, b))
    finally:
        h.uninstall()
    assert ref.general.non_max_suppression is orig_nms and common.non_max_suppression is orig_nms
    assert val.non_max_suppression is orig_nms and detect.non_max_suppression is orig_nms
    assert ref.general.box_iou is orig_box_iou
    assert type(det).forward is orig_forward and seg.out[4] is orig_up


@pytest.mark.parametrize("fused", [True, False])
def test_reference_val_run_unchanged_on_cpu(ref, scripts, fused):
    """scripts/val.py:125-345 ``run`` and :42-76 ``seg_validation``, unmodified, with the patch installed."""
    from desenet_b200 import patch
    val, _ = scripts
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    model.de_names = [str(i) for i in range(6)]
    model.se_names = ["a", "b"]
    data = {"de": {"nc": 6}, "se": {"nc": 2}}
    kw = dict(model=model, plots=False, half=False, conf_thres=0.001, iou_thres=0.6)

    def both():
        r = val.run(data, dataloader=_loader(), **kw)
        miou = val.seg_validation(model, 2, _loader(), half_precision=False)
        return r[0], r[1], miou

    base = both()
    with patch.install(model=model, fused=fused):
        got = both()
    assert np.array_equal(np.asarray(base[0], np.float64), np.asarray(got[0], np.float64), equal_nan=True)
    assert np.array_equal(base[1], got[1]) and base[2] == got[2]


@pytest.mark.parametrize("fused", [True, False])
def test_detect_call_sequence_on_cpu(ref, scripts, fused):
    """scripts/detect.py:158,190,198,218 with the names as detect.py resolves them (its own module globals)."""
    from desenet_b200 import patch
    _, detect = scripts
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    img = torch.rand(1, 3, 64, 96)
    im0_shape = (270, 480, 3)

    def seq():
        with torch.no_grad():
            (pred, _), seg_pred = model(img, augment=False, visualize=False)                        # detect.py:158
        pred = detect.non_max_suppression(pred, 0.001, 0.45, None, False, max_det=1000)             # detect.py:190
        seg_pred = detect.segoutput_to_target(seg_pred.cpu(), size=im0_shape[:2])                   # detect.py:198
        for det in pred:
            if len(det):
                det[:, :4] = detect.scale_coords(img.shape[2:], det[:, :4], im0_shape).round()      # detect.py:218
        return pred, seg_pred

    base = seq()
    with patch.install(model=model, fused=fused):
        got = seq()
    assert len(base[0]) == len(got[0]) and all(torch.equal(a, b) for a, b in zip(base[0], got[0]))
    assert len(base[0][0]) > 0
    assert torch.equal(base[1], got[1]) and tuple(got[1].shape) == (1, 270, 480)
