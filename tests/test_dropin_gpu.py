"""GPU: ``patch.install(model, fused=True|False)`` on CUDA, replaying the call sequences of scripts/val.py:211-234 and
scripts/detect.py:144-218 verbatim (positional NMS arguments, ``.cpu()`` on the seg output, ``half=True``) against a
stand-in model that honours the reference's attribute contract (tests/standin.py) and produces the raw head outputs of
the real random-init DeSeNet-s (fixture real_desenet_s128.npz).  Expected outputs are the reference's own results stored
in the fixture (masks: bit-exact; boxes / scores: 1e-5 relative, the device sigmoid differs from ATen's in the last ulp)
and, bit-exact, the oracle's NMS on the kernel's own decoded prediction."""
import numpy as np
import pytest
import torch

from tests import standin

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture()
def fx(golden):
    return golden("real_desenet_s128.npz")


def _close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and (a.size == 0 or (np.abs(a - b) <= REL * np.maximum(np.abs(b), 1e-30)).all())


def _decoded(fx_or_levels, fx, half=False):
    """The kernel's own decoded prediction for the fixture's raw head outputs (fp32, or fp16-rounded inputs)."""
    import desenet_b200 as dsn
    levels = [torch.from_numpy(fx[f"raw{i}"]).cuda() for i in range(3)]
    if half:
        levels = [l.half().float() for l in levels]
    return dsn.HeadOutputs(levels, fx["anchor_grid"], fx["stride"].tolist()).z.cpu()


@pytest.mark.parametrize("fused", [True, False])
def test_val_sequence_on_cuda(fx, fused):
    from desenet_b200 import api, patch
    from oracle import port
    ns, names = standin.fake_reference_modules(f"dsn_fake_gpu_val_{int(fused)}")
    model = standin.Model(fx).cuda().eval()
    img = torch.zeros(2, 3, 128, 128)
    tgt = torch.from_numpy(fx["target"])
    with patch.install(model=model, fused=fused, modules=names):
        out, seg_out, train_out = standin.val_sequence(ns, model, img, tgt)
        with torch.no_grad():
            (o2, _), s2 = model(img.cuda())
        assert isinstance(o2, api.HeadOutputs) == fused and isinstance(s2, api.SegLogits) == fused
        if not fused:
            assert o2.is_cuda and tuple(o2.shape) == (2, 1008, 11) and s2.is_cuda and tuple(s2.shape) == (2, 2, 128, 128)
            assert np.array_equal(s2.cpu().numpy(), fx["seg_full"])                # upsample kernel == ATen CPU, bit for bit
        # seg metrics through the patched names, as val.py's seg_validation / eval_batch calls them
        c, l = ns.val.batch_pix_accuracy(s2, tgt.cuda())
        i3, u3 = ns.val.batch_intersection_union(s2, tgt, 3)                       # CPU target is fine too
    want = port.non_max_suppression(_decoded(None, fx), 0.001, 0.6, multi_label=True, nms_impl="c")
    assert len(out) == 2
    for i in range(2):
        assert out[i].is_cuda and out[i].dtype == torch.float32
        assert torch.equal(out[i].cpu(), want[i]), f"image {i}: NMS differs from the oracle on the same decoded prediction"
        assert _close(out[i].cpu().numpy(), fx[f"nms_val_{i}"]), f"image {i}: NMS differs from the reference fixture"
        assert np.array_equal(train_out[i].cpu().numpy(), fx[f"xperm{i}"])        # the loss inputs (val.py:223-224)
    assert seg_out.is_cuda and np.array_equal(seg_out.cpu().numpy(), fx["mask"])
    assert [int(c), int(l)] == fx["pix_2"].tolist()
    assert np.array_equal(i3, fx["inter_3"]) and np.array_equal(u3, fx["union_3"])


@pytest.mark.parametrize("fused", [True, False])
def test_detect_sequence_on_cuda(fx, fused):
    from desenet_b200 import patch
    from oracle import port
    ns, names = standin.fake_reference_modules(f"dsn_fake_gpu_det_{int(fused)}")
    model = standin.Model(fx).cuda().eval()
    img = torch.zeros(2, 3, 128, 128)
    with patch.install(model=model, fused=fused, modules=names):
        # conf 0.25 (detect.py default): the random-init head has no detections; 0.002 exercises scale_coords + round
        pred, seg_pred = standin.detect_sequence(ns, model, img, (96, 170, 3))
        pred2, seg2 = standin.detect_sequence(ns, model, img, (96, 170, 3), conf_thres=0.002, classes=[0, 1, 2, 3], agnostic_nms=True)
    assert all(tuple(p.shape) == (0, 6) and p.is_cuda for p in pred)
    assert seg_pred.device.type == "cpu" and np.array_equal(seg_pred.numpy(), fx["mask_96x170"])    # detect.py:198
    assert np.array_equal(seg2.numpy(), fx["mask_96x170"])
    z = _decoded(None, fx)
    want = port.non_max_suppression(z, 0.002, 0.45, [0, 1, 2, 3], True, max_det=1000, nms_impl="c")
    assert sum(len(w) for w in want) > 0
    for p, w in zip(pred2, want):
        if len(w):
            w = w.clone()
            w[:, :4] = port.scale_coords((128, 128), w[:, :4], (96, 170, 3), do_round=True)
        assert torch.equal(p.cpu(), w)


@pytest.mark.parametrize("fused", [True, False])
def test_half_precision_sequences_on_cuda(fx, fused):
    """half=True (val.py:144,178-179 default): the heads hand over fp16 tensors; the kernels upcast on load, so the results
    are those of the fp32 path on the fp16-rounded logits."""
    from desenet_b200 import patch
    from oracle import cport, port
    ns, names = standin.fake_reference_modules(f"dsn_fake_gpu_half_{int(fused)}")
    model = standin.Model(fx).cuda().half().eval()
    img = torch.zeros(2, 3, 128, 128)
    tgt = torch.from_numpy(fx["target"])
    with patch.install(model=model, fused=fused, modules=names):
        out, seg_out, train_out = standin.val_sequence(ns, model, img, tgt, half=True)
        pred, seg_pred = standin.detect_sequence(ns, model, img, (96, 170, 3), half=True)
    z = _decoded(None, fx, half=True)
    if not fused:
        z = z.half().float()            # the strict path hands a materialised fp16 prediction to non_max_suppression
    want = port.non_max_suppression(z, 0.001, 0.6, multi_label=True, nms_impl="c")
    for i in range(2):
        assert torch.equal(out[i].cpu().float(), want[i])
        assert train_out[i].dtype == torch.float16
    low = torch.from_numpy(fx["seg_low"]).half().float().numpy()
    full = cport.upsample_bilinear_ac(low, 128, 128)
    if not fused:
        full = torch.from_numpy(full).half().float().numpy()      # materialised fp16 full-res logits
    assert np.array_equal(seg_out.cpu().numpy(), port.segoutput_to_target_np(full, (128, 128)))
    assert seg_pred.device.type == "cpu"
    assert np.array_equal(seg_pred.numpy(), port.segoutput_to_target_np(full, (96, 170)))


def test_function_level_install_takes_reference_tensors(fx):
    """``install()`` without a model: the reference model's own tensors (fp32 or fp16 z, full-res logits, any target dtype)
    go to the kernels; CPU tensors keep going to the reference."""
    from desenet_b200 import patch
    from oracle import port
    ns, names = standin.fake_reference_modules("dsn_fake_gpu_fn")
    z = torch.from_numpy(fx["z"])
    full = torch.from_numpy(fx["seg_full"])
    tgt = torch.from_numpy(fx["target"])
    with patch.install(modules=names):
        got = ns.val.non_max_suppression(z.cuda(), 0.001, 0.6, labels=[], multi_label=True, agnostic=False)
        got16 = ns.val.non_max_suppression(z.cuda().half(), 0.001, 0.6, labels=[], multi_label=True, agnostic=False)
        cpu = ns.val.non_max_suppression(z, 0.001, 0.6, labels=[], multi_label=True, agnostic=False)
        m = ns.detect.segoutput_to_target(full.cuda(), size=(96, 170))
        m16 = ns.detect.segoutput_to_target(full.cuda().half(), size=(96, 170))
        mc = ns.detect.segoutput_to_target(full.cuda().cpu(), size=(96, 170))
        pa = ns.val.batch_pix_accuracy(full.cuda(), tgt.to(torch.uint8))
        iu = ns.val.batch_intersection_union(full.cuda().half(), tgt.cuda().int(), 2)
    want16 = port.non_max_suppression(z.half().float(), 0.001, 0.6, multi_label=True, nms_impl="c")
    for i in range(2):
        assert np.array_equal(got[i].cpu().numpy(), fx[f"nms_val_{i}"])                 # bit-exact on the reference's z
        assert torch.equal(got16[i].cpu(), want16[i])
        assert not cpu[i].is_cuda and np.array_equal(cpu[i].numpy(), fx[f"nms_val_{i}"])
    assert m.is_cuda and np.array_equal(m.cpu().numpy(), fx["mask_96x170"])
    assert np.array_equal(m16.cpu().numpy(), port.segoutput_to_target_np(full.half().float().numpy(), (96, 170)))
    assert not mc.is_cuda and np.array_equal(mc.numpy(), fx["mask_96x170"])
    assert [int(pa[0]), int(pa[1])] == fx["pix_2"].tolist()
    i2, u2 = port.batch_intersection_union(full.half().float(), tgt, 2)
    assert np.array_equal(iu[0], i2) and np.array_equal(iu[1], u2)


def test_deferred_outputs_behave_like_tensors(fx):
    """Code that treats the deferred outputs as the reference's tensors keeps working: torch functions, indexing,
    in-place edits of the prediction (augmented inference, yolo.py:360-366), unknown attributes."""
    import torch.nn.functional as F

    import desenet_b200 as dsn
    levels = [torch.from_numpy(fx[f"raw{i}"]).cuda() for i in range(3)]
    ho = dsn.HeadOutputs(levels, fx["anchor_grid"], fx["stride"].tolist())
    zref = torch.from_numpy(fx["z"])
    assert tuple(ho.shape) == (2, 1008, 11) and ho.dtype == torch.float32 and len(ho) == 2
    assert _close(ho[..., :4].cpu().numpy(), zref[..., :4].numpy())
    cat = torch.cat((ho, ho), 1)
    assert tuple(cat.shape) == (2, 2016, 11)
    ho[..., :4] /= 2.0
    assert _close(ho.z[..., :4].cpu().numpy(), (zref[..., :4] / 2.0).numpy())
    assert _close(ho.cpu()[..., 4].numpy(), zref[..., 4].numpy())
    sl = dsn.SegLogits(torch.from_numpy(fx["seg_low"]).cuda())
    assert np.array_equal(sl.argmax(1).cpu().numpy(), torch.from_numpy(fx["seg_full"]).argmax(1).numpy())
    small = F.interpolate(sl, (64, 64), mode="bilinear", align_corners=False)          # val.py:47 on a deferred output
    want = F.interpolate(torch.from_numpy(fx["seg_full"]), (64, 64), mode="bilinear", align_corners=False)
    assert torch.allclose(small.cpu(), want, atol=1e-5)
    assert np.array_equal(sl.cpu().numpy(), fx["seg_full"])                             # .numpy() on the CPU stand-in


def test_head_convs_option_on_cuda(golden):
    """``install(model, fused=True, head_convs=True)``: the seg head's last 1x1 conv runs in the streaming kernel; the mask
    is the oracle's argmax of the upsampled logits that kernel produced, and stays within tolerance of the reference's."""
    import torch.nn as nn

    from desenet_b200 import api, ops, patch
    from oracle import cport, port
    g = golden("real_desenet_s128_headconv.npz")
    fx = golden("real_desenet_s128.npz")
    ns, names = standin.fake_reference_modules("dsn_fake_gpu_headconv")
    model = standin.Model(fx)
    conv = nn.Conv2d(128, 2, 1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g["weight"]))
        conv.bias.copy_(torch.from_numpy(g["bias"]))
    model.seg.out[0] = conv                                        # the head now really ends in Conv2d(c_hid, C, 1) + Upsample
    model.seg_low = torch.from_numpy(g["feat"]).clone()            # ... and the lookup feeds the conv's input
    model = model.cuda().eval()
    model.seg_low = model.seg_low.cuda()
    img = torch.zeros(2, 3, 128, 128)
    tgt = torch.from_numpy(fx["target"])
    with patch.install(model=model, fused=True, modules=names, head_convs=True):
        assert isinstance(model.seg.out[0], patch.FusedHeadConv)
        out, seg_out, _ = standin.val_sequence(ns, model, img, tgt)
        with torch.no_grad():
            (_, _), s2 = model(img.cuda())
    assert isinstance(model.seg.out[0], nn.Conv2d)
    assert isinstance(s2, api.SegLogits) and s2.low.dtype == torch.float32
    low = ops.seg_head_conv1x1(torch.from_numpy(g["feat"]).cuda(), conv.weight, conv.bias)
    assert torch.equal(s2.low, low)
    full = cport.upsample_bilinear_ac(low.cpu().numpy(), 128, 128)
    assert np.array_equal(seg_out.cpu().numpy(), port.segoutput_to_target_np(full, (128, 128)))
    ref_full = fx["seg_full"]
    margin = np.abs(ref_full[:, 1] - ref_full[:, 0])
    assert ((seg_out.cpu().numpy() == fx["mask"]) | (margin < 1e-4)).all()


@pytest.mark.parametrize("half", [True, False])
def test_head_convs_detect_convs_on_cuda(golden, half):
    """``install(model, fused=True, head_convs=True)`` with REAL ``Detect.m[i] = nn.Conv2d(ch, na*no, 1)`` (yolo.py:252,258):
    the patched forward produces the raw levels with dsn_detect_head_conv1x1 (tcgen05 under --half, exact fp32 FMA otherwise);
    they stay within the output type's rounding of the framework's own convolution, and the detections of the val.py call
    sequence are the oracle's for exactly those levels."""
    import torch.nn as nn

    from desenet_b200 import api, patch
    from oracle import port
    fx = golden("real_desenet_s128.npz")
    ns, names = standin.fake_reference_modules(f"dsn_fake_gpu_detconv_{int(half)}")
    model = standin.Model(fx)
    g = torch.Generator().manual_seed(11)
    chans = (128, 256, 512)
    feats = [torch.randn(2, c, 128 // s, 128 // s, generator=g) for c, s in zip(chans, (8, 16, 32))]
    for i, c in enumerate(chans):
        conv = nn.Conv2d(c, 33, 1)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(33, c, 1, 1, generator=g) * (2.0 / c ** 0.5))
            conv.bias.copy_(torch.randn(33, generator=g) - 3.0)
        model.det.m[i] = conv
        setattr(model, f"raw{i}", feats[i].clone())                       # the lookup now feeds the convs' inputs
    model = model.cuda().eval()
    if half:
        model = model.half()
    img = torch.zeros(2, 3, 128, 128)
    with torch.no_grad():
        xin = [getattr(model, f"raw{i}")[:2].clone() for i in range(3)]
        want_raw = [model.det.m[i](xin[i]) for i in range(3)]             # the framework's own convolution
        if not half:      # cuDNN's fp32 convolution runs in TF32 by default (~1e-3): compare with the exact convolution instead
            want_raw = [torch.nn.functional.conv2d(xin[i].double().cpu(), model.det.m[i].weight.double().cpu(),
                                                   model.det.m[i].bias.double().cpu()).float().cuda() for i in range(3)]
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False            # fp32 case: the exact kernel stands in for an exact fp32 convolution only
    try:
        with patch.install(model=model, fused=True, modules=names, head_convs=True):
            with torch.no_grad():
                (ho, _), _ = model(img.cuda().half() if half else img.cuda())
            assert isinstance(ho, api.HeadOutputs)
            out, _, _ = standin.val_sequence(ns, model, img, torch.from_numpy(fx["target"]), half=half)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32_was
    tol = 2e-3 if half else 2e-5
    for got, want in zip(ho.head.levels, want_raw):
        assert got.dtype == want.dtype and got.shape == want.shape
        assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol * 8), (got.float() - want.float()).abs().max()
    z = ho.z.float().cpu()
    want_det = port.non_max_suppression(z, 0.001, 0.6, multi_label=True, nms_impl="c")
    assert sum(d.shape[0] for d in out) > 0
    for d, w in zip(out, want_det):
        assert torch.equal(d.float().cpu(), w)


@pytest.mark.parametrize("fused", [True, False])
def test_autoshape_postprocess_sequence_on_cuda(fx, fused):
    """``AutoShape.forward``'s post-network part (core/models/common.py:775-778): NMS with keyword arguments, then
    ``scale_coords(shape1, y[i][:, :4], shape0[i])`` IN PLACE on a slice view, result discarded - both names are imported
    by common.py:22-24 and patched by ``install()``.  Expected: the oracle's NMS on the kernel's decoded prediction, rescaled
    by the oracle's scale_coords."""
    from desenet_b200 import patch
    from oracle import port
    ns, names = standin.fake_reference_modules(f"dsn_fake_gpu_autoshape_{int(fused)}")
    model = standin.Model(fx).cuda().eval()
    x = torch.zeros(2, 3, 128, 128).cuda()
    shape1, shape0 = [128, 128], [(96, 128), (240, 320)]
    with patch.install(model=model, fused=fused, modules=names):
        with torch.no_grad():
            y = model(x, False, False)[0][0]                                           # common.py:773 (the detection half)
        y = ns.detect.non_max_suppression(y, 0.001, iou_thres=0.45, classes=None, multi_label=False, max_det=1000)  # :775-776
        for i in range(2):
            ns.detect.scale_coords(shape1, y[i][:, :4], shape0[i])                     # :777-778, in place
    want = port.non_max_suppression(_decoded(None, fx), 0.001, 0.45, classes=None, multi_label=False, max_det=1000, nms_impl="c")
    assert sum(w.shape[0] for w in want) > 0
    for i in range(2):
        w = want[i].clone()
        w[:, :4] = port.scale_coords(shape1, w[:, :4].clone(), shape0[i])[:, :4]
        assert y[i].is_cuda and _close(y[i].cpu().numpy(), w.numpy()), f"image {i}"
