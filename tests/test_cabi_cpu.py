"""CPU (-m "not gpu"): the C-ABI library builds/loads, exports every symbol include/desenet_b200.h
declares, rejects bad arguments before touching CUDA, and the host-side logic (class masks, closed
form of the seg metrics, a-priori label rows) matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from desenet_b200 import _lib, api, ops
from oracle import port

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    with open(os.path.join(ROOT, "include", "desenet_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsn_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/desenet_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in desenet_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.dsn_version() == 200
    assert lib.dsn_error_string(0) == b"ok"
    assert b"invalid" in lib.dsn_error_string(-1)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    assert lib.dsn_seg_fused(null, 0, 1, 2, 8, 8, 64, 64, null, null, null, 64, 64, null, 0, 4, null) == -1
    assert lib.dsn_upsample_bilinear_ac(null, 0, 1, 8, 8, 64, 64, null, null) == -1
    assert lib.dsn_filter_z(null, 0, 1, 10, 11, 0.25, 0, None, null, 10, null, null) == -1
    assert lib.dsn_nms_z(null, 0, 1, 10, 11, null, 10, null, 0.25, 0.45, 0, 30000, 300, 4096.0, null, null, null) == -1
    buf = (ctypes.c_float * 64)()
    assert lib.dsn_upsample_bilinear_ac(buf, 7, 1, 2, 2, 4, 4, buf, null) == -1            # unknown dtype code


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        api.non_max_suppression(torch.zeros(1, 10, 11))
    with pytest.raises(RuntimeError, match="CUDA"):
        api.segoutput_to_target(torch.zeros(1, 2, 8, 8))
    with pytest.raises(AssertionError):
        api.non_max_suppression(torch.zeros(1, 10, 11), conf_thres=1.5)
    with pytest.raises(AssertionError):
        api.non_max_suppression(torch.zeros(1, 10, 11), iou_thres=-0.1)


def test_class_mask():
    assert ops.class_mask(None, 6) is None
    m = ops.class_mask([0, 2, 65], 80)
    assert list(m) == [0b101, 0b10, 0, 0]
    assert list(ops.class_mask([7], 6)) == [0, 0, 0, 0]      # class outside nc never matches


@pytest.mark.parametrize("C", [2, 3, 4])
@pytest.mark.parametrize("nclass", [2, 3, 4])
def test_seg_stats_closed_form_matches_reference_formulas(C, nclass):
    g = torch.Generator().manual_seed(C * 10 + nclass)
    out = torch.randn(2, C, 12, 12, generator=g)
    tgt = torch.randint(-2, C + 3, (2, 12, 12), generator=g)
    tgt[0, 0, :3] = 255
    pred = out.argmax(1).numpy()
    rows = ops.cm_rows_for(C, nclass)
    assert rows == port.cm_rows(C, nclass)
    cm = port.confusion_matrix(pred, tgt.numpy(), C, rows)
    c, l, inter, union = api.seg_stats_from_cm(cm, nclass)
    c2, l2 = port.batch_pix_accuracy(out, tgt)
    i2, u2 = port.batch_intersection_union(out, tgt, nclass)
    assert (c, l) == (c2, l2) and np.array_equal(inter, i2) and np.array_equal(union, u2)
    # 3-D (slots, rows, C) input sums over slots
    c3, l3, i3, u3 = api.seg_stats_from_cm(np.stack([cm, cm]), nclass)
    assert (c3, l3) == (2 * c, 2 * l) and np.array_equal(i3, 2 * inter) and np.array_equal(u3, 2 * union)


def test_label_rows_appended_like_reference():
    z = torch.rand(2, 5, 8)
    labels = [torch.tensor([[1.0, 10, 20, 4, 6]]), torch.zeros(0, 5)]
    za = api._append_label_rows(z, labels)
    assert za.shape == (2, 6, 8)
    assert torch.equal(za[:, :5], z)
    assert za[0, 5].tolist() == [10, 20, 4, 6, 1, 0, 1, 0]
    assert za[1, 5].abs().sum() == 0
    # and the oracle's own handling gives the same candidates
    a = port.nms_candidates(za[0], 0.5, None, False)
    b = port.nms_candidates(z[0], 0.5, None, False, labels[0])
    assert torch.equal(a, b)
