"""The ctypes stub INTEGRATION.md shows a reference maintainer (the code block headed ``core/utils/dsn_ffi.py``) is
executed as written - only the library path is made absolute - and must return what the package's own
``non_max_suppression`` returns for the same raw head outputs."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# core/utils/dsn_ffi\.py.*?)```", md, re.S)
    assert m, "INTEGRATION.md no longer holds the dsn_ffi.py stub"
    return m.group(1)


def test_stub_matches_header_arity():
    """CPU: the stub's argtypes list has one entry per parameter of dsn_detect_nms_levels in the header."""
    src = _stub_source()
    args = re.search(r"dsn_detect_nms_levels\.argtypes = \[(.*?)\]", src, re.S).group(1)
    n_stub = len([a for a in args.replace("\n", " ").split(",") if a.strip()])
    hdr = open(os.path.join(ROOT, "include", "desenet_b200.h")).read()
    decl = re.search(r"int dsn_detect_nms_levels\((.*?)\);", hdr, re.S).group(1)
    assert n_stub == len(decl.split(",")), (n_stub, len(decl.split(",")))


@pytest.mark.gpu
def test_stub_runs_and_agrees(golden):
    import desenet_b200 as dsn
    from desenet_b200 import build
    from oracle import port
    ns = {}
    exec(_stub_source().replace('"libdesenet_b200.so"', repr(build.LIB)), ns)
    g = golden("synthT_s160_nc6.npz")
    levels = [torch.from_numpy(g[f"raw{i}"]).cuda() for i in range(3)]
    anchor_grid = torch.tensor(port.DEFAULT_ANCHORS, dtype=torch.float32).view(3, 1, 3, 1, 1, 2)      # yolo.py:250
    stride = torch.tensor(port.DEFAULT_STRIDES)
    for lv in (levels, [l.half() for l in levels]):
        got = ns["nms_from_head"](lv, anchor_grid, stride, 0.25, 0.45, 300)
        want = dsn.non_max_suppression(dsn.HeadOutputs(lv, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES), 0.25, 0.45, max_det=300)
        assert len(got) == len(want) and sum(w.shape[0] for w in want) > 0
        for a, b in zip(got, want):
            assert torch.equal(a, b.float())
