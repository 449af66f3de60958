"""TEST INFRASTRUCTURE ONLY — in-place import shim for the real reference (splwany/DeSeNet).

Imports the *unmodified* reference from ``/root/reference`` (read-only, exists only in the
build container, never on the GPU box).  Used to (a) pin ``oracle/port.py`` against the real
reference and (b) generate the golden fixtures under ``tests/golden/`` (see
``tests/golden/make_golden.py``).  Nothing in the product package imports this module.

Recipe:
  * ``RANK=1`` before import so ``core/utils/plots.py:62-64`` skips ``check_font``;
  * empty stub modules for matplotlib / seaborn / imgviz (``core/utils/metrics.py:10``,
    ``core/utils/plots.py:13-18``) — plotting only, never on the hot path;
  * ``cwd`` inside the reference tree because ``core/models/yolo.py:17`` does
    ``ROOT.relative_to(Path.cwd())``.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DESENET_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "utils", "general.py"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return lambda *a, **k: None

    m.__dict__["__getattr__"] = _getattr
    return m


_loaded = None


def load():
    """Return a namespace with the reference's hot-path callables (imported in place)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    os.environ["RANK"] = "1"
    import torchvision  # noqa: F401  (before the stubs exist: its import walks sys.modules)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "imgviz"):
        if name not in sys.modules:
            sys.modules[name] = _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from core.models import yolo as ref_yolo
        from core.utils import general as ref_general
        from core.utils import metrics as ref_metrics
        from core.utils import plots as ref_plots
    finally:
        os.chdir(cwd)
    ns = types.SimpleNamespace(
        yolo=ref_yolo, general=ref_general, metrics=ref_metrics, plots=ref_plots,
        Detect=ref_yolo.Detect, Model=ref_yolo.Model,
        non_max_suppression=ref_general.non_max_suppression,
        xywh2xyxy=ref_general.xywh2xyxy,
        segoutput_to_target=ref_plots.segoutput_to_target,
        batch_pix_accuracy=ref_metrics.batch_pix_accuracy,
        batch_intersection_union=ref_metrics.batch_intersection_union,
    )
    _loaded = ns
    return ns


@contextlib.contextmanager
def in_reference_cwd():
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def build_desenet_s(n_segcls: int = 2, nc: int | None = None):
    """Random-init DeSeNet-s with the SegMaskPSP head (BASELINE config).

    Restores the commented ``SegMaskPSP`` line of ``core/models/yolov5s_seg.yaml:48`` in an
    in-memory copy of the yaml dict (matches ``README.md:42``)."""
    import yaml
    ref = load()
    with open(os.path.join(REFERENCE_ROOT, "core", "models", "yolov5s_seg.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["se_nc"] = n_segcls
    cfg["head"][-2] = [[16, 19, 22], 1, "SegMaskPSP", ["se_nc", 1, 256, False]]
    with in_reference_cwd():
        model = ref.Model(cfg, nc=nc) if nc is not None else ref.Model(cfg)
    return model.eval()
