"""desenet_b200 — B200-native (sm_100a) post-network inference path of splwany/DeSeNet.

Hand-written CUDA kernels behind a C ABI (``include/desenet_b200.h``), called from a thin PyTorch
host layer that mirrors the reference's own callables.  No CPU fallback: every entry point raises
if its inputs are not on a CUDA device or the CUDA library is missing."""
from .api import (HeadOutputs, SegLogits, box_iou, process_batch, batch_intersection_union, batch_pix_accuracy, detect_forward,  # noqa: F401
                  non_max_suppression, scale_coords, clip_coords, seg_confusion_matrix, seg_stats_from_cm, seg_validation, segoutput_to_target,
                  xywh2xyxy, ap_per_class, compute_ap, ConfusionMatrix)

__version__ = "0.1.0"
