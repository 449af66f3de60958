"""Runs each hot-path kernel a few times on BASELINE config 2 (or --config N) inputs so that ncu can
capture them:  ncu --set full -k regex:'filter_levels|nms_kernel|seg_fused' ... python profiles/prof_kernels.py"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from bench import CONFIGS  # noqa: E402
from desenet_b200 import ops  # noqa: E402
from desenet_b200.pipeline import PostProcessor  # noqa: E402
from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--bs", type=int, default=None)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
if a.bs:
    cfg["bs"] = a.bs
dev = torch.device("cuda", 0)
ev = cfg["mode"] == "eval"
levels, seg, tgt = synth_head_outputs(cfg["bs"], cfg["S"], cfg["nc"], cfg["C"], seed=0, with_target=ev)
levels = [l.to(dev) for l in levels]
seg = seg.to(dev)
tgt = tgt.to(dev) if ev else None
pp = PostProcessor(cfg["bs"], cfg["S"], cfg["nc"], cfg["C"], DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"],
                   iou_thres=cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"], mode=cfg["mode"],
                   nclass=max(cfg["C"], 2), device=dev, overlap_streams=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(a.reps):
    flush.zero_()                      # evict L2 between repetitions
    pp.run(levels, seg, tgt)
torch.cuda.synchronize()
print("ok", pp.det_count[:4].tolist(), pp.counts[:4].tolist())
