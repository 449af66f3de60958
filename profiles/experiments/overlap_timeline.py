"""Where the kernels of one step sit in time: the detection chain (side stream) and the seg kernel (main stream) of
PostProcessor.run_head, enqueued eagerly with a CUDA event between every two kernels.
    python profiles/experiments/overlap_timeline.py --config 3 [--bs N] [--order det|seg]"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from bench import CONFIGS  # noqa: E402
from desenet_b200 import ops  # noqa: E402
from desenet_b200.pipeline import PostProcessor  # noqa: E402
from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--bs", type=int, default=None)
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--order", default="det", help="which stream is enqueued first")
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
if a.bs:
    cfg["bs"] = a.bs
dev = torch.device("cuda", 0)
ev_mode = cfg["mode"] == "eval"
levels, seg, tgt = synth_head_outputs(cfg["bs"], cfg["S"], cfg["nc"], cfg["C"], seed=0, with_target=ev_mode)
levels = [l.to(dev) for l in levels]
seg = seg.to(dev)
tgt = tgt.to(dev) if ev_mode else None
pp = PostProcessor(cfg["bs"], cfg["S"], cfg["nc"], cfg["C"], DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"],
                   iou_thres=cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"], mode=cfg["mode"],
                   nclass=max(cfg["C"], 2), device=dev)
head = ops.Head(levels, DEFAULT_ANCHORS, DEFAULT_STRIDES)
kw = pp.kw
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream(dev)
side = pp.side


def E():
    return torch.cuda.Event(enable_timing=True)


def det_chain(marks):
    with torch.cuda.stream(side):
        marks.append(("det_start", E())); marks[-1][1].record(side)
        cut = None
        if pp.conf_cut:
            cut = ops.conf_cut_levels(head, kw["conf_thres"], kw["multi_label"], kw["classes"], target=pp.cut_target, cut=pp.cut)
            marks.append(("cut_end", E())); marks[-1][1].record(side)
        ops.filter_levels(head, kw["conf_thres"], kw["multi_label"], kw["classes"], cap=int(pp.keys.shape[1]), keys=pp.keys,
                          counts=pp.counts, cut=cut)
        marks.append(("filter_end", E())); marks[-1][1].record(side)
        ops.nms_levels(head, pp.keys, pp.counts, kw["conf_thres"], kw["iou_thres"], kw["agnostic"], kw["max_det"], det=pp.det,
                       det_count=pp.det_count, cut=cut)
        marks.append(("nms_end", E())); marks[-1][1].record(side)
        if cut is not None:
            ops.detect_nms_redo_levels(head, kw["conf_thres"], kw["iou_thres"], kw["classes"], kw["agnostic"], kw["multi_label"],
                                       kw["max_det"], keys=pp.keys, counts=pp.counts, det=pp.det, det_count=pp.det_count)
            marks.append(("redo_end", E())); marks[-1][1].record(side)


def seg_chain(marks):
    marks.append(("seg_start", E())); marks[-1][1].record(main)
    pp._seg(seg, tgt)
    marks.append(("seg_end", E())); marks[-1][1].record(main)


rows = []
for rep in range(a.reps):
    flush.zero_()
    torch.cuda.synchronize()
    t0 = E(); t0.record(main)
    side.wait_event(t0)
    marks = []
    if a.order == "det":
        det_chain(marks)
        if cfg["mode"] != "det_only":
            seg_chain(marks)
    else:
        if cfg["mode"] != "det_only":
            seg_chain(marks)
        det_chain(marks)
    torch.cuda.synchronize()
    rows.append({n: t0.elapsed_time(e) * 1e3 for n, e in marks})
names = list(rows[-1].keys())
print(f"cfg{a.config} bs{cfg['bs']} order={a.order}: median offsets from the step start (us), {a.reps} reps (first dropped)")
for n in names:
    print(f"  {n:12s} {statistics.median(r[n] for r in rows[1:]):9.1f}")
