"""[needs a build with a 1024-thread NMS instantiation (DSN_NMS_THREADS=1024 in a third translation unit, measured and
dropped: bs 64: NMS alone 39.8 -> 37.2 us, step 75.0 -> 79.8 us; bs 128: 41.3 -> 37.2 us, step 142.1 -> 150.3 us)]
Eval step (cfg 3 settings) at a small per-GPU batch - the shard of the 512-image batch on 8 GPUs - with the NMS in
512- and in 1024-thread CTAs: us per step with 3 batches in flight (CUDA-graph replay), and the NMS kernel alone.
    python profiles/experiments/x18_nms1024.py [bs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from bench import CONFIGS  # noqa: E402
from desenet_b200 import ops  # noqa: E402
from desenet_b200.pipeline import PostProcessor  # noqa: E402
from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(CONFIGS[3])
dev = torch.device("cuda", 0)
sets = []
for k in range(4):
    levels, seg, tgt = synth_head_outputs(bs, cfg["S"], cfg["nc"], cfg["C"], seed=k, with_target=True)
    sets.append(([l.to(dev) for l in levels], seg.to(dev), tgt.to(dev)))
res = {}
for cta in (512, 1024):
    pps = [PostProcessor(bs, cfg["S"], cfg["nc"], cfg["C"], DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"], iou_thres=cfg["iou"],
                         multi_label=True, max_det=cfg["max_det"], mode="eval", nclass=2, device=dev, nms_cta_threads=cta) for _ in range(3)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    for q in pps:
        for k, (lv, sg, tg) in enumerate(sets):
            q.capture(lv, sg, tg, key=k)
    torch.cuda.synchronize()

    def run(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % 3]):
                pps[i % 3].replay(i % 4)
    run(30)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    run(300)
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    step = a.elapsed_time(b) / 300 * 1e3
    q = pps[0]
    head = ops.Head(sets[0][0], DEFAULT_ANCHORS, DEFAULT_STRIDES)
    cut = ops.conf_cut_levels(head, cfg["conf"], True, None, target=q.cut_target, cut=q.cut)
    ops.filter_levels(head, cfg["conf"], True, None, cap=int(q.keys.shape[1]), keys=q.keys, counts=q.counts, cut=cut)
    torch.cuda.synchronize()
    a.record()
    for _ in range(100):
        ops.nms_levels(head, q.keys, q.counts, cfg["conf"], cfg["iou"], False, cfg["max_det"], det=q.det, det_count=q.det_count, cut=cut, cta_threads=cta)
    b.record()
    torch.cuda.synchronize()
    res[cta] = (step, a.elapsed_time(b) / 100 * 1e3, q.det.clone(), q.det_count.clone())
    print(f"bs {bs} eval, NMS in {cta}-thread CTAs: {step:7.1f} us/step (3 in flight), NMS alone {res[cta][1]:6.1f} us", flush=True)
print("identical detections:", torch.equal(res[512][2], res[1024][2]) and torch.equal(res[512][3], res[1024][3]))
