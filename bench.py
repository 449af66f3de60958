#!/usr/bin/env python
"""bench.py — post-process images/s of the DeSeNet post-network path on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl b200|reference]

One "step" = one pass of the whole hot path (decode+filter, NMS, seg upsample+argmax+mask / confusion
matrix) over one batch of synthetic head outputs.  Prints ONE JSON line (rank 0).  See DESIGN.md §6
for what every field means and how the roofline numbers are formed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs.  bs is PER GPU at N=1; config 3 is 512 images split over N GPUs.
CONFIGS = {
    2: dict(name="cfg2 DeSeNet-s 640x640 bs64/GPU nc6 C2 conf.25 iou.45 max_det1000 detect-mode (dets+fp32 mask)",
            S=640, bs=64, nc=6, C=2, conf=0.25, iou=0.45, multi_label=False, max_det=1000, mode="detect"),
    3: dict(name="cfg3 DeSeNet-s 640x640 bs512 total nc6 C2 conf.001 iou.6 multi_label max_det300 eval-mode (dets+CM)",
            S=640, bs=512, total=True, nc=6, C=2, conf=0.001, iou=0.6, multi_label=True, max_det=300, mode="eval"),
    4: dict(name="cfg4 DeSeNet-s 1280x1280 bs32/GPU nc6 C2 conf.25 iou.45 eval-mode (seg-dominated)",
            S=1280, bs=32, nc=6, C=2, conf=0.25, iou=0.45, multi_label=False, max_det=300, mode="eval"),
    5: dict(name="cfg5 Detect nc80 640x640 bs128/GPU conf.001 iou.6 multi_label max_det300 detection-only",
            S=640, bs=128, nc=80, C=2, conf=0.001, iou=0.6, multi_label=True, max_det=300, mode="det_only"),
}
FUSED_DET_DEFAULT = False    # the fused single-launch detection kernel is opt-in until it beats filter+NMS
NBUF = 4      # rotating input sets: the working set between two uses of a buffer exceeds the 126 MB L2


def algorithmic_bytes_per_image(cfg):
    """SURVEY.md §8d / BASELINE.md §3: fp32 head logits in, int64 target in (eval) or fp32 mask out (detect)."""
    S, nc, C = cfg["S"], cfg["nc"], cfg["C"]
    A = 3 * ((S // 8) ** 2 + (S // 16) ** 2 + (S // 32) ** 2)
    det_in = 4 * A * (nc + 5)
    seg_in = 4 * C * (S // 8) ** 2
    tgt = 8 * S * S
    mask = 4 * S * S
    if cfg["mode"] == "detect":
        return dict(det=det_in, seg=seg_in + mask, total=det_in + seg_in + mask)
    if cfg["mode"] == "eval":
        return dict(det=det_in, seg=seg_in + tgt, total=det_in + seg_in + tgt)
    return dict(det=det_in, seg=0, total=det_in)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# clocks sampler: `nvidia-smi -lms` in a SEPARATE process (the recipe's clocks line) so that no NVML call
# in this process competes with the kernel / graph launches; samples cover pre-warm + the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device_index: int, period_ms: int = 100):
        import subprocess
        import tempfile
        self.proc = None
        self.path = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            fd, self.path = tempfile.mkstemp(prefix="dsn_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms), "-f", self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, skip_first: int = 0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
               "how": "nvidia-smi -lms 100 in a separate process over pre-warm + warm-up + timed region"}
        if not self.path or not os.path.exists(self.path):
            return out
        mhz, mx, reasons = [], None, set()
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    mhz.append(float(parts[1]))
                    mx = float(parts[2])
                except ValueError:
                    continue
                for name, v in zip(self.NAMES, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        mhz = mhz[skip_first:] or mhz
        if mhz:
            out.update(sm_mhz=statistics.median(mhz), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(mhz))
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port = the same torch / torchvision /
# numpy wheels the reference calls; the Python reference itself cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------
def cpu_reference_pass(cfg, levels, seg_low, target, stages=None):
    """One pass of the reference path on CPU tensors (timed by the caller); `stages` accumulates per-stage seconds."""
    from oracle import port

    def lap(name, t0):
        if stages is not None:
            stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0)

    t0 = time.perf_counter()
    z, _ = port.decode_levels(levels, port.DEFAULT_ANCHORS, port.DEFAULT_STRIDES, cfg["nc"])
    lap("decode", t0)
    t0 = time.perf_counter()
    for i in range(z.shape[0]):     # per image, as the oracle does (no 10 s break)
        port.non_max_suppression(z[i:i + 1], cfg["conf"], cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"])
    lap("nms", t0)
    if cfg["mode"] == "det_only":
        return
    t0 = time.perf_counter()
    full = port.upsample_x8_aten(seg_low)
    lap("upsample", t0)
    t0 = time.perf_counter()
    if cfg["mode"] == "detect":
        port.segoutput_to_target(full, (cfg["S"], cfg["S"]))
        lap("argmax_mask", t0)
    else:
        port.batch_pix_accuracy(full, target)
        port.batch_intersection_union(full, target, max(cfg["C"], 2))
        lap("seg_metrics", t0)


def cpu_sample_size(cfg):
    if cfg["conf"] < 0.01:
        return 2 if cfg["nc"] > 6 else 4       # torchvision CPU NMS is seconds per image here
    return 16 if cfg["S"] <= 640 else 4


def time_cpu(cfg, seed, repeats, n=None, stages=None):
    from desenet_b200.synth import synth_head_outputs
    n = cpu_sample_size(cfg) if n is None else n
    levels, seg_low, target = synth_head_outputs(n, cfg["S"], cfg["nc"], cfg["C"], seed=seed, with_target=cfg["mode"] == "eval")
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_reference_pass(cfg, levels, seg_low, target, stages)
        times.append(time.perf_counter() - t0)
    return n, times


_JSON_OUT = sys.stdout


def run_reference_arm(args, cfg, rank, world):
    """--impl reference: the reference's CPU path (oracle port), all host threads, rank 0 only.  Each
    step is a bounded sample of the batch, sized from a 1-image probe so that W+K steps take ~2 min."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    _, probe = time_cpu(cfg, args.seed, 2, n=1)
    budget_s = 120.0
    n = int(budget_s / (max(args.steps + args.warmup, 1) * max(min(probe), 1e-4)))
    n = max(1, min(n, cfg["bs"]))
    if args.warmup:
        time_cpu(cfg, args.seed, args.warmup, n=n)
    n, times = time_cpu(cfg, args.seed, args.steps, n=n)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "post-process images/s", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": f"{n} images per step (bounded sample of the batch)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n} images x {len(times)} steps of the same synthetic workload, oracle port "
                                   "(torch/torchvision/numpy CPU kernels the reference itself calls)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of CUDA-graph replay")
    ap.add_argument("--python-loop", action="store_true", help="replay the step graphs one by one from Python instead of one native burst")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=2, help="slots (streams) of the host pipeline that e2e measures")
    ap.add_argument("--prewarm-ms", type=float, default=600.0)
    ap.add_argument("--two-kernel-det", action="store_true", help="filter + NMS as two launches instead of the fused kernel")
    ap.add_argument("--fused-det", action="store_true", help="use the fused single-launch detection kernel")
    ap.add_argument("--time-fused", action="store_true", help="also time the fused detection kernel alone (kernels_us.det_fused)")
    ap.add_argument("--in-flight", type=int, default=3, help="independent batches in flight (one CUDA stream + workspace each)")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): everything else a library prints there (NCCL prints its version banner
    # to stdout) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(CONFIGS[args.config])
    if cfg.get("total"):
        cfg["bs"] = cfg["bs"] // max(world, 1)

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the desenet_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from desenet_b200 import ops
    from desenet_b200.pipeline import HostPipeline, PostProcessor
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs

    S, bs, nc, C, mode = cfg["S"], cfg["bs"], cfg["nc"], cfg["C"], cfg["mode"]
    eval_mode = mode == "eval"
    # ---- synthetic inputs: NBUF rotating sets (set k = set 0 rolled by k images) -------------------
    levels0, seg0, tgt0 = synth_head_outputs(bs, S, nc, C, seed=args.seed + rank, with_target=eval_mode)
    h_sets = []
    for k in range(NBUF):
        h_sets.append(dict(levels=[l.roll(k, 0).contiguous().pin_memory() for l in levels0],
                           seg=seg0.roll(k, 0).contiguous().pin_memory(),
                           target=tgt0.roll(k, 0).contiguous().pin_memory() if eval_mode else None))
    d_sets = [dict(levels=[l.to(dev) for l in s["levels"]], seg=s["seg"].to(dev),
                   target=s["target"].to(dev) if eval_mode else None) for s in h_sets]

    def make_pp():
        return PostProcessor(bs, S, nc, C, DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"], iou_thres=cfg["iou"],
                             multi_label=cfg["multi_label"], max_det=cfg["max_det"], mode=mode, nclass=max(C, 2), device=dev,
                             fused_det=(FUSED_DET_DEFAULT or args.fused_det) and not args.two_kernel_det)

    # `in_flight` independent batches are processed concurrently, each on its own stream with its own
    # workspace (a serving loop keeps several batches in flight; the latency-bound detection chain of one
    # batch then overlaps the HBM-bound seg kernel of another).  Step i runs on stream i % in_flight.
    n_fl = max(1, args.in_flight)
    pps = [make_pp() for _ in range(n_fl)]
    pp = pps[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_fl)]
    heads = [ops.Head(s["levels"], DEFAULT_ANCHORS, DEFAULT_STRIDES) for s in d_sets]
    use_graph = not args.no_graph
    if use_graph:
        for q in pps:
            for k, s in enumerate(d_sets):
                q.capture(s["levels"], s["seg"], s["target"], key=k)

    def step(i):
        k = i % NBUF
        q = pps[i % n_fl]
        with torch.cuda.stream(streams[i % n_fl]):
            if use_graph:
                q.replay(k)
            else:
                q.run_head(heads[k], d_sets[k]["seg"], d_sets[k]["target"])

    def fork():
        ev = torch.cuda.Event()
        ev.record()
        for st in streams:
            st.wait_event(ev)

    def join():
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream(dev).wait_event(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    time.sleep(0.3)      # let nvidia-smi start; its idle samples are dropped below
    # ---- untimed: clock ramp + W warm-up steps ----------------------------------------------------
    t_end = time.perf_counter() + args.prewarm_ms / 1e3
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(20):
            step(i)
            i += 1
        torch.cuda.synchronize(dev)
    for w in range(max(args.warmup, 3)):
        step(w)
    join()
    barrier()
    # ---- timed: exactly K steps --------------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if eval_mode:
        warm = sum(q.cm.sum(0) for q in pps)     # load the reduction (and NCCL) kernels outside the timed region
        if world > 1:
            dist.all_reduce(warm)
            wd = pp.det.new_empty((world,) + tuple(pp.det.shape))
            dist.all_gather_into_tensor(wd.view(-1, *pp.det.shape[1:]), pp.det)
            wc = pp.det_count.new_empty((world,) + tuple(pp.det_count.shape))
            dist.all_gather_into_tensor(wc.view(-1), pp.det_count)
        for q in pps:
            q.cm.zero_()
        torch.cuda.synchronize(dev)
    # the K timed steps are enqueued by ONE native call (ops.GraphBurst -> dsn_graph_launch_many): step i = graph of input
    # set i % NBUF of batch slot i % in_flight, on that slot's stream - the same sequence step(i) launches from Python
    burst = None
    if use_graph and not args.python_loop:
        burst = ops.GraphBurst([pps[i % n_fl].graph(i % NBUF) for i in range(args.steps)],
                               [streams[i % n_fl] for i in range(args.steps)])
    t_wall0 = time.perf_counter()
    ev0.record()
    fork()
    if burst is not None:
        burst.launch()
    else:
        for i in range(args.steps):
            step(i)
    join()
    cm_total = None
    if eval_mode:
        cm_total = sum(q.cm.sum(0) for q in pps)
        if world > 1:
            dist.all_reduce(cm_total)                     # int64 confusion matrix, NCCL over NVLink
            # the other collective of the path: fixed-capacity detection buffers of the last batch to every rank
            all_det = pp.det.new_empty((world,) + tuple(pp.det.shape))
            all_cnt = pp.det_count.new_empty((world,) + tuple(pp.det_count.shape))
            dist.all_gather_into_tensor(all_det.view(-1, *pp.det.shape[1:]), pp.det)
            dist.all_gather_into_tensor(all_cnt.view(-1), pp.det_count)
    ev1.record()
    torch.cuda.synchronize(dev)
    t_wall1 = time.perf_counter()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    time.sleep(0.12)     # make sure at least one sample lands after the timed region started
    sampler.stop()
    clocks = sampler.summary(skip_first=3)

    # sanity on the last batch's results (outside the timed region)
    torch.cuda.synchronize(dev)
    n_det = pp.det_count.cpu()
    assert (n_det >= 0).all(), "candidate buffer overflow"
    det_per_img = float(n_det.float().mean())
    cand_per_img = float(pp.counts.cpu().float().mean())

    # ---- per-kernel durations: each kernel alone, CUDA events on its launch stream, rotating buffers ----
    def time_alone(fn, reps):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps * 1e3     # us per launch

    kw = pp.kw
    reps = max(args.steps, 50)

    def k_cut(i):
        ops.conf_cut_levels(heads[i % NBUF], kw["conf_thres"], kw["multi_label"], kw["classes"], cut=pp.cut)

    def k_filter(i):       # the filter alone; with the confidence cut it reuses the cut of the last k_cut call (same images)
        ops.filter_levels(heads[i % NBUF], kw["conf_thres"], kw["multi_label"], kw["classes"], cap=int(pp.keys.shape[1]),
                          keys=pp.keys, counts=pp.counts, cut=pp.cut if pp.conf_cut else None)

    def k_nms(i):
        ops.nms_levels(heads[0], pp.keys, pp.counts, kw["conf_thres"], kw["iou_thres"], kw["agnostic"], kw["max_det"],
                       det=pp.det, det_count=pp.det_count, cut=pp.cut if pp.conf_cut else None)

    def k_seg(i):
        pp._seg(d_sets[i % NBUF]["seg"], d_sets[i % NBUF]["target"])

    def k_det_fused(i):
        ops.detect_nms_levels(heads[i % NBUF], kw["conf_thres"], kw["iou_thres"], kw["classes"], kw["agnostic"],
                              kw["multi_label"], kw["max_det"], keys=pp.keys, counts=pp.counts, det=pp.det, det_count=pp.det_count)

    kern_us = {}
    if pp.conf_cut:
        kern_us["conf_cut"] = time_alone(k_cut, reps)
        k_cut(0)           # every image of a set is a roll of set 0: per-image cuts differ, so time the filter on set 0 only
        heads_f = [heads[0]]
        kern_us["filter"] = time_alone(lambda i: ops.filter_levels(heads_f[0], kw["conf_thres"], kw["multi_label"], kw["classes"],
                                                                   cap=int(pp.keys.shape[1]), keys=pp.keys, counts=pp.counts,
                                                                   cut=pp.cut), reps)
    else:
        kern_us["filter"] = time_alone(k_filter, reps)
    k_filter(0)
    kern_us["nms"] = time_alone(k_nms, reps)
    if args.time_fused or pp.fused_det:      # the opt-in single-launch detection kernel, for comparison (not part of the step)
        kern_us["det_fused"] = time_alone(k_det_fused, reps)
    if mode != "det_only":
        kern_us["seg"] = time_alone(k_seg, reps)

    ab = algorithmic_bytes_per_image(cfg)
    peak, peak_src = measured_peak_gbs()
    if mode != "det_only":
        dom, dom_bytes = "seg", ab["seg"] * bs
    else:
        dom, dom_bytes = "filter", ab["det"] * bs
    achieved = dom_bytes / (kern_us[dom] * 1e-6) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"cfg{args.config}", {}).get(dom)          # bytes per image (profiles/summarize.py)
            traffic = traffic * bs if traffic is not None else None
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": {"seg": "seg_fused_rows_kernel", "filter": "filter_dense_kernel" if kw["multi_label"] else "filter_levels_kernel"}[dom],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "us_per_launch": kern_us[dom],
        "kernels_us": kern_us,
        "step": {"algorithmic_bytes": ab["total"] * bs, "achieved": ab["total"] * bs / (ms_max / args.steps * 1e-3) / 1e9,
                 "frac": ab["total"] * bs / (ms_max / args.steps * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: host buffers in, host results out, copies inside the timed region -----------------------
    def time_e2e(make):
        hp = HostPipeline(make, [tuple(l.shape) for l in levels0], tuple(seg0.shape) if mode != "det_only" else None,
                          tuple(tgt0.shape) if eval_mode else None, depth=args.e2e_depth, device=dev)
        e_steps = max(3, min(args.steps, 50))
        for i in range(3):
            s = h_sets[i % NBUF]
            hp.submit(s["levels"], s["seg"], s["target"])
        hp.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(e_steps):
            s = h_sets[i % NBUF]
            hp.submit(s["levels"], s["seg"], s["target"])
        hp.synchronize()
        t1 = time.perf_counter()
        e_t = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        return {"value": bs * world * e_steps / float(e_t.item()), "unit": "images/s", "h2d_bytes_per_step": hp.h2d_bytes,
                "d2h_bytes_per_step": hp.d2h_bytes, "steps": e_steps, "ms_per_step": 1e3 * float(e_t.item()) / e_steps}

    e2e = None
    if not args.no_e2e:
        e2e = time_e2e(make_pp)
        e2e["note"] = ("pinned host inputs -> H2D -> kernels -> D2H of detections+counts and mask/CM every step; "
                       f"{args.e2e_depth} slots on {args.e2e_depth} streams; wall clock incl. final synchronize")
        if mode == "detect":
            # the same call with the mask returned as uint8 class ids instead of the reference's float32 (an option of the
            # host API, NOT the reference contract): shows how much of e2e is the PCIe cost of the fp32 mask
            u8 = time_e2e(lambda: PostProcessor(bs, S, nc, C, DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"],
                                                iou_thres=cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"],
                                                mode=mode, nclass=max(C, 2), device=dev, mask_dtype=torch.uint8))
            e2e["uint8_mask_variant"] = {k: u8[k] for k in ("value", "d2h_bytes_per_step", "ms_per_step")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        time_cpu(cfg, args.seed, 1)
        stages = {}
        n, times = time_cpu(cfg, args.seed, 3, stages=stages)
        cpu = {"value": n / min(times), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "stage_ms_per_image": {k: round(1e3 * v / (3 * n), 3) for k, v in stages.items()},
               "sample": f"{n} images of the same synthetic workload, best of 3 after 1 warm-up; oracle port = the "
                         "torch/torchvision/numpy CPU kernels the reference calls (the Python reference cannot travel)"}

    if rank == 0:
        value = bs * world * args.steps / (ms_max * 1e-3)
        line = {
            "metric": "post-process images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak" if not cfg.get("total") else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "images_per_gpu": bs, "l2": f"rotating {NBUF} input sets (working set > 126 MB L2 between reuses)",
                       "launch": ("cuda-graph replay, K steps enqueued by one native call" if burst is not None else
                                  "cuda-graph replay from Python") if use_graph else "direct launches",
                       "det_path": "fused single launch" if pp.fused_det else
                       ("confidence cut + filter + NMS (3 launches)" if pp.conf_cut else "filter + NMS (2 launches)"),
                       "in_flight_batches": n_fl,
                       "candidates_per_image": cand_per_img, "detections_per_image": det_per_img,
                       "algorithmic_bytes_per_image": ab["total"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": pp.kernels_per_step * args.steps, "clocks": clocks,
        }
        if cm_total is not None:
            line["config"]["cm_total"] = cm_total.cpu().tolist()
            line["config"]["collectives"] = ("int64 confusion-matrix all_reduce + detection all_gather (NCCL), once at the end of "
                                             "the timed region" if world > 1 else "none (1 GPU)")
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
