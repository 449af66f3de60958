#!/usr/bin/env python
"""bench.py — post-process images/s of the DeSeNet post-network path on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5] [--impl b200|reference]

One "step" = one pass of the whole hot path (decode+filter, NMS, seg upsample+argmax+mask / confusion
matrix) over one batch of synthetic head outputs.  Prints ONE JSON line (rank 0).  See DESIGN.md §6
for what every field means and how the roofline numbers are formed.

Multi-GPU (N > 1): ONE global batch, generated per image from one seed, is sharded contiguously over the ranks
(`distributed.shard_range`); every step's detections are gathered on rank 0 INSIDE the timed region (the NMS kernels
write straight into rank 0's buffer over NVLink, `desenet_b200.peer`), and in eval mode the int64 confusion matrix is
all-reduced (NCCL) before the closing event.  After the timed region the gathered result of the whole global batch is
hashed (`check.det_sha256`, `check.cm`), compared with the NCCL all-gather of the same detections, with the oracle on a
sample, and with the digest recorded from the 1-GPU run (profiles/bench/expected_checks.json) — a sharded run must
return exactly what the unsharded run returns.  The default run (config 2, weak scaling, 64 images per GPU) is followed
by a `secondary` block: config 3, the 512-image eval batch split over the N ranks (strong scaling).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs.  bs is PER GPU; "total" configs split bs over the N ranks (strong scaling).
CONFIGS = {
    1: dict(name="cfg1 DeSeNet-s 640x640 bs1 nc6 C2 conf.001 iou.6 multi_label max_det300 eval-mode (val.py post-process, the reference's CPU-runnable case)",
            S=640, bs=1, nc=6, C=2, conf=0.001, iou=0.6, multi_label=True, max_det=300, mode="eval"),
    2: dict(name="cfg2 DeSeNet-s 640x640 bs64/GPU nc6 C2 conf.25 iou.45 max_det1000 detect-mode (dets+fp32 mask)",
            S=640, bs=64, nc=6, C=2, conf=0.25, iou=0.45, multi_label=False, max_det=1000, mode="detect"),
    3: dict(name="cfg3 DeSeNet-s 640x640 bs512 total nc6 C2 conf.001 iou.6 multi_label max_det300 eval-mode (dets+CM)",
            S=640, bs=512, total=True, nc=6, C=2, conf=0.001, iou=0.6, multi_label=True, max_det=300, mode="eval"),
    4: dict(name="cfg4 DeSeNet-s 1280x1280 bs32/GPU nc6 C2 conf.25 iou.45 eval-mode (seg-dominated)",
            S=1280, bs=32, nc=6, C=2, conf=0.25, iou=0.45, multi_label=False, max_det=300, mode="eval"),
    5: dict(name="cfg5 Detect nc80 640x640 bs128/GPU conf.001 iou.6 multi_label max_det300 detection-only",
            S=640, bs=128, nc=80, C=2, conf=0.001, iou=0.6, multi_label=True, max_det=300, mode="det_only"),
}
FUSED_DET_DEFAULT = False    # the single-launch detection kernel is the on-device redo path, not the default chain
L2_BYTES = 126 * 1024 * 1024
EXPECTED_CHECKS = os.path.join(ROOT, "profiles", "bench", "expected_checks.json")


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
        return min(cfg["bs"], 2 if cfg["nc"] > 6 else 4)       # torchvision CPU NMS is seconds per image here
    return min(cfg["bs"], 16 if cfg["S"] <= 640 else 4)


def time_cpu(cfg, seed, repeats, n=None, stages=None):
    from desenet_b200.synth import synth_images
    n = cpu_sample_size(cfg) if n is None else n
    levels, seg_low, target = synth_images(0, n, cfg["S"], cfg["nc"], cfg["C"], seed=seed, with_target=cfg["mode"] == "eval")
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
# result checks (outside every timed region)
# ---------------------------------------------------------------------------------------------
def det_digest(det: torch.Tensor, cnt: torch.Tensor) -> str:
    """sha256 over (count, valid rows) of every image in order — equal digests = bit-identical detections."""
    det, cnt = det.cpu(), cnt.cpu()
    h = hashlib.sha256()
    for i in range(det.shape[0]):
        n = int(cnt[i])
        h.update(n.to_bytes(4, "little", signed=True))
        if n > 0:
            h.update(det[i, :n].contiguous().numpy().tobytes())
    return h.hexdigest()


def oracle_sample_check(cfg, levels, seg_low, target, det, cnt, n=2):
    """Rank 0: the first ``n`` global images against the oracle — NMS on the kernel's own decoded prediction (bit-exact)
    and the confusion matrix of the fused seg kernel against the C restatement (bit-exact)."""
    import desenet_b200 as dsn
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES
    from oracle import cport, port
    n = min(n, levels[0].shape[0])
    out = {"images": n}
    ho = dsn.HeadOutputs([l[:n].contiguous() for l in levels], DEFAULT_ANCHORS, DEFAULT_STRIDES)
    want = port.non_max_suppression(ho.z.cpu(), cfg["conf"], cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"], nms_impl="c")
    det, cnt = det.cpu(), cnt.cpu()
    out["nms_equal"] = all(int(cnt[i]) == want[i].shape[0] and torch.equal(det[i, :int(cnt[i])], want[i]) for i in range(n))
    if cfg["mode"] == "eval":
        S, C = cfg["S"], cfg["C"]
        rows = dsn.ops.cm_rows_for(C, max(C, 2))
        got = dsn.seg_confusion_matrix(dsn.SegLogits(seg_low[:n].contiguous()), target[:n].contiguous(), max(C, 2)).sum(0).cpu().numpy()
        full = cport.upsample_bilinear_ac(seg_low[:n].cpu().numpy(), S, S)
        ref = port.confusion_matrix(cport.argmax_first(full), target[:n].cpu().numpy(), C, rows)
        out["cm_equal"] = bool((got == ref).all())
    return out


def unsharded_detections_digest(cfg_id, seed, global_bs, dev):
    """1 GPU: the detections of global images [0, global_bs) computed WITHOUT sharding (chunks of the config's batch
    size through one PostProcessor), hashed like ``check.det_sha256`` — what an N-GPU run of the same global batch must
    reproduce.  Used by --record-expected for the weak-scaling configs (global batch = N x the per-GPU batch)."""
    from desenet_b200.pipeline import PostProcessor
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_images
    cfg = CONFIGS[cfg_id]
    bs = cfg["bs"]
    pp = PostProcessor(bs, cfg["S"], cfg["nc"], cfg["C"], DEFAULT_ANCHORS, DEFAULT_STRIDES, conf_thres=cfg["conf"], iou_thres=cfg["iou"],
                       multi_label=cfg["multi_label"], max_det=cfg["max_det"], mode="det_only", device=dev)
    dets, cnts = [], []
    for first in range(0, global_bs, bs):
        levels, _, _ = synth_images(first, bs, cfg["S"], cfg["nc"], cfg["C"], seed=seed, with_target=False)
        pp.run([l.to(dev) for l in levels])
        torch.cuda.synchronize(dev)
        dets.append(pp.det.cpu().clone())
        cnts.append(pp.det_count.cpu().clone())
    det, cnt = torch.cat(dets), torch.cat(cnts)
    return det_digest(det, cnt), int(cnt.sum())


def load_expected():
    try:
        with open(EXPECTED_CHECKS) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------
def run_config(args, cfg_id, rank, world, local_rank, dev, *, primary, host_info=None):
    """Measure one config; returns the JSON line (rank 0) or None."""
    import torch.distributed as dist

    from desenet_b200 import distributed as dd
    from desenet_b200 import ops, peer
    from desenet_b200.pipeline import HostPipeline, PostProcessor
    from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_images

    cfg = dict(CONFIGS[cfg_id])
    strong = bool(cfg.get("total"))
    global_bs = cfg["bs"] if strong else cfg["bs"] * world
    lo, hi = dd.shard_range(global_bs, rank, world)
    bs = hi - lo
    sizes = [b - a for a, b in (dd.shard_range(global_bs, r, world) for r in range(world))]
    S, nc, C, mode = cfg["S"], cfg["nc"], cfg["C"], cfg["mode"]
    eval_mode = mode == "eval"
    steps = args.steps if primary else max(20, args.steps // 4)
    warmup = max(args.warmup, 3)
    ab = algorithmic_bytes_per_image(cfg)
    # rotating input sets: the bytes touched between two uses of a buffer must exceed the L2
    nbuf = max(4, min(64, math.ceil(2.0 * L2_BYTES / max(ab["total"] * bs, 1))))

    # ---- synthetic inputs: this rank's shard of ONE global batch (image g depends on (seed, g) only) -------------
    levels0, seg0, tgt0 = synth_images(lo, bs, S, nc, C, seed=args.seed, with_target=eval_mode)
    want_host = primary and not args.no_e2e
    h_sets, d_sets = [], []
    for k in range(nbuf):       # set k = set 0 rolled by k images inside the shard (set 0 = the global batch in order)
        hl = [l.roll(k, 0).contiguous() for l in levels0]
        hs = seg0.roll(k, 0).contiguous()
        ht = tgt0.roll(k, 0).contiguous() if eval_mode else None
        if want_host and k < 4:
            hl, hs, ht = [l.pin_memory() for l in hl], hs.pin_memory(), (ht.pin_memory() if ht is not None else None)
            h_sets.append(dict(levels=hl, seg=hs, target=ht))
        d_sets.append(dict(levels=[l.to(dev) for l in hl], seg=hs.to(dev), target=ht.to(dev) if ht is not None else None))
    del levels0, seg0, tgt0

    # ---- gather of the per-image detections on rank 0: peer memory (default) or NCCL all-gather per step ----------
    n_fl = max(1, args.in_flight)
    gather = None
    gather_kind = "none (1 GPU)"
    if world > 1:
        gather_kind = "nccl all_gather per step"
        if args.gather == "peer":
            try:
                gather = peer.PeerGather(global_bs, cfg["max_det"], n_fl, dev)
                gather_kind = "peer memory: NMS kernels write into rank 0's buffer over NVLink + release/acquire flag per batch"
            except Exception as e:  # pragma: no cover - IPC refused on this box: keep going over NCCL
                print(f"[bench] peer gather unavailable on rank {rank}: {e!r}", file=sys.stderr)
                gather = None
            flag = torch.tensor([1 if gather is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and gather is not None:
                gather = None
                gather_kind = "nccl all_gather per step (peer memory unavailable)"

    def make_pp(slot=None, **over):
        kw = dict(conf_thres=cfg["conf"], iou_thres=cfg["iou"], multi_label=cfg["multi_label"], max_det=cfg["max_det"], mode=mode,
                  nclass=max(C, 2), device=dev, fused_det=(FUSED_DET_DEFAULT or args.fused_det))
        if slot is not None and gather is not None:
            kw.update(det_out=gather.det_view(slot, lo, hi), det_count_out=gather.cnt_view(slot, lo, hi), gather=(gather, slot))
        kw.update(over)
        return PostProcessor(bs, S, nc, C, DEFAULT_ANCHORS, DEFAULT_STRIDES, **kw)

    # `in_flight` independent batches are processed concurrently, each on its own stream with its own workspace (a serving
    # loop keeps several batches in flight; the latency-bound detection chain of one batch then overlaps the HBM-bound seg
    # kernel of another).  Step i runs on stream i % in_flight.
    pps = [make_pp(slot=q) for q in range(n_fl)]
    pp = pps[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_fl)]
    heads = [ops.Head(s["levels"], DEFAULT_ANCHORS, DEFAULT_STRIDES) for s in d_sets]
    use_graph = not args.no_graph
    if use_graph:
        for q in pps:
            for k, s in enumerate(d_sets):
                q.capture(s["levels"], s["seg"], s["target"], key=k)
    nccl_det = nccl_cnt = None
    if world > 1 and gather is None:
        nccl_det = [torch.empty((world * max(sizes), cfg["max_det"], 6), dtype=torch.float32, device=dev) for _ in range(n_fl)]
        nccl_cnt = [torch.empty((world * max(sizes),), dtype=torch.int32, device=dev) for _ in range(n_fl)]

    def step(i):
        k = i % nbuf
        q = pps[i % n_fl]
        with torch.cuda.stream(streams[i % n_fl]):
            if use_graph:
                q.replay(k)
            else:
                q.run_head(heads[k], d_sets[k]["seg"], d_sets[k]["target"])
            if nccl_det is not None:        # NCCL route: fixed-capacity all-gather of this step's detections
                dist.all_gather_into_tensor(nccl_det[i % n_fl], q.det)
                dist.all_gather_into_tensor(nccl_cnt[i % n_fl], q.det_count)

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

    sampler = ClockSampler(local_rank) if primary else None
    time.sleep(0.3 if primary else 0.0)      # let nvidia-smi start; its idle samples are dropped below
    # ---- untimed: clock ramp + W warm-up steps ----------------------------------------------------
    t_end = time.perf_counter() + (args.prewarm_ms if primary else 200.0) / 1e3
    i = 0
    if nccl_det is not None:
        # NCCL route: every rank must issue the same number of collectives - a fixed count instead of a time-based ramp
        for _ in range(200 if primary else 60):
            step(i)
            i += 1
        torch.cuda.synchronize(dev)
    else:
        while time.perf_counter() < t_end:
            for _ in range(20):
                step(i)
                i += 1
            torch.cuda.synchronize(dev)
    for w in range(warmup):
        step(w)
    join()
    barrier()
    # ---- timed: exactly K steps --------------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if eval_mode:
        warm = sum(q.cm.sum(0) for q in pps)     # load the reduction (and NCCL) kernels outside the timed region
        if world > 1:
            dist.all_reduce(warm)
        for q in pps:
            q.cm.zero_()
        torch.cuda.synchronize(dev)
    # root: per slot, the flag values every rank will have reached when its timed steps of that slot are done
    wait_targets = None
    if gather is not None:
        barrier()                                # every rank's warm-up signals have landed
        if rank == 0:
            per_slot = [len(range(q, steps, n_fl)) for q in range(n_fl)]
            wait_targets = [gather.flags(q).clone() + per_slot[q] for q in range(n_fl)]
        barrier()
    # the K timed steps are enqueued by ONE native call (ops.GraphBurst -> dsn_graph_launch_many): step i = graph of input
    # set i % nbuf of batch slot i % in_flight, on that slot's stream - the same sequence step(i) launches from Python
    burst = None
    if use_graph and not args.python_loop and nccl_det is None:
        burst = ops.GraphBurst([pps[i % n_fl].graph(i % nbuf) for i in range(steps)], [streams[i % n_fl] for i in range(steps)])
    ev0.record()
    fork()
    if burst is not None:
        burst.launch()
    else:
        for i in range(steps):
            step(i)
    join()
    cm_total = None
    if eval_mode:
        cm_total = sum(q.cm.sum(0) for q in pps)
        if world > 1:
            dist.all_reduce(cm_total)                     # int64 confusion matrix, NCCL over NVLink (val.py:67-70 totals)
    if wait_targets is not None:                          # rank 0: every rank's detections of every timed step have landed
        for q in range(n_fl):
            gather.wait(q, wait_targets[q])
    ev1.record()
    torch.cuda.synchronize(dev)
    barrier()
    if gather is not None and rank == 0:
        gather.check()
    ms = ev0.elapsed_time(ev1)
    ms_max = dd.max_over_ranks(ms, device=dev)
    clocks = None
    if sampler is not None:
        time.sleep(0.12)     # make sure at least one sample lands after the timed region started
        sampler.stop()
        clocks = sampler.summary(skip_first=3)

    # ---- checks on ONE pass over the global batch in order (input set 0), outside the timed region -------------------
    for q in pps:
        if q.cm is not None:
            q.cm.zero_()
    barrier()
    base_flags = gather.flags(0).clone() if (gather is not None and rank == 0) else None
    barrier()
    with torch.cuda.stream(streams[0]):
        pp.run_head(heads[0], d_sets[0]["seg"], d_sets[0]["target"])
    streams[0].synchronize()
    check = {"global_images": global_bs, "images_per_rank": sizes}
    local_det = local_cnt = None
    if gather is not None:
        # the same batch once more into LOCAL buffers, gathered over NCCL: must equal what the peer-memory gather delivered
        chk = make_pp()
        chk.run_head(heads[0], d_sets[0]["seg"], d_sets[0]["target"])
        local_det, local_cnt = chk.det, chk.det_count
        torch.cuda.synchronize(dev)
    else:
        local_det, local_cnt = pp.det, pp.det_count
    cm_chk = None
    if eval_mode:
        cm_chk = dd.allreduce_confusion_matrix(pp.cm)
    if world > 1:
        cap = max(sizes)
        pad_det = torch.zeros((cap,) + tuple(local_det.shape[1:]), dtype=torch.float32, device=dev)
        pad_cnt = torch.zeros((cap,), dtype=torch.int32, device=dev)
        pad_det[:bs], pad_cnt[:bs] = local_det, local_cnt
        all_det = torch.empty((world * cap,) + tuple(local_det.shape[1:]), dtype=torch.float32, device=dev)
        all_cnt = torch.empty((world * cap,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_det, pad_det)
        dist.all_gather_into_tensor(all_cnt, pad_cnt)
        keep = torch.cat([torch.arange(r * cap, r * cap + sizes[r], device=dev) for r in range(world)])
        g_det, g_cnt = all_det[keep], all_cnt[keep]
    else:
        g_det, g_cnt = local_det, local_cnt
    if rank == 0:
        assert int(g_cnt.min()) >= 0, f"detection chain reported {int(g_cnt.min())}"
        check["det_sha256"] = det_digest(g_det, g_cnt)
        check["detections"] = int(g_cnt.sum())
        if gather is not None:
            gather.wait(0, base_flags + 1)
            torch.cuda.synchronize(dev)
            gather.check()
            p_det, p_cnt = gather.gathered(0)
            check["peer_gather_equals_nccl_gather"] = det_digest(p_det, p_cnt) == check["det_sha256"]
            assert check["peer_gather_equals_nccl_gather"], "peer-memory gather differs from the NCCL gather"
        if cm_chk is not None:
            check["cm"] = cm_chk.cpu().tolist()
        check["oracle_sample"] = oracle_sample_check(cfg, d_sets[0]["levels"], d_sets[0]["seg"], d_sets[0]["target"], g_det, g_cnt)
        assert check["oracle_sample"]["nms_equal"], "detections differ from the oracle on the sample"
        assert check["oracle_sample"].get("cm_equal", True), "confusion matrix differs from the oracle on the sample"
        # the unsharded (1-GPU) digest of the same global batch, recorded once: a sharded run must reproduce it
        exp = load_expected().get(f"cfg{cfg_id}_images{global_bs}_seed{args.seed}")
        if exp is not None:
            same = exp.get("det_sha256") == check["det_sha256"] and (cm_chk is None or exp.get("cm") == check["cm"])
            check["equals_recorded_1gpu_result"] = bool(same)
            assert same, "sharded result differs from the recorded 1-GPU result of the same global batch"
        else:
            check["equals_recorded_1gpu_result"] = None
    det_per_img = float(g_cnt.float().mean())
    cand_per_img = float(pp.counts.cpu().float().mean())

    # ---- per-kernel durations: each kernel alone, CUDA events on its launch stream, rotating inputs AND outputs -------
    kern_us = {}
    roofline = None
    peak, peak_src = measured_peak_gbs()
    if primary or args.kernels:
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
        reps = max(steps, 50)
        lpp = [make_pp() for _ in range(n_fl)]      # local output buffers (no peer writes / signals in the kernel-alone timings)

        def k_filter(i, cut=None):
            q = lpp[i % n_fl]
            ops.filter_levels(heads[i % nbuf], kw["conf_thres"], kw["multi_label"], kw["classes"], cap=int(q.keys.shape[1]),
                              keys=q.keys, counts=q.counts, cut=cut)

        def k_nms(i):
            q = lpp[0]
            ops.nms_levels(heads[0], q.keys, q.counts, kw["conf_thres"], kw["iou_thres"], kw["agnostic"], kw["max_det"],
                           det=lpp[i % n_fl].det, det_count=lpp[i % n_fl].det_count, cut=q.cut if q.conf_cut else None,
                           cta_threads=pp.nms_cta_threads)                    # as PostProcessor._det launches it

        if pp.conf_cut:
            kern_us["conf_cut"] = time_alone(lambda i: ops.conf_cut_levels(heads[i % nbuf], kw["conf_thres"], kw["multi_label"], kw["classes"],
                                                                           target=pp.cut_target, cut=lpp[0].cut), reps)
            # per-image cuts belong to one input set: time the filter (and the NMS) on set 0 with set 0's cut
            ops.conf_cut_levels(heads[0], kw["conf_thres"], kw["multi_label"], kw["classes"], target=pp.cut_target, cut=lpp[0].cut)
            kern_us["filter"] = time_alone(lambda i: ops.filter_levels(heads[0], kw["conf_thres"], kw["multi_label"], kw["classes"],
                                                                       cap=int(lpp[0].keys.shape[1]), keys=lpp[0].keys, counts=lpp[0].counts,
                                                                       cut=lpp[0].cut), reps)
        else:
            kern_us["filter"] = time_alone(k_filter, reps)
            ops.filter_levels(heads[0], kw["conf_thres"], kw["multi_label"], kw["classes"], cap=int(lpp[0].keys.shape[1]),
                              keys=lpp[0].keys, counts=lpp[0].counts)
        kern_us["nms"] = time_alone(k_nms, reps)
        if mode != "det_only":
            kern_us["seg"] = time_alone(lambda i: lpp[i % n_fl]._seg(d_sets[i % nbuf]["seg"], d_sets[i % nbuf]["target"]), reps)
        del lpp
        if mode != "det_only":
            dom, dom_bytes = "seg", ab["seg"] * bs
        else:
            dom, dom_bytes = "filter", ab["det"] * bs
        achieved = dom_bytes / (kern_us[dom] * 1e-6) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f).get(f"cfg{cfg_id}", {})
            if t.get(dom) is not None and t.get("images_per_launch") == bs:
                traffic, traffic_src = t[dom], t.get("source")
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "kernel": {"seg": "seg_mask_rows_kernel" if mode == "detect" else "seg_fused_rows_kernel",
                                       "filter": ("filter_dense_tma_kernel" if nc <= 8 else "filter_dense_kernel") if kw["multi_label"] else "filter_levels_kernel"}[dom],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "us_per_launch": kern_us[dom],
            "kernels_us": kern_us,
            "step": {"algorithmic_bytes": ab["total"] * bs, "achieved": ab["total"] * bs / (ms_max / steps * 1e-3) / 1e9,
                     "frac": ab["total"] * bs / (ms_max / steps * 1e-3) / 1e9 / peak},
        }

    # ---- e2e: host buffers in, host results out, copies inside the timed region -----------------------
    e2e = None
    if want_host:
        level_shapes = [tuple(l.shape) for l in h_sets[0]["levels"]]
        seg_shape = tuple(h_sets[0]["seg"].shape) if mode != "det_only" else None
        tgt_shape = tuple(h_sets[0]["target"].shape) if eval_mode else None

        def time_e2e(make, device_inputs=False):
            hp = HostPipeline(make, level_shapes, seg_shape, tgt_shape, depth=args.e2e_depth, device=dev)
            e_steps = max(3, min(steps, 50))

            def sub(i):
                if device_inputs:
                    s = d_sets[i % nbuf]
                    return hp.submit_device(s["levels"], s["seg"], s["target"])
                s = h_sets[i % len(h_sets)]
                return hp.submit(s["levels"], s["seg"], s["target"])

            for i in range(3):
                sub(i)
            hp.synchronize()
            barrier()
            t0 = time.perf_counter()
            for i in range(e_steps):
                sub(i)
            hp.synchronize()
            t1 = time.perf_counter()
            e_t = dd.max_over_ranks(t1 - t0, device=dev)
            return {"value": global_bs * e_steps / e_t, "unit": "images/s", "h2d_bytes_per_step": 0 if device_inputs else hp.h2d_bytes,
                    "d2h_bytes_per_step": hp.d2h_bytes, "steps": e_steps, "ms_per_step": 1e3 * e_t / e_steps}

        e2e = time_e2e(make_pp)
        e2e["note"] = ("pinned host inputs -> H2D -> kernels -> D2H of detections+counts and mask/CM every step; "
                       f"{args.e2e_depth} slots on {args.e2e_depth} streams; wall clock incl. final synchronize; every rank moves its own "
                       "shard over its own PCIe link (bytes are per rank)")
        e2e["host"] = host_info
        dv = time_e2e(make_pp, device_inputs=True)
        e2e["device_inputs_variant"] = dict({k: dv[k] for k in ("value", "d2h_bytes_per_step", "ms_per_step")},
                                            note="head outputs already in HBM (what the patched model hands over); results still copied to pinned host memory every step")
        if mode == "detect":
            # the same call with the mask returned as uint8 class ids instead of the reference's float32 (an option of the
            # host API, NOT the reference contract): shows how much of e2e is the PCIe cost of the fp32 mask
            u8 = time_e2e(lambda: make_pp(mask_dtype=torch.uint8))
            e2e["uint8_mask_variant"] = {k: u8[k] for k in ("value", "d2h_bytes_per_step", "ms_per_step")}

    cpu = None
    if primary and rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))      # the CPU arm uses every core (the e2e binding is undone)
        except (AttributeError, OSError):
            pass
        torch.set_num_threads(os.cpu_count() or 1)
        time_cpu(cfg, args.seed, 1)
        stages = {}
        n, times = time_cpu(cfg, args.seed, 3, stages=stages)
        cpu = {"value": n / min(times), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "stage_ms_per_image": {k: round(1e3 * v / (3 * n), 3) for k, v in stages.items()},
               "sample": f"{n} images of the same synthetic workload, best of 3 after 1 warm-up; oracle port = the "
                         "torch/torchvision/numpy CPU kernels the reference calls (the Python reference cannot travel)"}

    if gather is not None:
        gather.close()
    if rank != 0:
        return None
    value = global_bs * steps / (ms_max * 1e-3)
    collectives = "none (1 GPU)"
    if world > 1:
        collectives = f"detection gather every step inside the timed region ({gather_kind})"
        if eval_mode:
            collectives += "; int64 confusion-matrix all_reduce (NCCL) once before the closing event (val.py:67-73 totals)"
    line = {
        "metric": "post-process images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "images_per_gpu": bs, "global_images": global_bs,
                   "inputs": "one global batch, image g seeded by (seed, g), sharded contiguously over the ranks",
                   "l2": f"rotating {nbuf} input sets (working set > 126 MB L2 between reuses)",
                   "launch": ("cuda-graph replay, K steps enqueued by one native call" if burst is not None else
                              "cuda-graph replay from Python") if use_graph else "direct launches",
                   "det_path": "fused single launch" if pp.fused_det else
                   ("confidence cut + filter + NMS + on-device redo of starved images (4 launches)" if pp.conf_cut else "filter + NMS (2 launches)"),
                   "in_flight_batches": n_fl, "collectives": collectives,
                   "candidates_per_image": cand_per_img, "detections_per_image": det_per_img,
                   "algorithmic_bytes_per_image": ab["total"]},
        "check": check,
        "gpu_launches": pp.kernels_per_step * steps + (n_fl if gather is not None else 0),
    }
    if roofline is not None:
        line["roofline"] = roofline
    if primary:
        line["cpu_baseline"] = cpu
        line["e2e"] = e2e
        line["clocks"] = clocks
    if cm_total is not None:
        line["config"]["cm_total"] = cm_total.cpu().tolist()
    return line


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-3 strong-scaling block that follows the default config-2 run")
    ap.add_argument("--kernels", action="store_true", help="per-kernel timings for the secondary block as well")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="N > 1: how every step's detections reach rank 0")
    ap.add_argument("--e2e-depth", type=int, default=2, help="slots (streams) of the host pipeline that e2e measures")
    ap.add_argument("--prewarm-ms", type=float, default=600.0)
    ap.add_argument("--fused-det", action="store_true", help="use the single-launch detection kernel for the whole chain")
    ap.add_argument("--in-flight", type=int, default=3, help="independent batches in flight (one CUDA stream + workspace each)")
    ap.add_argument("--record-expected", action="store_true", help="1 GPU: write this run's result digests to profiles/bench/expected_checks.json")
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
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))

    if args.impl == "reference":
        cfg = dict(CONFIGS[args.config])
        if cfg.get("total"):
            cfg["bs"] = cfg["bs"] // max(world, 1)
        run_reference_arm(args, cfg, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the desenet_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import torch.distributed as dist

    from desenet_b200 import hostbind
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host placement BEFORE any pinned allocation: the intra-op pool is spun up first (its threads keep the full mask,
    # the CPU baseline uses them), then this thread - the one that allocates pinned memory and launches - is bound to the
    # cores of the GPU's NUMA node
    torch.set_num_threads(os.cpu_count() or 1)
    torch.randn(1 << 22).sum()
    props = torch.cuda.get_device_properties(local_rank)
    bus = None
    if hasattr(props, "pci_bus_id"):
        bus = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0"
    host_info = hostbind.bind_rank(local_rank, local_world, bus)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    line = run_config(args, args.config, rank, world, local_rank, dev, primary=True, host_info=host_info)
    if args.config == 2 and not args.no_secondary:
        torch.cuda.empty_cache()
        sec = run_config(args, 3, rank, world, local_rank, dev, primary=False)
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        if args.record_expected and world == 1:
            exp = load_expected()
            for ln, cid in ((line, args.config), (line.get("secondary"), 3)):
                if ln:
                    c = ln["check"]
                    exp[f"cfg{cid}_images{c['global_images']}_seed{args.seed}"] = {
                        "det_sha256": c["det_sha256"], "cm": c.get("cm"), "detections": c["detections"],
                        "gpu": torch.cuda.get_device_name(dev), "how": "bench.py --record-expected on 1 GPU (unsharded run of the global batch)"}
            if not CONFIGS[args.config].get("total"):          # weak scaling: the global batches of the 2/4/8-GPU runs
                for n in (2, 4, 8):
                    g = CONFIGS[args.config]["bs"] * n
                    dg, nd = unsharded_detections_digest(args.config, args.seed, g, dev)
                    exp[f"cfg{args.config}_images{g}_seed{args.seed}"] = {
                        "det_sha256": dg, "cm": None, "detections": nd, "gpu": torch.cuda.get_device_name(dev),
                        "how": "bench.py --record-expected on 1 GPU (unsharded: chunks of the per-GPU batch through one PostProcessor)"}
            os.makedirs(os.path.dirname(EXPECTED_CHECKS), exist_ok=True)
            with open(EXPECTED_CHECKS, "w") as f:
                json.dump(exp, f, indent=1, sort_keys=True)
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
