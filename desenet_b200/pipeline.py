"""PostProcessor: the whole post-network path for one batch with preallocated workspaces — decode +
filter, per-image NMS, fused seg upsample/argmax/mask/confusion matrix — enqueue-only (no host sync),
optionally replayed as one CUDA graph, plus a host-buffer entry (pinned H2D -> kernels -> D2H) that is
what a caller holding host tensors pays end to end."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class PostProcessor:
    def __init__(self, bs: int, size, nc: int, n_segcls: int, anchor_grid, stride, *, conf_thres=0.25,
                 iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300, mode="detect",
                 nclass: Optional[int] = None, mask_dtype=torch.float32, device=None, overlap_streams=True,
                 na: int = 3, fused_det: bool = False, scale_params: Optional[torch.Tensor] = None, round_boxes: bool = False,
                 conf_cut: Optional[bool] = None, det_out=None, det_count_out=None, gather=None,
                 nms_cta_threads: Optional[int] = None):
        """``det_out`` / ``det_count_out``: write the detections there instead of into buffers of this object (tensors, or
        ``peer.RawView`` slices of a gather root's buffer); ``gather = (PeerGather, slot)``: after the detection chain of a
        batch, signal the root that this rank's rows of that slot are complete (see ``desenet_b200.peer``)."""
        assert mode in ("detect", "eval", "det_only")
        assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.bs, self.nc, self.C, self.mode = bs, nc, n_segcls, mode
        self.H, self.W = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
        self.anchor_grid, self.stride = anchor_grid, stride
        self.kw = dict(conf_thres=float(conf_thres), iou_thres=float(iou_thres), classes=classes,
                       agnostic=bool(agnostic), multi_label=bool(multi_label) and nc > 1, max_det=int(max_det))
        self.nclass = int(nclass if nclass is not None else max(n_segcls, 2))
        self.cm_rows = ops.cm_rows_for(n_segcls, self.nclass)
        self.na = na
        self.fused_det = fused_det
        # optional fused scale_coords/clip_coords (+ round) of the detections, general.py:598-623 / detect.py:218
        self.scale_params = scale_params
        self.round_boxes = round_boxes
        st = [float(s) for s in torch.as_tensor(stride).reshape(-1).tolist()]
        self.level_shapes = [(bs, na * (nc + 5), self.H // int(s), self.W // int(s)) for s in st]
        A = na * sum(s[2] * s[3] for s in self.level_shapes)
        cap = A * (nc if self.kw["multi_label"] else 1)
        d = self.device
        with torch.cuda.device(d):
            self.keys = torch.empty((bs, cap), dtype=torch.int64, device=d)
            self.counts = torch.zeros((bs,), dtype=torch.int32, device=d)
            self.det = torch.zeros((bs, max(int(max_det), 0), 6), dtype=torch.float32, device=d) if det_out is None else det_out
            self.det_count = torch.zeros((bs,), dtype=torch.int32, device=d) if det_count_out is None else det_count_out
            assert tuple(self.det.shape) == (bs, max(int(max_det), 0), 6) and tuple(self.det_count.shape) == (bs,)
            self.gather = gather
            self.mask = torch.empty((bs, self.H, self.W), dtype=mask_dtype, device=d) if mode == "detect" else None
            self.cm = torch.zeros((min(bs, 64), self.cm_rows, n_segcls), dtype=torch.int64, device=d) if mode == "eval" else None
            self.side = torch.cuda.Stream(device=d, priority=-1) if (overlap_streams and mode != "det_only") else None
            self.ev_in = torch.cuda.Event()
            self.ev_seg = torch.cuda.Event()
        # sampled confidence cut (ops.conf_cut_levels): worth it when an image can emit far more than the max_nms
        # candidates non_max_suppression ever looks at (general.py:728-729), i.e. multi-label at a low threshold
        # candidates the greedy NMS is expected to consume before max_det survive; the cut keeps about twice as many
        self.cut_target = ops.cut_target_for(max_det)
        self.conf_cut = (self.kw["multi_label"] and cap >= 8 * self.cut_target) if conf_cut is None else bool(conf_cut)
        self.conf_cut = self.conf_cut and not fused_det
        with torch.cuda.device(d):
            self.cut = ops.ConfCut(bs, d) if self.conf_cut else None
        # NMS CTA size (ops.nms_levels).  Eval mode: the persistent eval seg kernel leaves no room for a second kernel's CTAs
        # on an SM, so the NMS runs before / after it at full width and the 512-thread CTA is the right one.  Detect mode:
        # the library's rule (256 threads between the short-lived CTAs of the mask kernel).  (A 1024-thread instantiation
        # was measured for small eval batches - NMS alone 39.8 -> 37.2 us at 64 images, step 75.0 -> 79.8 us - and dropped.)
        if nms_cta_threads is None:
            nms_cta_threads = 512 if mode == "eval" else 0
        self.nms_cta_threads = int(nms_cta_threads)
        self._last_head = None
        # cut + filter + NMS + redo (all CTAs of the redo exit at once unless an image was starved) | filter + NMS | fused
        self.kernels_per_step = (1 if fused_det else (4 if self.conf_cut else 2)) + (0 if mode == "det_only" else 1) \
            + (1 if gather is not None else 0)
        self._graphs = {}

    # ---- enqueue-only device path --------------------------------------------------------------
    def run(self, levels: Sequence[torch.Tensor], seg_low: Optional[torch.Tensor] = None,
            target: Optional[torch.Tensor] = None):
        """Enqueue one batch on the current stream.  Results land in ``self.det`` / ``self.det_count``
        and ``self.mask`` (detect mode) or accumulate into ``self.cm`` (eval mode)."""
        head = ops.Head(levels, self.anchor_grid, self.stride)
        return self.run_head(head, seg_low, target)

    def run_head(self, head: ops.Head, seg_low=None, target=None):
        kw = self.kw
        cur = torch.cuda.current_stream(self.device)
        seg = self.mode != "det_only"
        # Critical path = filter -> NMS (both latency-bound, few bytes); the seg kernel is the HBM
        # stream.  The detection chain runs on a HIGH-PRIORITY side stream so its CTAs are placed
        # first, and the seg kernel (current stream) fills the rest of the machine around it.
        if self.side is not None:
            self.ev_in.record(cur)
            self.side.wait_event(self.ev_in)
            with torch.cuda.stream(self.side):
                self._det(head)
                self.ev_seg.record(self.side)
            if seg:
                self._seg(seg_low, target)
            cur.wait_event(self.ev_seg)
        else:
            self._det(head)
            if seg:
                self._seg(seg_low, target)
        return self.det, self.det_count, (self.mask if self.mode == "detect" else self.cm)

    def _det(self, head, allow_cut: bool = True):
        kw = self.kw
        self._last_head = head
        if self.fused_det:
            ops.detect_nms_levels(head, kw["conf_thres"], kw["iou_thres"], kw["classes"], kw["agnostic"], kw["multi_label"],
                                  kw["max_det"], keys=self.keys, counts=self.counts, det=self.det, det_count=self.det_count)
            self._scale()
            return
        cut = None
        if self.conf_cut and allow_cut:
            cut = ops.conf_cut_levels(head, kw["conf_thres"], kw["multi_label"], kw["classes"], target=self.cut_target,
                                      cut=self.cut)
        ops.filter_levels(head, kw["conf_thres"], kw["multi_label"], kw["classes"], cap=int(self.keys.shape[1]),
                          keys=self.keys, counts=self.counts, cut=cut)
        ops.nms_levels(head, self.keys, self.counts, kw["conf_thres"], kw["iou_thres"], kw["agnostic"], kw["max_det"],
                       det=self.det, det_count=self.det_count, cut=cut, cta_threads=self.nms_cta_threads)
        if cut is not None:
            # images the sampled cut starved (det_count == -2) are redone uncut on the device, inside the same enqueue /
            # graph: a negative count never reaches a consumer (HostPipeline, DetEvaluator, graph replay)
            ops.detect_nms_redo_levels(head, kw["conf_thres"], kw["iou_thres"], kw["classes"], kw["agnostic"], kw["multi_label"],
                                       kw["max_det"], keys=self.keys, counts=self.counts, det=self.det, det_count=self.det_count)
        self._scale()

    def _scale(self):
        if self.scale_params is not None and self.det.shape[1] > 0:
            ops.scale_coords_(self.det, self.scale_params, self.det_count, self.round_boxes)
        if self.gather is not None:
            self.gather[0].signal(self.gather[1])

    def _seg(self, seg_low, target):
        if self.mode == "detect":
            ops.seg_fused(seg_low, (self.H, self.W), mask_out=self.mask)
        else:
            ops.seg_fused(seg_low, (self.H, self.W), target=target, cm=self.cm, cm_rows=self.cm_rows)

    # ---- CUDA-graph replay of run() on static input buffers ---------------------------------------
    def capture(self, levels, seg_low=None, target=None, key=0):
        head = ops.Head(levels, self.anchor_grid, self.stride)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.run_head(head, seg_low, target)          # warm-up outside capture (lazy module loads)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_head(head, seg_low, target)
        self._graphs[key] = (g, head, seg_low, target)
        return g

    def replay(self, key=0):
        self._graphs[key][0].replay()

    def graph(self, key=0) -> torch.cuda.CUDAGraph:
        """The captured step graph (for ``ops.GraphBurst``: many steps enqueued by one native call)."""
        return self._graphs[key][0]

    # ---- results ---------------------------------------------------------------------------------
    def detections(self) -> List[torch.Tensor]:
        """Host-synchronising: list of (n_i, 6) tensors like the reference's non_max_suppression."""
        n = self.det_count.cpu().tolist()
        if any(v < 0 for v in n):           # -1: key buffer too small; -2 cannot occur (redone on the device in _det)
            raise RuntimeError(f"desenet_b200: detection chain reported {min(n)} (internal capacity error)")
        return [self.det[i, :n[i]].clone() for i in range(self.bs)]


class HostPipeline:
    """End-to-end entry for callers holding HOST tensors: per step, pinned host inputs are copied to
    the device, the post-process kernels run, and the results (detections + counts, and the mask or the
    confusion matrix) are copied back to pinned host memory.  ``depth`` slots on ``depth`` streams
    let the H2D of one step overlap the D2H of the previous one (PCIe is full duplex)."""

    def __init__(self, make_pp, level_shapes, seg_shape, target_shape=None, depth: int = 2, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self.slots = []
        for _ in range(depth):
            pp = make_pp()
            d = self.device
            slot = dict(
                pp=pp, stream=torch.cuda.Stream(device=d),
                levels=[torch.empty(s, dtype=torch.float32, device=d) for s in level_shapes],
                seg=torch.empty(seg_shape, dtype=torch.float32, device=d) if seg_shape else None,
                target=torch.empty(target_shape, dtype=torch.int64, device=d) if target_shape else None,
                h_det=torch.empty(tuple(pp.det.shape), dtype=torch.float32).pin_memory(),
                h_cnt=torch.empty(tuple(pp.det_count.shape), dtype=torch.int32).pin_memory(),
                h_mask=torch.empty(tuple(pp.mask.shape), dtype=pp.mask.dtype).pin_memory() if pp.mask is not None else None,
                h_cm=torch.empty(tuple(pp.cm.shape), dtype=torch.int64).pin_memory() if pp.cm is not None else None,
                done=torch.cuda.Event(),
            )
            slot["head"] = ops.Head(slot["levels"], pp.anchor_grid, pp.stride)
            self.slots.append(slot)
        s0 = self.slots[0]
        self.h2d_bytes = sum(l.numel() * 4 for l in s0["levels"]) + (s0["seg"].numel() * 4 if s0["seg"] is not None else 0) \
            + (s0["target"].numel() * 8 if s0["target"] is not None else 0)
        self.d2h_bytes = s0["h_det"].numel() * 4 + s0["h_cnt"].numel() * 4 \
            + (s0["h_mask"].numel() * s0["h_mask"].element_size() if s0["h_mask"] is not None else 0) \
            + (s0["h_cm"].numel() * 8 if s0["h_cm"] is not None else 0)
        self.i = 0

    def submit(self, h_levels, h_seg=None, h_target=None):
        """Enqueue one step (inputs are pinned host tensors).  Returns the slot holding the pinned
        host outputs; call ``slot['done'].synchronize()`` before reading them."""
        slot = self.slots[self.i % self.depth]
        self.i += 1
        pp = slot["pp"]
        with torch.cuda.stream(slot["stream"]):
            for d, h in zip(slot["levels"], h_levels):
                d.copy_(h, non_blocking=True)
            if slot["seg"] is not None:
                slot["seg"].copy_(h_seg, non_blocking=True)
            if slot["target"] is not None:
                slot["target"].copy_(h_target, non_blocking=True)
            pp.run_head(slot["head"], slot["seg"], slot["target"])
            self._copy_back(slot, pp)
        return slot

    def submit_device(self, levels, seg=None, target=None):
        """Same step for inputs that are ALREADY in device memory (what the patched model hands over): no H2D, the
        results still go back to pinned host memory."""
        slot = self.slots[self.i % self.depth]
        self.i += 1
        pp = slot["pp"]
        head = slot.get("dev_heads", {}).get(levels[0].data_ptr())
        if head is None:
            head = ops.Head(levels, pp.anchor_grid, pp.stride)
            slot.setdefault("dev_heads", {})[levels[0].data_ptr()] = head
        slot["stream"].wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(slot["stream"]):
            pp.run_head(head, seg, target)
            self._copy_back(slot, pp)
        return slot

    def _copy_back(self, slot, pp):
        slot["h_det"].copy_(pp.det, non_blocking=True)
        slot["h_cnt"].copy_(pp.det_count, non_blocking=True)
        if slot["h_mask"] is not None:
            slot["h_mask"].copy_(pp.mask, non_blocking=True)
        if slot["h_cm"] is not None:
            slot["h_cm"].copy_(pp.cm, non_blocking=True)
        slot["done"].record(slot["stream"])

    def synchronize(self):
        for s in self.slots:
            s["stream"].synchronize()
