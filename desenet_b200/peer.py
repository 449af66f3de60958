"""Detection gather over peer memory: the per-image detections of every rank land in ONE buffer on the
gather root, written there by the NMS kernels themselves over NVLink / NVSwitch.

The reference never shards evaluation (scripts/val.py:170-172), so the split is new: images are sharded contiguously
(``distributed.shard_range``), the root allocates ``slots`` gathered buffers

    det   (global_bs, max_det, 6) fp32      cnt   (global_bs) int32      flags (world) uint32

exports them as one CUDA IPC allocation, and every rank maps it.  A rank's ``PostProcessor`` is constructed with
``det`` / ``det_count`` pointing at ITS rows of the root's buffer, so the compute step (NMS) and the collective (gather)
are one kernel: result rows cross the link as they are produced — 24 B per detection + 4 B per image, instead of an
all-gather of fixed-capacity buffers (1.5 MB per rank per batch at max_det 1000) — and ``signal()`` publishes "batch s of
rank r is complete" with a system-scope release store that the root's ``wait()`` acquires.  ``torch.distributed`` is used
once, to hand the IPC handle around.  The NCCL route (``distributed.gather_detections``) stays available for callers
that want the detections on every rank."""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


class RawView:
    """A typed view of device memory this process did not allocate through torch (a slice of the root's gathered
    buffer mapped over CUDA IPC).  Only what the kernel wrappers need: ``data_ptr()`` and ``shape``.  No torch op ever
    touches it (a torch tensor over peer-mapped memory would be attributed to the OWNING device and drag a context of
    that GPU into this process)."""

    is_cuda = True

    def __init__(self, ptr: int, shape: Tuple[int, ...], dtype: torch.dtype):
        self._ptr, self.shape, self.dtype = int(ptr), tuple(shape), dtype

    def data_ptr(self) -> int:
        return self._ptr


class _CudaArray:
    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}
        self._owner = owner


_TYPESTR = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerGather:
    """``slots`` gathered detection buffers on ``root``; see the module docstring.  Collective constructor (every rank of
    ``group`` calls it with the same arguments)."""

    def __init__(self, global_bs: int, max_det: int, slots: int, device, group=None, root: int = 0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.group = group
        ok = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if ok else 0
        self.world = dist.get_world_size(group) if ok else 1
        self.root, self.global_bs, self.max_det, self.slots = root, int(global_bs), int(max_det), int(slots)
        self.det_bytes = _align(self.global_bs * max(self.max_det, 1) * 6 * 4)
        self.cnt_bytes = _align(self.global_bs * 4)
        self.flag_bytes = _align(self.world * 4)
        self.slot_bytes = self.det_bytes + self.cnt_bytes + self.flag_bytes
        total = self.slot_bytes * self.slots
        self._owned = self._mapped = None
        err = None
        with torch.cuda.device(self.device):
            payload = [None]
            if self.rank == root:
                try:
                    p = c_void_p()
                    _lib.check(self.lib.dsn_peer_alloc(total, ctypes.byref(p)), "dsn_peer_alloc")
                    self._owned = p.value
                    h = ctypes.create_string_buffer(64)
                    _lib.check(self.lib.dsn_peer_export(c_void_p(self._owned), h), "dsn_peer_export")
                    payload = [h.raw]
                except Exception as e:          # keep taking part in the collectives below, fail together afterwards
                    err = e
            if self.world > 1:
                dist.broadcast_object_list(payload, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
            if self.rank == root:
                self.base = self._owned
            elif payload[0] is not None:
                try:
                    p = c_void_p()
                    _lib.check(self.lib.dsn_peer_open(payload[0], ctypes.byref(p)), "dsn_peer_open")
                    self._mapped = p.value
                    self.base = self._mapped
                except Exception as e:
                    err = e
            # per-slot step counters of THIS rank (device words the signal kernel increments)
            self.local_step = torch.zeros((self.slots,), dtype=torch.int32, device=self.device)
            self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)
            ok = torch.tensor([0 if (err is not None or getattr(self, "base", None) is None) else 1], dtype=torch.int32, device=self.device)
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # also: nobody signals before everybody has mapped the buffer
        if int(ok.item()) == 0:
            self._release()
            raise RuntimeError(f"desenet_b200.peer: the gathered buffer could not be shared over CUDA IPC on every rank "
                               f"(rank {self.rank}: {err!r})")

    # ---- addresses ------------------------------------------------------------------------------
    def _slot(self, q: int) -> int:
        return self.base + self.slot_bytes * q

    def det_view(self, q: int, lo: int, hi: int) -> RawView:
        return RawView(self._slot(q) + lo * self.max_det * 6 * 4, (hi - lo, self.max_det, 6), torch.float32)

    def cnt_view(self, q: int, lo: int, hi: int) -> RawView:
        return RawView(self._slot(q) + self.det_bytes + lo * 4, (hi - lo,), torch.int32)

    def flag_ptr(self, q: int, rank: int) -> int:
        return self._slot(q) + self.det_bytes + self.cnt_bytes + 4 * rank

    # ---- root-side tensors ----------------------------------------------------------------------
    def gathered(self, q: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(det (global_bs, max_det, 6), det_count (global_bs)) of slot ``q`` as torch tensors — root only."""
        assert self.rank == self.root, "the gathered buffers live on the root"
        with torch.cuda.device(self.device):
            det = torch.as_tensor(_CudaArray(self._slot(q), (self.global_bs, self.max_det, 6), "<f4", self), device=self.device)
            cnt = torch.as_tensor(_CudaArray(self._slot(q) + self.det_bytes, (self.global_bs,), "<i4", self), device=self.device)
        return det, cnt

    # ---- per-batch protocol -----------------------------------------------------------------------
    def signal(self, q: int, stream=None):
        """Enqueue on ``stream`` (default: current): everything this rank wrote for the batch of slot ``q`` is complete."""
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        rc = self.lib.dsn_peer_signal(c_void_p(self.local_step.data_ptr() + 4 * q), c_void_p(self.flag_ptr(q, self.rank)),
                                      c_void_p(st.cuda_stream))
        _lib.check(rc, "dsn_peer_signal")

    def flags(self, q: int) -> torch.Tensor:
        """Root: the (world,) int32 step flags of slot ``q`` (flags[r] = batches of that slot rank r has completed)."""
        assert self.rank == self.root
        with torch.cuda.device(self.device):
            return torch.as_tensor(_CudaArray(self.flag_ptr(q, 0), (self.world,), "<i4", self), device=self.device)

    def wait(self, q: int, targets: torch.Tensor, timeout_ms: int = 5000, stream=None):
        """Root: enqueue a wait until flags(q)[r] has reached ``targets[r]`` for every rank ((world,) int32 device tensor,
        e.g. a snapshot of ``flags(q)`` plus the batches enqueued since)."""
        assert self.rank == self.root and targets.is_cuda and targets.dtype == torch.int32 and targets.numel() == self.world
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        rc = self.lib.dsn_peer_wait(c_void_p(self.flag_ptr(q, 0)), c_void_p(targets.data_ptr()), self.world, int(timeout_ms),
                                    c_void_p(self.status.data_ptr()), c_void_p(st.cuda_stream))
        _lib.check(rc, "dsn_peer_wait")

    def check(self):
        """Host-synchronising: raise if a wait timed out."""
        s = int(self.status.item())
        if s:
            raise RuntimeError(f"desenet_b200.peer: rank {s - 1} never signalled its batch (peer gather timed out)")

    def detections(self, q: int, images_per_rank: Optional[List[int]] = None) -> List[torch.Tensor]:
        """Root, host-synchronising: list of (n_i, 6) tensors in global image order (like ``gather_detections``)."""
        det, cnt = self.gathered(q)
        n = cnt.cpu().tolist()
        return [det[i, :max(n[i], 0)] for i in range(self.global_bs)]

    def _release(self):
        with torch.cuda.device(self.device):
            if self._mapped:
                self.lib.dsn_peer_close(c_void_p(self._mapped))
                self._mapped = None
            if self._owned:
                self.lib.dsn_peer_free(c_void_p(self._owned))
                self._owned = None

    def close(self):
        """Collective: every rank unmaps, then the root frees (nobody unmaps / frees while a peer may still be writing)."""
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            if self._mapped:
                self.lib.dsn_peer_close(c_void_p(self._mapped))
                self._mapped = None
        if self.world > 1:
            dist.barrier(group=self.group)
        self._release()
