"""Multi-GPU plumbing for the post-network path: one process per GPU, images sharded contiguously
across ranks, NO data-path collective per kernel.  The only exchanges are the two the north star
names: an all-reduce of the int64 confusion matrix and a gather of the per-image
detections, both through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The reference itself never shards evaluation ("Multi-GPU disabled", scripts/val.py:170-172)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of ``n_images`` over ``world`` ranks; the first ``n % world`` ranks get one more.
    Concatenating the shards in rank order restores the unsharded image order."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_confusion_matrix(cm: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the int64 confusion matrix over ranks (integer sum: order-independent, bit-exact).  Accepts the
    (slots, rows, C) accumulator of ``seg_confusion_matrix`` or an already reduced (rows, C) matrix."""
    total = cm.sum(0) if cm.dim() == 3 else cm.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def gather_detections(det: torch.Tensor, det_count: torch.Tensor, group=None,
                      images_per_rank: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
    """All-gather the fixed-capacity detection buffers ``det`` (bs_local, max_det, 6) and ``det_count``
    (bs_local) and return, on EVERY rank, the list of (n_i, 6) tensors in global image order (rank-major =
    the order of ``shard_range``).  ``images_per_rank`` is needed only when shards are uneven (the buffers are
    padded to the largest shard for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        n = det_count.cpu().tolist()
        return [det[i, :n[i]] for i in range(det.shape[0])]
    world = dist.get_world_size(group)
    if images_per_rank is None:
        images_per_rank = [int(det.shape[0])] * world
    cap = max(images_per_rank)
    if det.shape[0] < cap:      # pad the short shard
        det = torch.cat((det, det.new_zeros((cap - det.shape[0],) + tuple(det.shape[1:]))), 0)
        det_count = torch.cat((det_count, det_count.new_zeros(cap - det_count.shape[0])), 0)
    all_det = det.new_empty((world,) + tuple(det.shape))
    all_cnt = det_count.new_empty((world,) + tuple(det_count.shape))
    dist.all_gather_into_tensor(all_det.view(-1, *det.shape[1:]), det.contiguous(), group=group)
    dist.all_gather_into_tensor(all_cnt.view(-1), det_count.contiguous(), group=group)
    counts = all_cnt.cpu().tolist()
    out: List[torch.Tensor] = []
    for r in range(world):
        for i in range(images_per_rank[r]):
            out.append(all_det[r, i, :counts[r][i]])
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a scalar over ranks (used for device-timed durations: the slowest rank defines the step)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
