"""Seeded synthetic head outputs (family T): what a DeSeNet head emits right before
the post-network path.  CPU RNG so every rank / run / box sees the same bytes for a given seed."""
from __future__ import annotations

import torch

DEFAULT_ANCHORS = ((10, 13, 16, 30, 33, 23), (30, 61, 62, 45, 59, 119), (116, 90, 156, 198, 373, 326))
DEFAULT_STRIDES = (8.0, 16.0, 32.0)


def synth_head_outputs(bs: int, S: int, nc: int, C: int = 2, seed: int = 0, obj_mean: float = -4.0,
                       cls_mean: float = -2.0, na: int = 3, strides=(8, 16, 32), with_target: bool = True):
    """Raw Detect conv outputs per level (bs, na*no, S/s, S/s): box channels ~N(0,1), objectness
    ~N(obj_mean, 1.5^2), class channels ~N(cls_mean, 1.5^2); low-res seg logits (bs, C, S/8, S/8)
    ~N(0,1); int64 targets randint(0, C) (bs, S, S)."""
    g = torch.Generator().manual_seed(seed)
    no = nc + 5
    levels = []
    for s in strides:
        n = S // s
        t = torch.randn(bs, na, no, n, n, generator=g)
        t[:, :, 4] = t[:, :, 4] * 1.5 + obj_mean
        t[:, :, 5:] = t[:, :, 5:] * 1.5 + cls_mean
        levels.append(t.reshape(bs, na * no, n, n).contiguous())
    seg = torch.randn(bs, C, S // 8, S // 8, generator=g)
    target = torch.randint(0, C, (bs, S, S), generator=g, dtype=torch.int64) if with_target else None
    return levels, seg, target


def synth_images(first: int, count: int, S: int, nc: int, C: int = 2, seed: int = 0, obj_mean: float = -4.0,
                 cls_mean: float = -2.0, na: int = 3, strides=(8, 16, 32), with_target: bool = True):
    """The same family, but seeded PER IMAGE: global image ``g`` depends only on ``(seed, g)``, so any rank of any world
    size generates exactly the bytes of its shard of ONE global batch (``first`` = the shard's first global image) and
    the sharded result can be compared with the unsharded one image by image."""
    no = nc + 5
    levels = [torch.empty(count, na * no, S // s, S // s) for s in strides]
    seg = torch.empty(count, C, S // 8, S // 8)
    target = torch.empty(count, S, S, dtype=torch.int64) if with_target else None
    g = torch.Generator()
    for i in range(count):
        g.manual_seed(int(seed) * 1_000_003 + first + i)
        for lvl, s in zip(levels, strides):
            n = S // s
            t = torch.randn(na, no, n, n, generator=g)
            t[:, 4] = t[:, 4] * 1.5 + obj_mean
            t[:, 5:] = t[:, 5:] * 1.5 + cls_mean
            lvl[i] = t.reshape(na * no, n, n)
        seg[i] = torch.randn(C, S // 8, S // 8, generator=g)
        if with_target:
            target[i] = torch.randint(0, C, (S, S), generator=g, dtype=torch.int64)
    return levels, seg, target
