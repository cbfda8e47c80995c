"""Frame sharding across the GPUs of one box: independent frames, no collective on the hot path.

Mirrors what the reference does with ``gluon.utils.split_and_load(batch, ctx_list, batch_axis=0,
even_split=False)`` (detect_yolo3.py:211-213, train_yolov3.py:445-451) and the host-side
concatenation of ``as_numpy`` (utils/general.py:6-18), but as one process per GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def split_frames(n: int, num_slice: int) -> List[Tuple[int, int]]:
    """Slice boundaries of gluon ``split_data(..., even_split=False)``: step = n // num_slice, the last
    slice takes the remainder; with n < num_slice only n single-frame slices exist (the rest are empty)."""
    if num_slice < 1:
        raise ValueError("num_slice must be >= 1")
    if n < num_slice:
        return [(i, i + 1) if i < n else (n, n) for i in range(num_slice)]
    step = n // num_slice
    return [(i * step, (i + 1) * step if i < num_slice - 1 else n) for i in range(num_slice)]


def shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    return split_frames(n, world)[rank]


def window_halo(k: int, step: int = 1) -> int:
    """Frames of context a shard of a continuous stream needs on each side for a centred temporal
    window of K frames (datasets/imgnetvid.py:480-506): floor(K/2)*step."""
    return (k // 2) * step


def gather_detections(local: Sequence[torch.Tensor], n_total: int, group=None) -> List[torch.Tensor]:
    """Collect every rank's (B_local, post_nms, *) tensors on all ranks, in shard order.

    Off the hot path (614 KB per 256-frame shard).  Shards may be ragged, so each tensor is padded to
    the largest shard, all-gathered, and trimmed with the same split_frames boundaries."""
    world = dist.get_world_size(group)
    bounds = split_frames(n_total, world)
    biggest = max(e - s for s, e in bounds)
    out = []
    for t in local:
        pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out.append(torch.cat([p[: e - s] for p, (s, e) in zip(parts, bounds)], dim=0))
    return out
