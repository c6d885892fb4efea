"""Multi-GPU: one process per GPU, images sharded, ONE collective at the end.

The reference shards the test set over ``multiprocessing.Process`` workers in contiguous index ranges of
``ceil(N / n_gpu)`` and ships results back through an ``mp.Queue`` (pickle over a pipe), re-ordered by rank
(``lib/test.py:324-344``).  Here each rank runs the same partition and the per-image result buffers
(fixed capacity + counts) are exchanged with a single ``all_gather`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests); the payload is a few MB, so the collective is latency-bound and needs no overlap.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """``lib/test.py:329-335``: rank r owns [r*ceil(N/G), min((r+1)*ceil(N/G), N))."""
    per = int(np.ceil(1.0 * n_images / world))
    return min(per * rank, n_images), min(per * (rank + 1), n_images)


def gather_detections(out_dets: torch.Tensor, out_count: torch.Tensor, world: int):
    """all_gather of (B, cap, 5) boxes and (B,) counts -> lists ordered by rank (device tensors)."""
    if world == 1:
        return [out_dets], [out_count]
    dets = [torch.empty_like(out_dets) for _ in range(world)]
    cnts = [torch.empty_like(out_count) for _ in range(world)]
    dist.all_gather(dets, out_dets.contiguous())
    dist.all_gather(cnts, out_count.contiguous())
    return dets, cnts


def merge_gathered(dets: List[torch.Tensor], cnts: List[torch.Tensor], n_images: int, world: int) -> List[np.ndarray]:
    """Rank-ordered concatenation, trimmed to each rank's real shard size (``lib/test.py:342-344``)."""
    out: List[np.ndarray] = []
    for r in range(world):
        a, b = shard_range(n_images, world, r)
        d = dets[r].cpu().numpy()
        c = cnts[r].cpu().numpy()
        for i in range(b - a):
            out.append(d[i, :int(c[i])].copy())
    return out
