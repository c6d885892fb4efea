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


def gather_detections(out_dets: torch.Tensor, out_count: torch.Tensor, world: int, rows: int = 4096):
    """The path's ONE collective: every rank contributes a (B, rows + 1, 5) float32 block -- row 0 of each image carries
    its detection count (int32 bits in column 0), rows 1.. the first ``rows`` boxes -- and receives everybody's
    (replaces the mp.Queue pickling of ``lib/test.py:336-344``).  Returns the gathered (world, B, rows + 1, 5) tensor;
    ``merge_gathered`` unpacks it and RAISES if an image had more than ``rows`` detections (nothing is cut silently)."""
    B = out_dets.shape[0]
    rows = min(rows, out_dets.shape[1])
    pack = torch.empty((B, rows + 1, 5), dtype=torch.float32, device=out_dets.device)
    pack[:, 1:, :] = out_dets[:, :rows, :]
    pack[:, 0, :] = 0.0
    pack[:, 0, 0] = out_count.to(torch.int32).view(torch.float32)
    if world == 1:
        return pack.unsqueeze(0)
    out = torch.empty((world,) + tuple(pack.shape), dtype=torch.float32, device=pack.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, pack)               # ncclAllGather straight into the result tensor
    else:
        dist.all_gather(list(out.unbind(0)), pack)           # gloo (CPU tests)
    return out


def merge_gathered(gathered: torch.Tensor, n_images: int, world: int) -> List[np.ndarray]:
    """Rank-ordered concatenation, trimmed to each rank's real shard size (``lib/test.py:342-344``)."""
    g = gathered.cpu()
    counts = g[:, :, 0, 0].contiguous().view(torch.int32).numpy()
    g = g.numpy()
    rows = g.shape[2] - 1
    out: List[np.ndarray] = []
    for r in range(world):
        a, b = shard_range(n_images, world, r)
        for i in range(b - a):
            c = int(counts[r, i])
            if c > rows:
                raise RuntimeError("image %d of rank %d has %d detections, the gather payload carries %d rows: raise "
                                   "gather_rows" % (i, r, c, rows))
            out.append(g[r, i, 1:1 + c].copy())
    return out
