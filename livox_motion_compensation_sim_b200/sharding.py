"""Frame-sharded multi-GPU execution.

Frames are independent, so rank r owns a contiguous frame range chosen by prefix sums of the
per-frame point counts (points, not frames, are balanced).  Contiguity makes the concatenation of
the ranks' outputs equal np.vstack frame-major order (LMC:888) with no permutation: every rank
runs the fused kernel over its point range [p_begin, p_end) and writes at the GLOBAL offsets of a
full-size merged buffer.  The only exchange step is assembling the merged cloud -- an
all-gather(v) of the disjoint slices over NCCL/NVLink (gloo on CPU for the tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .frames import partition_frames


def shard_ranges(frame_off: np.ndarray, world_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(frame cuts[W+1], point cuts[W+1])."""
    cuts = partition_frames(frame_off, world_size)
    return cuts, np.asarray(frame_off, np.int64)[cuts]


def all_gather_merged(buffers: Sequence[torch.Tensor], point_cuts: np.ndarray, group=None) -> None:
    """In-place all-gather(v): on entry rank r's slice [point_cuts[r], point_cuts[r+1]) of every
    buffer (first dim = points) is valid; on exit every rank holds every slice.

    Equal shards use one all_gather_into_tensor per buffer (in place); unequal shards use one
    broadcast per (rank, buffer), issued asynchronously and waited together."""
    W = dist.get_world_size(group)
    if W == 1:
        return
    rank = dist.get_rank(group)
    sizes = np.diff(point_cuts)
    equal = len(set(int(s) for s in sizes)) == 1 and int(point_cuts[0]) == 0
    works = []
    for buf in buffers:
        if buf is None:
            continue
        n = buf.shape[0]
        raw = buf.view(torch.uint8).reshape(n, -1) if n else buf        # bytes are bytes: one dtype for every backend
        if equal and int(point_cuts[-1]) == n:
            mine = raw[int(point_cuts[rank]):int(point_cuts[rank + 1])]
            works.append(dist.all_gather_into_tensor(raw, mine, group=group, async_op=True))
        else:
            for r in range(W):
                b, e = int(point_cuts[r]), int(point_cuts[r + 1])
                if e > b:
                    src = dist.get_global_rank(group, r) if group is not None else r
                    works.append(dist.broadcast(raw[b:e], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
