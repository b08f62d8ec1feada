"""Frame-sharded multi-GPU execution.

Frames are independent, so rank r owns a contiguous frame range chosen by prefix sums of the
per-frame point counts (points, not frames, are balanced).  Contiguity makes the concatenation of
the ranks' outputs equal np.vstack frame-major order (LMC:888) with no permutation: every rank
runs the fused kernel over its point range [p_begin, p_end) and writes at the GLOBAL offsets of a
full-size merged buffer.  The only exchange step is assembling the merged cloud -- an
all-gather(v) of the disjoint slices over NCCL/NVLink (gloo on CPU for the tests)."""
from __future__ import annotations

from typing import Sequence, Tuple

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


class SymmetricMerged:
    """Merged-cloud buffers in NVLink-symmetric memory (torch.distributed._symmetric_memory): every rank
    holds a full-size copy of the aligned cloud (and LVX records) and can address the other ranks' copies
    directly.  Passing ``peer_ptrs()`` (and, where the fabric offers it, ``mc_ptrs()``) to the fused operators
    makes the kernel epilogue store each result into every rank's copy -- remote st.global over NVLink, or one
    multimem.st per result that the NVSwitch replicates -- so the all-gather overlaps the transform instead of
    following it; ``barrier()`` after the launch makes the remote writes visible.  torch is only the
    allocator / rendezvous plumbing here -- the data moves inside the sm_100a kernel."""

    def __init__(self, n_points: int, device, *, dtype=torch.float32, lvx: bool = True, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.out = symm.empty((n_points, 4), dtype=dtype, device=device)
        self.h_out = symm.rendezvous(self.out, self.group)
        self.lvx14 = self.h_lvx = None
        if lvx:
            self.lvx14 = symm.empty((n_points, 14), dtype=torch.uint8, device=device)
            self.h_lvx = symm.rendezvous(self.lvx14, self.group)

    @staticmethod
    def _off(h) -> int:
        """byte offset of the tensor inside its symmetric allocation (0 unless a memory pool packs several tensors into one)"""
        o = getattr(h, "offset", 0)
        return int(o) if isinstance(o, int) else 0

    def peer_ptrs(self):
        others = [r for r in range(self.world) if r != self.rank]
        po = [int(self.h_out.buffer_ptrs[r]) + self._off(self.h_out) for r in others]
        pl = [int(self.h_lvx.buffer_ptrs[r]) + self._off(self.h_lvx) for r in others] if self.h_lvx is not None else []
        return po, pl

    def mc_ptrs(self):
        """(multicast address of out, of lvx14), or (0, 0) when the buffers have no NVSwitch multicast mapping."""
        if self.h_lvx is None:
            return 0, 0
        mo, ml = int(getattr(self.h_out, "multicast_ptr", 0) or 0), int(getattr(self.h_lvx, "multicast_ptr", 0) or 0)
        return (mo + self._off(self.h_out), ml + self._off(self.h_lvx)) if (mo and ml) else (0, 0)

    def barrier(self):
        """Cross-rank barrier ON THE CURRENT STREAM (a signal-pad kernel: every rank's earlier work on its stream,
        the fused kernel's remote stores included, is complete and visible before any rank's later work starts).
        Nothing blocks on the host."""
        self.h_out.barrier()
