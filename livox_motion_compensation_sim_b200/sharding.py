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


# ------------------------------------------------------------------------------------------------------------------
# Replication-free file production: every rank turns ITS frame range into ITS byte range of the
# output file and writes it at the right offset with pwrite().  Nothing but a few integers crosses ranks:
#   LVX v1.1   byte offsets are closed-form in the frame sizes (lvx.frame_layout)              -> no exchange at all
#   LAS 1.2    record offsets are closed-form; the header needs the global integer extremes    -> all-reduce of 6 int32
#   ASCII PCD  '%.6f' lines have variable length                                               -> all-gather of W text sizes
# The concatenation of the ranks' byte ranges IS the file the single-GPU path builds (tests/multi_gpu_fused_merge.py,
# bench.py config.merge.sharded_files).
# ------------------------------------------------------------------------------------------------------------------
def file_offsets(sizes, header_bytes: int = 0) -> np.ndarray:
    """Byte offset of every rank's range in a file that starts with `header_bytes` bytes owned by rank 0:
    off[r] = header_bytes + sum(sizes[:r]) for r >= 1, off[0] = 0 (rank 0's range includes the header)."""
    sizes = np.asarray(sizes, np.int64)
    off = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum(sizes, out=off[1:])
    off[1:] += header_bytes
    return off                                         # off[W] = file size


def all_gather_sizes(n_local: int, device, group=None) -> np.ndarray:
    """int64[W]: every rank's n_local (one tiny all-gather; device tensors so NCCL and gloo both work)."""
    W = dist.get_world_size(group)
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    allv = torch.empty(W, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    return allv.cpu().numpy()


def pwrite_range(path: str, offset: int, data) -> int:
    """Write `data` (bytes-like / uint8 host array) at byte `offset` of `path` without truncating it: every rank writes
    its own disjoint range of the same file.  Returns the number of bytes written."""
    import os
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        mv = memoryview(data).cast('B')
        done = 0
        while done < len(mv):
            done += os.pwrite(fd, mv[done:done + (1 << 30)], offset + done)
        return done
    finally:
        os.close(fd)


def lvx_v11_shard(pts: torch.Tensor, frame_off: np.ndarray, frame_time: np.ndarray, frame_id: np.ndarray, f_begin: int, f_end: int):
    """This rank's byte range of the LVX v1.1 file (LMC:58-272) built on its GPU from the GLOBAL frame table.
    Returns (uint8 device tensor, file offset of its first byte, status tensor)."""
    from . import ops
    from .lvx import frame_layout
    dev = pts.device
    frame_off = np.asarray(frame_off, np.int64)
    _, fpos = frame_layout(frame_off)
    pos0 = 0 if f_begin == 0 else int(fpos[f_begin])
    nbytes = int(fpos[f_end]) - pos0 if f_end > f_begin else 0
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a, dtype=t)).to(dev)          # noqa: E731
    counts = np.diff(frame_off[f_begin:f_end + 1])
    out, status = ops.build_lvx_v11_range(pts, d(frame_off, np.int64), d(fpos, np.int64), d(frame_time, np.float64), d(frame_id, np.int64),
                                          f_begin, f_end, pos0, nbytes, int(counts.max()) if len(counts) else 0)
    return out, pos0, status


def las_pf3_shard(pts: torch.Tensor, p_begin: int, p_end: int, rank: int, group=None, **las_kw):
    """This rank's byte range of the LAS 1.2 / PF3 file (LMC:950-963): records of [p_begin, p_end); rank 0's range also
    carries the header, filled in after the six integer extremes have been min / max-reduced over the ranks.
    Returns (uint8 device tensor, file offset of its first byte, status tensor)."""
    from . import _capi as C
    from . import ops
    hdr_kw = {k: las_kw.pop(k) for k in ("year", "day_of_year") if k in las_kw}
    pos0 = 0 if rank == 0 else C.LAS_HEADER_BYTES + C.LAS_RECORD_BYTES * int(p_begin)
    out, mm, status = ops.las_pf3_records(pts, p_begin, p_end, pos0, **las_kw)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        lo, hi = mm[0::2].contiguous(), mm[1::2].contiguous()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        mm = torch.stack([lo, hi], dim=1).reshape(6).contiguous()
    if rank == 0:
        ops.las_pf3_header(out, mm, pts.shape[0], scale=las_kw.get("scale", (0.01,) * 3), offset=las_kw.get("offset", (0.0,) * 3), **hdr_kw)
    return out, pos0, status


def pcd_ascii_shard(pts_shard: torch.Tensor, n_total: int, rank: int, group=None):
    """This rank's byte range of the ASCII PCD file (LMC:932-948): the '%.6f' lines of its own points; rank 0's range is
    preceded by the header (returned separately as bytes).  One all-gather of the W text sizes places the ranges.
    Returns (uint8 device text tensor, file offset of its first byte, header bytes or b'', status tensor)."""
    from . import ops
    from .simulator import LiDARMotionSimulator
    text, status = ops.pcd_ascii_body(pts_shard)
    header = LiDARMotionSimulator._pcd_header(int(n_total))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        sizes = all_gather_sizes(int(text.numel()), pts_shard.device, group)
    else:
        sizes = np.array([int(text.numel())], np.int64)
    off = file_offsets(sizes, len(header))
    pos0 = int(off[rank]) + (len(header) if rank == 0 else 0)
    return text, pos0, (header if rank == 0 else b""), status
