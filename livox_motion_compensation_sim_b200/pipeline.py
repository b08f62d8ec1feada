"""Host-buffer entry of the hot path: pinned host arrays in, pinned host arrays out.

The stream is cut into chunks of whole frames; chunk i+1's host->device copy, chunk i's fused
kernel and chunk i-1's device->host copy run concurrently on three CUDA streams over
double-buffered device staging, so the end-to-end rate is the PCIe rate of the larger direction
rather than the sum of the three stages.  The kernel calls go through the C ABI
(lmc_deskew_slerp_f32 / lmc_align_rigid_f32) exactly as in ops.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _trace
from . import _capi as C
from . import ops


def bind_host_to_gpu(device_index: int) -> Optional[list]:
    """Pin this process to the CPUs NVML reports as local to the GPU (its NUMA node), so that pinned host
    buffers allocated afterwards land in memory attached to the GPU's PCIe root -- on a two-socket 8-GPU
    box un-bound ranks otherwise share one socket's memory and the host<->device copies collapse.
    Returns the CPU list, or None if NVML / affinity calls are unavailable (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:                                      # noqa: BLE001
        pass
    return None


@dataclass
class HostStream:
    """Pinned host buffers of one stream (inputs + outputs)."""
    pts: torch.Tensor                 # (N,4) f32 pinned
    ts_off: Optional[torch.Tensor]    # (N) uint32 pinned (Mode C/B)
    out: torch.Tensor                 # (N,4) f32 pinned
    lvx14: Optional[torch.Tensor]     # (N,14) uint8 pinned
    frame_off: np.ndarray
    frame_start: np.ndarray


class StreamingAligner:
    """Chunked H2D -> fused kernel -> D2H pipeline for the f32 throughput layout."""

    def __init__(self, device, frame_off: np.ndarray, frame_start: np.ndarray, *, mode: str = "slerp",
                 sample_ts: Optional[torch.Tensor] = None, seg: Optional[torch.Tensor] = None,
                 pose_Rt: Optional[torch.Tensor] = None, chunk_points: int = 1 << 24, lvx: bool = True,
                 pose_samples: Optional[tuple] = None, nbuf: int = 3):
        """pose_samples = (quat_xyzw (S,4) f64, pos (S,3) f64, ts (S) int64) as pinned HOST tensors: the pose stream then
        travels with the points -- every run() uploads it and builds the segment table on the device
        (lmc_build_slerp_table) instead of taking a host-built ``seg``."""
        self.device = torch.device(device)
        self.mode, self.lvx = mode, lvx
        self.frame_off = np.asarray(frame_off, np.int64)
        self.frame_start = np.asarray(frame_start, np.int64)
        self.sample_ts, self.seg, self.pose_Rt = sample_ts, seg, pose_Rt
        self.pose_samples = pose_samples
        if pose_samples is not None:
            S = pose_samples[0].shape[0]
            self.d_quat = torch.empty((S, 4), dtype=torch.float64, device=self.device)
            self.d_pos = torch.empty((S, 3), dtype=torch.float64, device=self.device)
            self.sample_ts = torch.empty(S, dtype=torch.int64, device=self.device)
            self.seg = torch.empty((S, 22), dtype=torch.float64, device=self.device)
        F = len(self.frame_off) - 1
        # chunk boundaries at whole frames, ~chunk_points each
        cuts = [0]
        while cuts[-1] < F:
            target = self.frame_off[cuts[-1]] + chunk_points
            nxt = int(np.searchsorted(self.frame_off, target, side='right') - 1)
            nxt = max(nxt, cuts[-1] + 1)
            cuts.append(min(nxt, F))
        self.cuts = cuts
        self.max_pts = int(max(self.frame_off[b] - self.frame_off[a] for a, b in zip(cuts[:-1], cuts[1:])))
        self.max_frames = int(max(b - a for a, b in zip(cuts[:-1], cuts[1:])))
        dev = self.device
        self.nbuf = max(2, int(nbuf))                               # staging buffer sets in flight (H2D | kernel | D2H)
        self.d_pts = [torch.empty((self.max_pts, 4), dtype=torch.float32, device=dev) for _ in range(self.nbuf)]
        self.d_ts = [torch.empty(self.max_pts, dtype=torch.uint32, device=dev) for _ in range(self.nbuf)]
        self.d_out = [torch.empty((self.max_pts, 4), dtype=torch.float32, device=dev) for _ in range(self.nbuf)]
        self.d_lvx = [torch.empty((self.max_pts, 14), dtype=torch.uint8, device=dev) for _ in range(self.nbuf)] if lvx else None
        self.d_status = torch.zeros(1, dtype=torch.int32, device=dev)
        # per-chunk CSR / frame-start tables are tiny: upload them all once
        self.d_off, self.d_fs, self.d_pose = [], [], []
        for a, b in zip(cuts[:-1], cuts[1:]):
            self.d_off.append(torch.from_numpy(self.frame_off[a:b + 1] - self.frame_off[a]).to(dev))
            self.d_fs.append(torch.from_numpy(np.ascontiguousarray(self.frame_start[a:b])).to(dev))
            self.d_pose.append(None if pose_Rt is None else pose_Rt[a:b].contiguous())
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @_trace.traced("StreamingAligner.run")
    def run(self, hs: HostStream) -> None:
        """One pass over the whole host stream.  Returns after all copies are enqueued; call
        torch.cuda.synchronize() (or wait on the streams) before reading hs.out / hs.lvx14."""
        cuts, off = self.cuts, self.frame_off
        ev_in: List[torch.cuda.Event] = [torch.cuda.Event() for _ in range(self.nbuf)]
        ev_k: List[torch.cuda.Event] = [torch.cuda.Event() for _ in range(self.nbuf)]
        ev_out: List[Optional[torch.cuda.Event]] = [None] * self.nbuf
        self.launches = self.h2d_bytes = self.d2h_bytes = 0
        cur = torch.cuda.current_stream(self.device)
        self.d_status.zero_()                                        # NaN / overflow flags of THIS pass (flags() / raise_for_flags())
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(cur)
        if self.pose_samples is not None:                           # pose stream: H2D + segment table on the device
            hq, hp, ht = self.pose_samples
            with torch.cuda.stream(self.s_in):
                self.d_quat.copy_(hq, non_blocking=True); self.d_pos.copy_(hp, non_blocking=True); self.sample_ts.copy_(ht, non_blocking=True)
                self.h2d_bytes += hq.numel() * 8 + hp.numel() * 8 + ht.numel() * 8
                ev_pose = torch.cuda.Event(); ev_pose.record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(ev_pose)
                ops.build_slerp_table(self.d_quat, self.d_pos, self.sample_ts, out=self.seg)
                self.launches += 1
        for ci, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            sl = ci % self.nbuf
            p0, p1 = int(off[a]), int(off[b])
            n = p1 - p0
            if n == 0:
                continue
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(ev_k[sl])                     # staging input free once its kernel ran
                self.d_pts[sl][:n].copy_(hs.pts[p0:p1], non_blocking=True)
                self.h2d_bytes += n * 16
                if self.mode != "rigid":
                    self.d_ts[sl][:n].copy_(hs.ts_off[p0:p1], non_blocking=True)
                    self.h2d_bytes += n * 4
                ev_in[sl].record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(ev_in[sl])
                if ev_out[sl] is not None:
                    self.s_k.wait_event(ev_out[sl])                # staging output drained by the previous D2H
                spec = None
                if self.lvx:
                    spec = ops.ExportSpec(lvx=True, lvx_mode=C.LVX_TYPE2_OF_INPUT,
                                          into=ops.ExportBuffers(lvx14=self.d_lvx[sl][:n], status=self.d_status))
                if self.mode == "rigid":
                    ops.align_rigid(self.d_pts[sl][:n], self.d_off[ci], self.d_pose[ci], out=self.d_out[sl][:n], export=spec)
                else:
                    ops.deskew_slerp(self.d_pts[sl][:n], self.d_ts[sl][:n], self.d_off[ci], self.d_fs[ci],
                                     self.sample_ts, self.seg, out=self.d_out[sl][:n], export=spec)
                self.launches += 1
                ev_k[sl].record(self.s_k)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_k[sl])
                hs.out[p0:p1].copy_(self.d_out[sl][:n], non_blocking=True)
                self.d2h_bytes += n * 16
                if self.lvx:
                    hs.lvx14[p0:p1].copy_(self.d_lvx[sl][:n], non_blocking=True)
                    self.d2h_bytes += n * 14
                e = torch.cuda.Event()
                e.record(self.s_out)
                ev_out[sl] = e
        for s in (self.s_in, self.s_k, self.s_out):
            cur.wait_stream(s)

    def flags(self) -> int:
        """NaN / overflow bits (C.FLAG_*) of the last run(); synchronises the device."""
        torch.cuda.synchronize(self.device)
        return int(self.d_status.item())

    def raise_for_flags(self) -> None:
        """What the reference's packer would have raised on the same stream (LMC:257 int(nan) -> ValueError)."""
        f = self.flags()
        if f & C.FLAG_NAN:
            raise ValueError("cannot convert float NaN to integer")
        if f & C.FLAG_OVERFLOW:
            raise OverflowError("LVX record field out of range")
