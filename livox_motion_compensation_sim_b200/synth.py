"""Synthetic Mid-70-shaped streams for the parity tests and bench.py:

  M-1H     36 000 frames x 10 000 pts (1 h at 10 Hz, 3.6e8 pts), figure-eight trajectory sampled at
           200 Hz (720 001 samples) with GPS / IMU noise on
  M-SWEEP  10 k - 1 M points per frame, ~1e8 points

Points are generated on the device (torch.Generator, seeded) because 3.6e8 points take minutes
with NumPy; the small pose / sample tables come from NumPy + SciPy on the host.  All f32 values are
exactly representable in f64, so the oracle sees bit-identical inputs after an up-cast.

Sensor model (CS:63-73): azimuth U(-35.2, 35.2) deg, elevation U(-38.6, 38.6) deg, range U(0.05, 90) m,
intensity U(0,1); per-point time = frame start + i * point_dt_ns (10 us spreads 10 000 points over
the 100 ms frame; the reference's own spacing is 1 us, CS:1048).
Trajectory: the reference's figure-eight (LMC:380-394) + N(0,0.03) m GPS noise (LMC:321) and
N(0,0.01) rad orientation noise (LMC:323).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from . import frames as FR


@dataclass
class Stream:
    n_frames: int
    n_points: int
    frame_off: np.ndarray           # int64[F+1]
    frame_start: np.ndarray         # int64[F] ns
    frame_t: np.ndarray             # f64[F] s
    sample_ts: np.ndarray           # int64[S] ns (200 Hz)
    sample_quat: np.ndarray         # f64 (S,4) xyzw
    sample_pos: np.ndarray          # f64 (S,3)
    sample_euler: np.ndarray        # f64 (S,3)
    seg: np.ndarray                 # f64 (S,20) Mode C table
    gps_t: np.ndarray               # f64[n_t] 5 Hz grid (Mode A trajectory)
    gps_Rt: np.ndarray              # f64 (n_t,12)
    pts: Optional[torch.Tensor] = None       # (N,4) f32/f64 on device
    ts_off: Optional[torch.Tensor] = None    # (N) uint32 on device


def figure_eight(t: np.ndarray, max_speed: float = 12.0, rate: float = 200.0):
    """LMC:380-394 evaluated on time grid t."""
    scale, freq = 30.0, 0.05
    x = scale * np.sin(2 * np.pi * freq * t)
    y = scale * np.sin(4 * np.pi * freq * t)
    z = np.zeros_like(t) + 1.5
    vx, vy = np.gradient(x) * rate, np.gradient(y) * rate
    yaw = np.arctan2(vy, vx)
    roll = np.radians(5) * np.sin(0.5 * yaw) * (np.sqrt(vx ** 2 + vy ** 2) / max_speed)
    return np.column_stack([x, y, z]), np.column_stack([roll, np.zeros_like(t), yaw])


def make_tables(n_frames: int, seed: int, fps: float = 10.0, sample_rate: float = 200.0, gps_rate: float = 5.0):
    rng = np.random.default_rng(seed)
    duration = n_frames / fps
    S = int(round(duration * sample_rate)) + 1
    sample_ts = (np.arange(S, dtype=np.int64) * int(round(1e9 / sample_rate)))
    t = sample_ts * 1e-9
    pos, eul = figure_eight(t, rate=sample_rate)
    pos = pos + rng.normal(0, 0.03, pos.shape)
    eul = eul + rng.normal(0, 0.01, eul.shape)
    quat = Rotation.from_euler('xyz', eul).as_quat()
    seg = FR.slerp_segment_table(quat, pos, sample_ts)
    frame_start = (np.arange(n_frames, dtype=np.int64) * int(round(1e9 / fps)))
    frame_t = frame_start * 1e-9
    n_t = max(int(duration * gps_rate), 2)
    gps_t = np.linspace(0, duration, n_t)
    gidx = np.clip(np.searchsorted(t, gps_t), 0, S - 1)
    gps_Rt = FR.pose_table(pos[gidx], eul[gidx])
    return dict(sample_ts=sample_ts, sample_quat=quat, sample_pos=pos, sample_euler=eul, seg=seg,
                frame_start=frame_start, frame_t=frame_t, gps_t=gps_t, gps_Rt=gps_Rt)


def make_points(n_points: int, seed: int, device, dtype=torch.float32, chunk: int = 1 << 25) -> torch.Tensor:
    """(N,4) sensor-frame points on `device`, values exactly f32-representable."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_points, 4), dtype=dtype, device=device)
    d2r = np.pi / 180.0
    for b in range(0, n_points, chunk):
        n = min(chunk, n_points - b)
        u = torch.rand((n, 4), generator=g, device=device, dtype=torch.float32)
        az = (u[:, 0] * 70.4 - 35.2) * d2r
        el = (u[:, 1] * 77.2 - 38.6) * d2r
        r = u[:, 2] * (90.0 - 0.05) + 0.05
        ce = torch.cos(el)
        o = out[b:b + n]
        o[:, 0] = (r * ce * torch.cos(az)).to(dtype)
        o[:, 1] = (r * ce * torch.sin(az)).to(dtype)
        o[:, 2] = (r * torch.sin(el)).to(dtype)
        o[:, 3] = u[:, 3].to(dtype)
    return out


def make_stream(n_frames: int, pts_per_frame, seed: int, device=None, dtype=torch.float32,
                point_dt_ns: int = 10_000, with_points: bool = True) -> Stream:
    """pts_per_frame: int (uniform) or int array (ragged, zeros allowed)."""
    counts = np.full(n_frames, pts_per_frame, np.int64) if np.isscalar(pts_per_frame) else np.asarray(pts_per_frame, np.int64)
    assert len(counts) == n_frames
    off = np.zeros(n_frames + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    N = int(off[-1])
    tb = make_tables(n_frames, seed)
    st = Stream(n_frames=n_frames, n_points=N, frame_off=off, **tb)
    if with_points:
        st.pts = make_points(N, seed + 1, device, dtype)
        off_d = torch.from_numpy(off).to(device)
        st.ts_off = torch.empty(N, dtype=torch.uint32, device=device)
        chunk = 1 << 26
        for b in range(0, N, chunk):
            idx = torch.arange(b, min(b + chunk, N), device=device, dtype=torch.int64)
            fidx = torch.searchsorted(off_d, idx, right=True) - 1
            within = idx - off_d[fidx]
            st.ts_off[b:b + len(idx)] = (within * point_dt_ns).clamp_(0, 2 ** 32 - 1).to(torch.uint32)
    return st
