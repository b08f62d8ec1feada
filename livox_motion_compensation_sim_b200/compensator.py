"""
Host-side mirror of the reference's per-point deskew, ``MotionCompensator``
(/root/reference/livox_mid70_complete_simulator.py = CS, lines 1426-1536), B200-native underneath.

Same constructor (config dict, key ``enable_motion_compensation``), same
``compensate_point_cloud(points, imu_data, frame_start_time, frame_duration_ns)`` signature over
the reference's ``LiDARPoint`` / ``IMUData`` carriers, same no-op rules (CS:1439-1440).  The
reference walks Python dataclass lists point by point (0.02 Mpts/s); here the list API only
converts to flat arrays and back -- use ``compensate_arrays`` / ``compensate_frames`` to stay in
array form.  No CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _trace
from . import _capi as C
from . import ops


@dataclass
class IMUData:                       # CS:97-106
    timestamp: int
    gyro_x: float
    gyro_y: float
    gyro_z: float
    accel_x: float
    accel_y: float
    accel_z: float


@dataclass
class LiDARPoint:                    # CS:120-129
    x: float
    y: float
    z: float
    intensity: int
    timestamp: int
    ring: int
    tag: int


def imu_arrays(imu_data: List[IMUData]):
    """List[IMUData] -> (int64 ts[S], f64 gyro (S,3))."""
    ts = np.fromiter((s.timestamp for s in imu_data), np.int64, count=len(imu_data))
    gy = np.array([[s.gyro_x, s.gyro_y, s.gyro_z] for s in imu_data], np.float64).reshape(-1, 3)
    return ts, gy


class MotionCompensator:
    """Advanced motion compensation using IMU data (CS:1426-1536) on the B200."""

    def __init__(self, config: Dict):
        self.config = config
        self.enable_compensation = config.get('enable_motion_compensation', True)
        self.device = torch.device(config.get('device', 'cuda:0'))

    def _dev(self, a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.device)

    # -- array API ---------------------------------------------------------------------------
    @_trace.traced("MotionCompensator.compensate_arrays")
    def compensate_arrays(self, pts: np.ndarray, ts: np.ndarray, frame_off: np.ndarray, frame_start: np.ndarray,
                          imu_ts: np.ndarray, imu_gyro: np.ndarray, *, lvx2: bool = False,
                          tag: Optional[np.ndarray] = None):
        """pts (N,4) f64 [x y z intensity], ts int64 ns -> compensated (N,4) f64 (+ LVX2 records)."""
        pts = np.asarray(pts, np.float64).reshape(-1, 4)
        if not self.enable_compensation or len(imu_ts) == 0 or len(pts) == 0:      # CS:1439-1440
            out = pts.copy()
            if not lvx2:
                return out, None
        spec = None
        if lvx2:
            spec = ops.ExportSpec(lvx=True, lvx_mode=C.LVX2_OF_OUTPUT,
                                  tag=None if tag is None else self._dev(tag, np.uint8))
        if not self.enable_compensation or len(imu_ts) == 0 or len(pts) == 0:
            if len(pts) == 0:
                return out, np.zeros((0, 14), np.uint8)
            bufs = ops.quantize(self._dev(pts, np.float64), spec)
            bufs.raise_for_flags()
            return out, bufs.lvx14.cpu().numpy()
        out, bufs = ops.deskew_gyro(self._dev(pts, np.float64), self._dev(ts, np.int64),
                                    self._dev(frame_off, np.int64), self._dev(frame_start, np.int64),
                                    self._dev(imu_ts, np.int64), self._dev(imu_gyro, np.float64), export=spec)
        rec = None
        if bufs is not None:
            bufs.raise_for_flags()
            rec = bufs.lvx14.cpu().numpy()
        return out.cpu().numpy(), rec

    # -- the reference's operator (CS:1435-1480) ---------------------------------------------------
    @_trace.traced("MotionCompensator.compensate_point_cloud")
    def compensate_point_cloud(self, points: List[LiDARPoint], imu_data: List[IMUData],
                               frame_start_time: int, frame_duration_ns: int) -> List[LiDARPoint]:
        """Apply motion compensation to point cloud using IMU data (one frame)."""
        if not self.enable_compensation or not imu_data:
            return points
        if not points:
            return []
        pts = np.array([[p.x, p.y, p.z, p.intensity] for p in points], np.float64)
        ts = np.fromiter((p.timestamp for p in points), np.int64, count=len(points))
        imu_ts, imu_gyro = imu_arrays(imu_data)
        out, _ = self.compensate_arrays(pts, ts, np.array([0, len(points)], np.int64),
                                        np.array([frame_start_time], np.int64), imu_ts, imu_gyro)
        return [LiDARPoint(x=float(o[0]), y=float(o[1]), z=float(o[2]), intensity=p.intensity,
                           timestamp=p.timestamp, ring=p.ring, tag=p.tag) for o, p in zip(out, points)]

    # -- batched CS:2086-2105 (_apply_motion_compensation) -----------------------------------------
    @_trace.traced("MotionCompensator.compensate_frames")
    def compensate_frames(self, frames_data: List[Dict], imu_data: List[IMUData]) -> List[Dict]:
        """All frames in one device call; returns new frame dicts with 'motion_compensated': True."""
        if not self.enable_compensation or not imu_data:
            return [dict(f, motion_compensated=True) for f in frames_data]
        counts = np.array([len(f['points']) for f in frames_data], np.int64)
        off = np.zeros(len(frames_data) + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        allp = [p for f in frames_data for p in f['points']]
        pts = np.array([[p.x, p.y, p.z, p.intensity] for p in allp], np.float64).reshape(-1, 4)
        ts = np.fromiter((p.timestamp for p in allp), np.int64, count=len(allp))
        fstart = np.array([f['timestamp'] for f in frames_data], np.int64)
        imu_ts, imu_gyro = imu_arrays(imu_data)
        out, _ = self.compensate_arrays(pts, ts, off, fstart, imu_ts, imu_gyro)
        res = []
        for i, f in enumerate(frames_data):
            o = out[off[i]:off[i + 1]]
            newp = [LiDARPoint(x=float(r[0]), y=float(r[1]), z=float(r[2]), intensity=p.intensity,
                               timestamp=p.timestamp, ring=p.ring, tag=p.tag) for r, p in zip(o, f['points'])]
            g = f.copy(); g['points'] = newp; g['motion_compensated'] = True
            res.append(g)
        return res
