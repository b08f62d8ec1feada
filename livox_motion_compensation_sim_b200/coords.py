""" Coordinate-frame chain: host mirror of the reference's ``CoordinateTransformer``
(/root/reference/livox_mid70_complete_simulator.py = CS, lines 153-233), with the per-point work
on the B200.

A 4x4 homogeneous transform applied to (n,3) points is the Mode A kernel with the pose row
[T[:3,:3] row-major | T[:3,3]]: the reference's ``(T @ homog.T).T`` runs through dgemm as
fma(T3,1, fma(T2,z, fma(T1,y, T0*x))) and fma(t,1,acc) == acc + t, so the result is bit-identical
(golden ``coord_chain.npz``) for n >= 2 points; a single point goes through a 4-term gemv in the
reference, (T0*x + T2*z) + (T1*y + T3) with unfused products -- ``lmc_transform_homog_*`` implements
both orders, so the per-point calls of ``_transform_coordinates`` (CS:2107-2163) are reproduced bit
for bit as well (``transform_frames``).  The 4x4 matrices themselves are tiny host-side bookkeeping
built with the same NumPy calls as CS:176-212.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _capi as C
from . import ops

from . import _trace


class CoordinateSystem:                 # CS:145-151
    SENSOR = "sensor"
    VEHICLE = "vehicle"
    LOCAL = "local"
    UTM = "utm"
    WGS84 = "wgs84"


def pose_rows_from_matrices(T: np.ndarray) -> np.ndarray:
    """(F,4,4) homogeneous matrices -> (F,12) pose rows for lmc_align_rigid_*."""
    T = np.asarray(T, np.float64).reshape(-1, 4, 4)
    return np.ascontiguousarray(np.concatenate([T[:, :3, :3].reshape(-1, 9), T[:, :3, 3]], axis=1))


def fold_chain(T_chain: np.ndarray, pose_Rt: np.ndarray) -> np.ndarray:
    """Pre-multiply a fixed frame chain into a per-frame pose table: the fused kernels then emit the
    target frame directly at zero extra traffic.  (One rounding of the product instead of two
    successive transforms: agrees with applying them one after the other to ~1e-12 m, not bitwise.)"""
    pose_Rt = np.asarray(pose_Rt, np.float64).reshape(-1, 12)
    F = len(pose_Rt)
    Tp = np.zeros((F, 4, 4)); Tp[:, 3, 3] = 1.0
    Tp[:, :3, :3] = pose_Rt[:, :9].reshape(F, 3, 3); Tp[:, :3, 3] = pose_Rt[:, 9:]
    return pose_rows_from_matrices(np.matmul(np.asarray(T_chain, np.float64), Tp))


class CoordinateTransformer:
    """Advanced coordinate system transformations (CS:153-233), points transformed on the device."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        self.transformations: Dict[Tuple[str, str], np.ndarray] = {}
        self._setup_default_transformations()

    def _setup_default_transformations(self):               # CS:160-174
        identity = np.eye(4)
        self.transformations[(CoordinateSystem.SENSOR, CoordinateSystem.SENSOR)] = identity
        self.transformations[(CoordinateSystem.VEHICLE, CoordinateSystem.VEHICLE)] = identity
        self.transformations[(CoordinateSystem.SENSOR, CoordinateSystem.VEHICLE)] = np.array(
            [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.5], [0, 0, 0, 1]])

    def set_transformation(self, from_frame: str, to_frame: str, translation: List[float], rotation: List[float]):
        T = self._create_transform_matrix(translation, rotation)       # CS:176-186
        self.transformations[(from_frame, to_frame)] = T
        self.transformations[(to_frame, from_frame)] = np.linalg.inv(T)

    def _create_transform_matrix(self, translation, rotation) -> np.ndarray:   # CS:188-212 (same NumPy calls)
        roll, pitch, yaw = rotation
        Rx = np.array([[1, 0, 0], [0, np.cos(roll), -np.sin(roll)], [0, np.sin(roll), np.cos(roll)]])
        Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        T = np.eye(4)
        T[:3, :3] = Rz @ Ry @ Rx
        T[:3, 3] = translation
        return T

    @_trace.traced("CoordinateTransformer.transform_points")
    def transform_points(self, points: np.ndarray, from_frame: str, to_frame: str) -> np.ndarray:
        """CS:214-233: (n,3) -> (n,3) in the target frame; unknown pair -> points returned unchanged.
        n == 1 reproduces the reference's single-point (gemv) summation order, n >= 2 the dgemm order."""
        if (from_frame, to_frame) not in self.transformations:
            return points
        points = np.asarray(points, np.float64)
        if points.ndim != 2 or points.shape[1] != 3:
            # CS:223-228 treats any other width as already homogeneous (column 3 multiplies the translation, other widths fail
            # in the matmul); none of the reference's own callers does that, and the device kernel fixes w = 1 -- refuse
            # instead of silently computing something else
            raise ValueError(f"transform_points expects (n, 3) points, got {points.shape}")
        n = len(points)
        if n == 0:
            return points[:, :3].copy()
        p4 = np.zeros((n, 4)); p4[:, :3] = points[:, :3]
        d = torch.from_numpy(p4).to(self.device)
        out = ops.transform_homog(d, self.transformations[(from_frame, to_frame)], C.HOMOG_SINGLE if n == 1 else C.HOMOG_BATCH)
        return out.cpu().numpy()[:, :3]

    # -- batched CS:2107-2163 (_transform_coordinates) ---------------------------------------------
    @_trace.traced("CoordinateTransformer.transform_frames")
    def transform_frames(self, frames_data: List[Dict], target_system: str, utm_offsets: Optional[np.ndarray] = None) -> List[Dict]:
        """All frames in one device call.  The reference transforms ONE point per ``transform_points`` call
        (CS:2117-2138), so the single-point summation order applies to every point.  For the UTM target
        the reference adds a per-frame ``[utm_x, utm_y, 0]`` taken from the ``utm`` package (CS:2122-2131);
        that package is not a dependency here, so the caller passes the (F,2) easting / northing per frame
        (``utm_offsets``; None = what the reference does when ``utm`` is missing: points unchanged).
        Frames hold ``List[LiDARPoint]``; returns new frame dicts with ``'coordinate_system'`` set."""
        from .compensator import LiDARPoint
        counts = np.array([len(f['points']) for f in frames_data], np.int64)
        off = np.zeros(len(frames_data) + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        allp = [p for f in frames_data for p in f['points']]
        xyz = np.zeros((len(allp), 4), np.float64)
        if allp:
            xyz[:, :3] = [(p.x, p.y, p.z) for p in allp]
        if target_system == CoordinateSystem.UTM:
            if utm_offsets is not None and len(allp):
                o = np.asarray(utm_offsets, np.float64).reshape(len(frames_data), 2)
                fr = np.repeat(np.arange(len(frames_data)), counts)
                pose = np.zeros((len(frames_data), 12)); pose[:, [0, 4, 8]] = 1.0; pose[:, 9:11] = o
                dd = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)   # noqa: E731
                out, _ = ops.align_rigid(dd(xyz), dd(off), dd(pose))       # 1*x (+0*y +0*z) + utm: exactly point + offset
                xyz = out.cpu().numpy()
                del fr
        elif len(allp):
            key = (CoordinateSystem.SENSOR, target_system)
            if key in self.transformations:                               # unknown pair: unchanged (CS:216-218)
                xyz = ops.transform_homog(torch.from_numpy(xyz).to(self.device), self.transformations[key], C.HOMOG_SINGLE).cpu().numpy()
        res = []
        for i, f in enumerate(frames_data):
            o = xyz[off[i]:off[i + 1]]
            g = f.copy()
            g['points'] = [LiDARPoint(x=r[0], y=r[1], z=r[2], intensity=p.intensity, timestamp=p.timestamp, ring=p.ring, tag=p.tag)
                           for r, p in zip(o, f['points'])]
            g['coordinate_system'] = target_system
            res.append(g)
        return res
