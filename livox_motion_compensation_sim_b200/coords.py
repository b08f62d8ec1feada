""" Coordinate-frame chain: host mirror of the reference's ``CoordinateTransformer``
(/root/reference/livox_mid70_complete_simulator.py = CS, lines 153-233), with the per-point work
on the B200.

A 4x4 homogeneous transform applied to (n,3) points is the Mode A kernel with the pose row
[T[:3,:3] row-major | T[:3,3]]: the reference's ``(T @ homog.T).T`` runs through dgemm as
fma(T3,1, fma(T2,z, fma(T1,y, T0*x))) and fma(t,1,acc) == acc + t, so the result is bit-identical
(golden ``coord_chain.npz``) for n >= 2 points; a single point goes through a 4-term gemv in the
reference whose summation order differs, so that one case agrees to 1 ulp instead.  The 4x4 matrices themselves are tiny host-side bookkeeping built
with the same NumPy calls as CS:176-212.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops


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

    def transform_points(self, points: np.ndarray, from_frame: str, to_frame: str) -> np.ndarray:
        """CS:214-233: (n,3) -> (n,3) in the target frame; unknown pair -> points returned unchanged."""
        if (from_frame, to_frame) not in self.transformations:
            return points
        points = np.asarray(points, np.float64)
        n = len(points)
        if n == 0:
            return points[:, :3].copy()
        p4 = np.zeros((n, 4)); p4[:, :3] = points[:, :3]
        pose = pose_rows_from_matrices(self.transformations[(from_frame, to_frame)])
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)   # noqa: E731
        out, _ = ops.align_rigid(d(p4), torch.tensor([0, n], dtype=torch.int64, device=self.device), d(pose))
        return out.cpu().numpy()[:, :3]
