"""Host-side plumbing around the device path: flattening the reference's per-frame arrays into
one frame-major buffer + CSR offsets, and building the small per-frame / per-sample pose tables.

The pose tables are tiny (<= 720 001 rows for a 1 h stream at 200 Hz) and are built on the host
with SciPy exactly as the reference builds its rotation matrices (LMC:774
``R.from_euler('xyz', rpy).as_matrix()``), so the f64 matrices the kernels consume are bit-identical
to the reference's.  Everything per-point happens on the GPU.

LMC = lidar_motion_compensation.py        CS = livox_mid70_complete_simulator.py
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Sequence, Tuple

import numpy as np
from scipy.spatial.transform import Rotation

SEG_STRIDE = 22        # doubles per Mode C sample row (include/lmc_b200.h)


def frame_offsets(frames: Sequence[np.ndarray]) -> np.ndarray:
    """int64 CSR offsets[F+1] of a list of per-frame arrays."""
    counts = np.fromiter((len(f) for f in frames), dtype=np.int64, count=len(frames))
    off = np.zeros(len(frames) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return off


def _host_threads() -> int:
    try:
        return max(1, min(16, len(os.sched_getaffinity(0))))
    except AttributeError:                                   # pragma: no cover
        return max(1, min(16, os.cpu_count() or 1))


def _addr(a: np.ndarray) -> int:
    """Address of an array's first byte (the cheapest route CPython offers: ~0.8 us per array)."""
    if a.nbytes == 0:
        return 0
    try:
        return ctypes.addressof(ctypes.c_char.from_buffer(a))
    except (TypeError, ValueError):                          # read-only buffer
        return a.__array_interface__['data'][0]


def flatten_frames_into(frames: Sequence[np.ndarray], flat: np.ndarray) -> None:
    """Frame-major concatenation into a caller-provided (N,4) buffer (e.g. a pinned staging array).

    Frames that already have the buffer's dtype and are C-contiguous (what the reference's scanner produces) are
    packed by liblmc_b200's multi-threaded ``lmc_host_gather``; anything else goes through one NumPy pass."""
    if not len(flat):
        return
    arrs = [np.asarray(f) for f in frames]
    if flat.flags.c_contiguous and all(a.dtype == flat.dtype and a.flags.c_contiguous for a in arrs):
        from . import _capi as C
        n = len(arrs)
        ptrs = np.fromiter((_addr(a) for a in arrs), dtype=np.uintp, count=n)
        boff = np.zeros(n + 1, np.int64)
        np.cumsum(np.fromiter((a.nbytes for a in arrs), dtype=np.int64, count=n), out=boff[1:])
        if int(boff[-1]) != flat.nbytes:
            raise ValueError(f"frames hold {int(boff[-1])} bytes, the buffer {flat.nbytes}")
        C.check(C.lib().lmc_host_gather(ptrs.ctypes.data, boff.ctypes.data, n, flat.ctypes.data, _host_threads()))
        return
    np.concatenate([a.reshape(-1, 4) for a in arrs], axis=0, out=flat, casting='unsafe')


def legacy_normal(scale: float, n: int, out: np.ndarray = None) -> np.ndarray:
    """``np.random.normal(0, scale, n)`` drawn from NumPy's GLOBAL legacy generator -- bit-identical values, and the
    generator is left exactly where NumPy would have left it -- through ``lmc_host_legacy_normal``: the word stream
    and the rejection loop stay sequential, the sqrt / log part runs on all host threads (3x faster than NumPy's
    13 ns per sample, which is most of a whole scan + align run on the device).  ``out``: optional flat float64
    buffer of n elements (e.g. pinned staging)."""
    from . import _capi as C
    if scale < 0:
        raise ValueError("scale < 0")                          # as NumPy
    kind, key, pos, has_gauss, cached = np.random.get_state()
    if kind != 'MT19937':                                       # pragma: no cover  (the legacy global generator is always MT19937)
        raise RuntimeError(f"unexpected global bit generator {kind}")
    key = np.ascontiguousarray(key, np.uint32).copy()
    st_pos, st_has, st_cached = ctypes.c_int32(int(pos)), ctypes.c_int32(int(has_gauss)), ctypes.c_double(float(cached))
    if out is None:
        out = np.empty(int(n), np.float64)
    if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != int(n):
        raise ValueError("out: flat C-contiguous float64 buffer of n elements expected")
    C.check(C.lib().lmc_host_legacy_normal(key.ctypes.data, ctypes.addressof(st_pos), ctypes.addressof(st_has), ctypes.addressof(st_cached),
                                           0.0, float(scale), int(n), out.ctypes.data, _host_threads()))
    np.random.set_state((kind, key, st_pos.value, st_has.value, st_cached.value))
    return out


def flatten_frames(frames: Sequence[np.ndarray], dtype=np.float64) -> Tuple[np.ndarray, np.ndarray]:
    """list of (n_f,4) arrays -> ((N,4) frame-major array, int64 CSR offsets[F+1]).

    Frame-major concatenation is np.vstack order (LMC:888); empty frames (``np.array([]).reshape(0,4)``,
    LMC:719/748) become zero-length rows."""
    counts = np.fromiter((len(f) for f in frames), dtype=np.int64, count=len(frames))
    off = np.zeros(len(frames) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    flat = np.empty((int(off[-1]), 4), dtype)
    if off[-1]:
        np.concatenate([np.asarray(f).reshape(-1, 4) for f in frames], axis=0, out=flat, casting='unsafe')   # one C-level pass
    return flat, off


def split_frames(flat: np.ndarray, frame_off: np.ndarray) -> List[np.ndarray]:
    """Inverse of flatten_frames: per-frame views (no copies)."""
    return [flat[frame_off[i]:frame_off[i + 1]] for i in range(len(frame_off) - 1)]


def pose_table(position: np.ndarray, euler_xyz: np.ndarray) -> np.ndarray:
    """(n,12) f64 rows [R row-major | t] with R = SciPy from_euler('xyz') as at LMC:774."""
    position = np.asarray(position, np.float64).reshape(-1, 3)
    euler_xyz = np.asarray(euler_xyz, np.float64).reshape(-1, 3)
    Rm = Rotation.from_euler('xyz', euler_xyz).as_matrix().reshape(-1, 9)
    return np.ascontiguousarray(np.concatenate([Rm, position], axis=1))


def lidar_frame_times(duration: float, lidar_fps: float) -> np.ndarray:
    """LMC:792-793 (linspace WITH endpoint: dt = duration/(n-1), not 1/fps)."""
    return np.linspace(0, duration, int(duration * lidar_fps))


def slerp_segment_table(sample_quat_xyzw: np.ndarray, sample_pos: np.ndarray, sample_ts: np.ndarray) -> np.ndarray:
    """Per-sample table consumed by lmc_deskew_slerp_*: for sample k
    [R_k (9) | pos_k (3) | unit axis of R_k^-1 R_{k+1} (3) | angle | pos_{k+1}-pos_k (3) | 1/(t_{k+1}-t_k) | t_k bits | (t_{k+1}-t_k) bits]."""
    q = np.asarray(sample_quat_xyzw, np.float64).reshape(-1, 4)
    pos = np.asarray(sample_pos, np.float64).reshape(-1, 3)
    S = len(q)
    rot = Rotation.from_quat(q)
    seg = np.zeros((S, SEG_STRIDE), np.float64)
    seg[:, 0:9] = rot.as_matrix().reshape(S, 9)
    seg[:, 9:12] = pos
    seg[:, 12] = 1.0
    if S > 1:
        rv = (rot[:-1].inv() * rot[1:]).as_rotvec()
        th = np.linalg.norm(rv, axis=1)
        nz = th > 0
        seg[:-1][nz, 12:15] = rv[nz] / th[nz, None]
        seg[:-1, 15] = th
        seg[:-1, 16:19] = pos[1:] - pos[:-1]
        dts = np.diff(np.asarray(sample_ts, np.int64)).astype(np.float64)
        seg[:-1, 19] = np.where(dts > 0, 1.0 / np.where(dts > 0, dts, 1.0), 0.0)
    # columns 20/21: t_k and dt_k = t_{k+1} - t_k as raw int64 bits (the row then verifies its own bracket)
    tsi = np.asarray(sample_ts, np.int64)
    seg[:, 20] = tsi.view(np.float64)
    dti = np.zeros(S, np.int64)
    dti[:-1] = np.diff(tsi)
    seg[:, 21] = dti.view(np.float64)
    return seg


def partition_frames(frame_off: np.ndarray, world_size: int) -> np.ndarray:
    """Contiguous frame ranges balanced by POINTS: returns int64[world_size+1] frame
    cut indices; rank r owns frames [cuts[r], cuts[r+1]) = points [off[cuts[r]], off[cuts[r+1]])."""
    frame_off = np.asarray(frame_off, np.int64)
    F, N = len(frame_off) - 1, int(frame_off[-1])
    targets = (np.arange(1, world_size, dtype=np.float64) * N / world_size)
    cuts = np.searchsorted(frame_off, targets, side='left')
    # choose the nearer of the two neighbouring frame boundaries
    for i, t in enumerate(targets):
        c = int(cuts[i])
        if c > 0 and abs(frame_off[c - 1] - t) <= abs(frame_off[min(c, F)] - t):
            cuts[i] = c - 1
    cuts = np.clip(cuts, 0, F)
    cuts = np.maximum.accumulate(cuts)
    return np.concatenate([[0], cuts, [F]]).astype(np.int64)
