"""CPU baseline for bench.py -- TEST/BENCH INFRASTRUCTURE ONLY (see lmc_oracle.py header).

What is timed is the reference's own CPU path for the hot path, restated with the same library
calls ("port": the Python reference itself cannot travel to the GPU box):

  per frame   pose_idx = min(np.searchsorted(traj_t, t), n-1)                     LMC:804-806
              R = scipy Rotation.from_euler('xyz', rpy).as_matrix()               LMC:774
              (R @ pts[:, :3].T).T + t ; np.column_stack([.., pts[:, 3]])         LMC:775-776
  once        np.vstack(aligned)                                                  LMC:888
  once        LVX int32-mm records of the raw points                              LMC:257-267
              (vectorised NumPy; the reference's per-point struct packing runs at 0.04 Mpts/s)

on float64 (n,4) arrays exactly like the reference.  Frames are spread over a thread pool
(NumPy/OpenBLAS release the GIL inside matmul and the big copies), so `cores` = threads used.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import lmc_oracle as orc


def make_sample(n_frames: int, pts_per_frame: int, seed: int = 4242):
    """Bounded sample of the M-1H stream: same sensor model and trajectory as synth.py."""
    rng = np.random.default_rng(seed)
    n = n_frames * pts_per_frame
    az = np.radians(rng.uniform(-35.2, 35.2, n)); el = np.radians(rng.uniform(-38.6, 38.6, n))
    r = rng.uniform(0.05, 90.0, n)
    pts = np.column_stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el), rng.uniform(0, 1, n)])
    pts = pts.astype(np.float32).astype(np.float64)
    frames = [pts[i * pts_per_frame:(i + 1) * pts_per_frame] for i in range(n_frames)]
    duration = n_frames / 10.0
    n_t = max(int(duration * 5), 2)
    traj_t = np.linspace(0, duration, n_t)
    pos = np.column_stack([30 * np.sin(2 * np.pi * 0.05 * traj_t), 30 * np.sin(4 * np.pi * 0.05 * traj_t), np.full(n_t, 1.5)])
    pos += rng.normal(0, 0.03, pos.shape)
    eul = np.column_stack([0.05 * np.sin(traj_t), np.zeros(n_t), np.unwrap(np.arctan2(np.gradient(pos[:, 1]), np.gradient(pos[:, 0])))])
    eul += rng.normal(0, 0.01, eul.shape)
    frame_t = np.arange(n_frames) / 10.0
    return dict(frames=frames, traj_t=traj_t, pos=pos, eul=eul, frame_t=frame_t, n_points=n)


def run_port(sample, threads: int) -> float:
    """One pass of the reference-equivalent NumPy path over the sample; returns seconds."""
    frames, traj_t, pos, eul, frame_t = (sample[k] for k in ('frames', 'traj_t', 'pos', 'eul', 'frame_t'))
    F = len(frames)
    t0 = time.perf_counter()
    idx = np.minimum(np.searchsorted(traj_t, frame_t), len(traj_t) - 1)

    def block(j):
        a, b = j * F // threads, (j + 1) * F // threads
        return [orc.transform_pointcloud_np(frames[i], {'translation': pos[idx[i]], 'rotation': eul[idx[i]]})
                for i in range(a, b)]
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(block, range(threads)))
        aligned = [x for p in parts for x in p]
    else:
        aligned = block(0)
    merged = np.vstack(aligned)

    def qblock(j):
        a, b = j * F // threads, (j + 1) * F // threads
        return orc.quantize_lvx_type2_np(np.vstack(frames[a:b])) if b > a else np.zeros((0, 14), np.uint8)
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            rec = np.vstack(list(ex.map(qblock, range(threads))))
    else:
        rec = qblock(0)
    dt = time.perf_counter() - t0
    assert merged.shape[0] == rec.shape[0] == sample['n_points']
    return dt


def run_c_port_mode_c(n_frames: int = 60, pts_per_frame: int = 10_000):
    """Extra context line: the single-thread C restatement of the north_star kernel (Mode C + LVX)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(1)
    n = n_frames * pts_per_frame
    pts = np.column_stack([rng.uniform(-90, 90, (n, 3)), rng.uniform(0, 1, n)])
    S = n_frames * 20 + 1
    sample_ts = np.arange(S, dtype=np.int64) * 5_000_000
    quat = Rotation.from_euler('xyz', np.cumsum(rng.normal(0, 0.01, (S, 3)), axis=0)).as_quat()
    seg = orc.slerp_segment_table(quat, np.cumsum(rng.normal(0, 0.05, (S, 3)), axis=0), sample_ts)
    off = np.arange(n_frames + 1, dtype=np.int64) * pts_per_frame
    ts = np.repeat(np.arange(n_frames, dtype=np.int64) * 100_000_000, pts_per_frame) + np.tile(np.arange(pts_per_frame, dtype=np.int64) * 10_000, n_frames)
    orc.C.deskew_slerp_f64(pts[:1000], ts[:1000], np.array([0, 1000]), sample_ts, seg)
    t0 = time.perf_counter()
    orc.C.deskew_slerp_f64(pts, ts, off, sample_ts, seg)
    orc.C.quantize_lvx_type2(pts)
    return n / (time.perf_counter() - t0)


def run_writers_port(n: int = 200_000):
    """Context for the device writers: the reference's own per-point Python loops restated as the fastest
    single-thread NumPy/Python equivalents ('%.6f' formatting per row; LVX records + container layout)."""
    rng = np.random.default_rng(2)
    pts = np.column_stack([rng.uniform(-90, 90, (n, 3)), rng.uniform(0, 1, n)])
    t0 = time.perf_counter()
    body = "".join("%.6f %.6f %.6f %.6f\n" % tuple(r) for r in pts)
    t_pcd = time.perf_counter() - t0
    assert len(body) > 0
    t0 = time.perf_counter()
    rec = orc.quantize_lvx_type2_np(pts)
    t_lvx = time.perf_counter() - t0
    return {"pcd_ascii": n / t_pcd, "lvx_records": n / t_lvx}


def default_threads() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))
