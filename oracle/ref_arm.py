"""bench.py's reference arm -- TEST / BENCH INFRASTRUCTURE ONLY (see lmc_oracle.py header).

Times the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) through its own stock code path
for the hot path, on every host core of the box:

  stage A  per frame   pose_idx = min(np.searchsorted(traj_t, t), n-1)            LMC:804-806 (inline in run_simulation)
                       LiDARMotionSimulator.transform_pointcloud(points, {...})   LMC:772-776, called as LMC:826-832
           per worker  np.vstack(aligned)                                         LMC:888
  stage B  per frame   LivoxLVXWriter._write_frame -> _write_package ->           LMC:172-250
                       _write_point_data_type2 per point, into a BytesIO          LMC:252-272

The reference is single-threaded Python; "all the host threads it can use" = one forked worker process per
core, each running the stock functions on its own contiguous slice of the frames (frames are independent,
exactly the partition the GPU ranks use).  Nothing large crosses a process boundary: the sample is inherited
copy-on-write, every worker returns counts and a checksum.

The per-point Python LVX writer runs at ~0.04 Mpts/s per core, so stage B gets a smaller frame sample than
stage A; a step's rate is  1 / (T_A / n_A + T_B / n_B)  points/s (per-point costs add, both stages are linear
in the number of points), and both sample sizes are stated in `sample`.
"""
from __future__ import annotations

import io
import multiprocessing as mp
import os
import time

import numpy as np

from . import cpu_baseline as cb
from . import make_ref

_G = {}


def _init():
    mods = make_ref.load()
    _G["LMC"] = mods[0]
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        _G["sim"] = mods[0].LiDARMotionSimulator({})
    _G["writer"] = mods[0].LivoxLVXWriter()


def _stage_a(args):
    a, b = args
    s, sim = _G["sample"], _G["sim"]
    traj_t, pos, eul, frame_t, frames = s["traj_t"], s["pos"], s["eul"], s["frame_t"], s["frames"]
    aligned = []
    for i in range(a, b):
        pose_idx = np.searchsorted(traj_t, frame_t[i])                        # LMC:804-806
        pose_idx = max(min(pose_idx, len(traj_t) - 1), 0)
        aligned.append(sim.transform_pointcloud(frames[i], {'translation': pos[pose_idx], 'rotation': eul[pose_idx]}))
    merged = np.vstack(aligned) if aligned else np.zeros((0, 4))              # LMC:888
    return merged.shape[0], float(merged[:, :3].sum())


def _stage_b(args):
    a, b = args
    s, w = _G["sample"], _G["writer"]
    f = io.BytesIO()
    n = 0
    positions = [0] * (b - a + 1)
    for j, i in enumerate(range(a, b)):
        w._write_frame(f, {'frame_id': i, 'timestamp': float(s["frame_t"][i]), 'points': s["frames"][i]}, positions, j)
        n += len(s["frames"][i])
    return n, f.tell()


class ReferenceArm:
    def __init__(self, frames_a: int, ppf: int, frames_b: int = None, workers: int = None):
        if not make_ref.available():
            raise RuntimeError("oracle/_ref is not staged (python oracle/make_ref.py in the build container)")
        self.workers = workers or cb.default_threads()
        self.frames_a, self.ppf = frames_a, ppf
        self.frames_b = frames_b if frames_b is not None else self.workers     # one frame per worker and step
        _G["sample"] = cb.make_sample(max(frames_a, self.frames_b), ppf)
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.workers, initializer=_init)

    def _cuts(self, n):
        W = self.workers
        return [(j * n // W, (j + 1) * n // W) for j in range(W) if (j + 1) * n // W > j * n // W]

    def step(self):
        t0 = time.perf_counter()
        ra = self.pool.map(_stage_a, self._cuts(self.frames_a))
        t1 = time.perf_counter()
        rb = self.pool.map(_stage_b, self._cuts(self.frames_b))
        t2 = time.perf_counter()
        n_a, n_b = sum(r[0] for r in ra), sum(r[0] for r in rb)
        assert n_a == self.frames_a * self.ppf and n_b == self.frames_b * self.ppf
        sec_per_pt = (t1 - t0) / n_a + (t2 - t1) / n_b
        return dict(seconds=t2 - t0, t_a=t1 - t0, t_b=t2 - t1, n_a=n_a, n_b=n_b, points_per_s=1.0 / sec_per_pt)

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self) -> str:
        return (f"UNMODIFIED reference (oracle/_ref) on {self.workers} forked workers: per step stage A = {self.frames_a} frames x {self.ppf} pts "
                f"through LiDARMotionSimulator.transform_pointcloud per frame + np.vstack (LMC:802-832, 888), stage B = {self.frames_b} frames x "
                f"{self.ppf} pts through LivoxLVXWriter._write_frame / _write_point_data_type2 (LMC:172-272); rate = 1 / (T_A/n_A + T_B/n_B)")


def run(steps: int, warmup: int, frames_a: int, ppf: int):
    arm = ReferenceArm(frames_a, ppf)
    try:
        for _ in range(warmup):
            arm.step()
        rs = [arm.step() for _ in range(steps)]
    finally:
        arm.close()
    sec = float(np.mean([r["seconds"] for r in rs]))
    spp = float(np.mean([1.0 / r["points_per_s"] for r in rs]))
    return dict(value=1.0 / spp, seconds_per_step=sec, cores=arm.workers, sample=arm.describe(),
                stage_a_points_per_s=float(np.mean([r["n_a"] / r["t_a"] for r in rs])),
                stage_b_points_per_s=float(np.mean([r["n_b"] / r["t_b"] for r in rs])))
