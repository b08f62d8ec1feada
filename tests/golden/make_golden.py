#!/usr/bin/env python
"""
Generate the golden fixtures in tests/golden/ by running the REAL reference.

Runs only in the build container (needs /root/reference, read-only).  The reference is
pure Python; it is imported with stub modules for its missing third-party imports
(laspy for LMC; matplotlib / mpl_toolkits for CS) -- no reference code is copied, the
fixtures hold only inputs and the reference's outputs.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz + MANIFEST.json

LMC = /root/reference/lidar_motion_compensation.py
CS  = /root/reference/livox_mid70_complete_simulator.py
"""
import contextlib
import hashlib
import io
import json
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LMC_REFERENCE_DIR", "/root/reference")


def import_reference():
    """Import both reference modules with stubs for absent third-party packages."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:                                     # the real laspy, if a wheel ever appears: LAS parity then pins itself (--only las)
        import laspy                         # noqa: F401
    except ImportError:
        sys.modules.setdefault('laspy', types.ModuleType('laspy'))
    for m in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.animation', 'mpl_toolkits',
              'mpl_toolkits.mplot3d']:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib.animation'].FuncAnimation = object
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    logging.disable(logging.CRITICAL)
    with contextlib.redirect_stdout(io.StringIO()):
        import lidar_motion_compensation as LMC
        import livox_mid70_complete_simulator as CS
    return LMC, CS


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# the five anchor configs of SURVEY.md section 4
CONFIGS = {
    'C1a': dict(trajectory_type='linear', environment_complexity='simple', duration=60.0, max_speed=25.0, lidar_fps=10),
    'C1b': dict(trajectory_type='linear', environment_complexity='simple', duration=60.0, max_speed=25.0, lidar_fps=15),
    'C2a': dict(trajectory_type='figure_eight', environment_complexity='complex', duration=60.0, max_speed=12.0, lidar_fps=10),
    'C2b': dict(trajectory_type='figure_eight', environment_complexity='complex', duration=120.0, max_speed=12.0, lidar_fps=10),
    'C3':  dict(trajectory_type='circular', environment_complexity='medium', duration=30.0, max_speed=5.0, lidar_fps=20),
}
# which frames are stored in full (None = all)
KEEP = {'C1a': None, 'C2a': 10, 'C3': 8}


def run_lmc(LMC, cfg):
    sim = LMC.LiDARMotionSimulator(dict(cfg))
    with contextlib.redirect_stdout(io.StringIO()):
        res = sim.run_simulation()
    return sim, res


def lmc_fixture(LMC, name, manifest):
    sim, res = run_lmc(LMC, CONFIGS[name])
    raws = [s['points_local'] for s in res['raw_scans']]
    al = res['aligned_pointclouds']
    counts = np.array([len(r) for r in raws], np.int64)
    raw_all, al_all = np.vstack(raws), np.vstack(al)
    traj = res['trajectory']
    frame_t = np.array([s['timestamp'] for s in res['raw_scans']], np.float64)
    # the reference's own pose index per frame (LMC:804-806), recovered from the pose it stored
    pos_used = np.array([s['sensor_pose']['position'] for s in res['raw_scans']])
    eul_used = np.array([s['sensor_pose']['orientation'] for s in res['raw_scans']])
    pose_idx = np.array([int(np.flatnonzero((traj['position_gps'] == p).all(1))[0]) for p in pos_used], np.int32)
    entry = dict(config=CONFIGS[name], frames=len(raws), traj_samples=len(traj['time']),
                 total_points=int(counts.sum()), empty_frames=int((counts == 0).sum()),
                 single_point_frames=int((counts == 1).sum()),
                 max_points_per_frame=int(counts.max()), raw_sha256=sha(raw_all), aligned_sha256=sha(al_all))
    manifest['lmc'][name] = entry
    if name not in KEEP:
        return
    keep = KEEP[name]
    if keep is None:
        ids = np.arange(len(raws))
    else:
        ids = np.unique(np.linspace(0, len(raws) - 1, keep).astype(int))
    sel_counts = counts[ids]
    off = np.zeros(len(ids) + 1, np.int64)
    np.cumsum(sel_counts, out=off[1:])
    np.savez_compressed(
        os.path.join(HERE, f'lmc_{name}.npz'),
        frame_ids=ids.astype(np.int32), frame_off=off,
        raw=np.vstack([raws[i] for i in ids]), aligned=np.vstack([al[i] for i in ids]),
        pose_position=pos_used[ids], pose_euler=eul_used[ids],
        traj_time=traj['time'], traj_position_gps=traj['position_gps'],
        traj_orientation_imu=traj['orientation_imu'],
        frame_t_all=frame_t, pose_idx_all=pose_idx, counts_all=counts)
    entry['fixture'] = f'lmc_{name}.npz'
    entry['fixture_frames'] = int(len(ids))


def lmc_edge_fixture(LMC, manifest):
    """Reference transform_pointcloud (LMC:772-776) on ragged / tiny frames: n = 0, 1 (NumPy
    routes a 3x1 right-hand side through gemv: different FMA order), 2, 3, ... and large
    translations."""
    rng = np.random.default_rng(5)
    sim = LMC.LiDARMotionSimulator()
    counts = [0, 1, 2, 1, 3, 0, 0, 1, 17, 64, 1, 1000, 1, 0]
    raws, outs, pos, eul = [], [], [], []
    for n in counts:
        p = np.column_stack([rng.uniform(-90, 90, (n, 3)), rng.uniform(0.1, 0.9, n)]).reshape(n, 4)
        t = rng.uniform(-1500, 1500, 3); e = rng.normal(0, 0.7, 3)
        raws.append(p); pos.append(t); eul.append(e)
        outs.append(sim.transform_pointcloud(p, {'translation': t, 'rotation': e}))
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    np.savez_compressed(os.path.join(HERE, 'lmc_edge.npz'), raw=np.vstack(raws), aligned=np.vstack(outs),
                        frame_off=off, pose_position=np.array(pos), pose_euler=np.array(eul))
    manifest['lmc_edge'] = dict(frames=len(counts), points=int(off[-1]), aligned_sha256=sha(np.vstack(outs)))


def lvx_type2_fixture(LMC, manifest):
    """Reference per-point packer LMC:252-272 on adversarial + random rows."""
    rng = np.random.default_rng(20261018)
    rows = [rng.uniform(-90, 90, (4000, 3))]
    rows.append(rng.uniform(-1, 1, (500, 3)) * 1e-3)                     # sub-millimetre
    k = rng.integers(-90000, 90000, (1500, 3))
    rows.append(k / 1000.0)                                              # exact-mm decimals (x*1000 may land 1 ulp under)
    rows.append((k[:500] + 0.5) / 1000.0)
    rows.append(rng.uniform(-4e6, 4e6, (300, 3)))                        # clip to int32 range
    rows.append(np.array([[2147483.647, -2147483.648, 0.0], [2147483.648, -2147483.649, -0.0],
                          [np.inf, -np.inf, 1e300], [0.0009999999, -0.0009999999, 5e-324]]))
    xyz = np.vstack(rows)
    inten = rng.uniform(0, 1, len(xyz))
    inten[:50] = rng.uniform(-0.5, 1.5, 50)                              # clip of reflectivity
    inten[50:60] = [0.0, 1.0, 0.999999999, 1 / 255, 2 / 255, 254.9999 / 255, 0.5, 0.25, 1e-9, 0.00392156862745098]
    pts = np.column_stack([xyz, inten])
    w = LMC.LivoxLVXWriter()
    buf = io.BytesIO()
    for p in pts:
        w._write_point_data_type2(buf, p)
    out = np.frombuffer(buf.getvalue(), np.uint8).reshape(-1, 14)
    np.savez_compressed(os.path.join(HERE, 'lvx_type2.npz'), pts=pts, records=out)
    manifest['lvx_type2'] = dict(points=len(pts), records_sha256=sha(out))


def lvx_file_fixture(LMC, manifest):
    """Whole-file bytes of LivoxLVXWriter.write_compatible_lvx (LMC:58-250) on tiny frames."""
    rng = np.random.default_rng(7)
    counts = [0, 1, 95, 96, 97, 200, 0, 3]
    frames, raw = [], []
    for i, n in enumerate(counts):
        p = np.column_stack([rng.uniform(-50, 50, (n, 3)), rng.uniform(0, 1, n)]).reshape(n, 4)
        frames.append({'frame_id': i, 'timestamp': i * 0.1, 'points': p})
        raw.append(p)
    path = os.path.join(HERE, '_tmp.lvx')
    with contextlib.redirect_stdout(io.StringIO()):
        ok = LMC.LivoxLVXWriter().write_compatible_lvx(path, frames)
    assert ok
    data = np.fromfile(path, np.uint8)
    os.remove(path)
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    np.savez_compressed(os.path.join(HERE, 'lvx_file.npz'), raw=np.vstack(raw), frame_off=off,
                        timestamps=np.arange(len(counts)) * 0.1, file_bytes=data)
    manifest['lvx_file'] = dict(frames=len(counts), bytes=int(len(data)), sha256=sha(data))


def coord_chain_fixture(CS, manifest):
    """(N3) Reference CoordinateTransformer (CS:153-233): sensor->vehicle default, a user-set
    vehicle->local transform and their inverse, applied with transform_points (4x4 homogeneous)."""
    rng = np.random.default_rng(314)
    ct = CS.CoordinateTransformer()
    ct.set_transformation(CS.CoordinateSystem.VEHICLE, CS.CoordinateSystem.LOCAL, [12.5, -3.25, 0.75], [0.02, -0.03, 1.1])
    cases = [(CS.CoordinateSystem.SENSOR, CS.CoordinateSystem.VEHICLE), (CS.CoordinateSystem.VEHICLE, CS.CoordinateSystem.LOCAL),
             (CS.CoordinateSystem.LOCAL, CS.CoordinateSystem.VEHICLE)]
    pts, outs, mats, counts = [], [], [], []
    for (a, b), n in zip(cases, [1500, 2, 777]):
        p = rng.uniform(-90, 90, (n, 3))
        o = ct.transform_points(p, a, b)
        pts.append(p); outs.append(o); mats.append(ct.transformations[(a, b)]); counts.append(n)
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    np.savez_compressed(os.path.join(HERE, 'coord_chain.npz'), pts=np.vstack(pts), out=np.vstack(outs),
                        T=np.array(mats), frame_off=off)
    manifest['coord_chain'] = dict(frames=len(counts), points=int(off[-1]), out_sha256=sha(np.vstack(outs)))


def pcd_fixture(LMC, manifest):
    """(N2) Reference save_pcd (LMC:932-948) on adversarial values: exact ties (odd k / 128), carries,
    negative zero, tiny / denormal, large magnitudes, nan / inf."""
    rng = np.random.default_rng(2718)
    vals = [rng.uniform(-90, 90, 4000), rng.uniform(-1, 1, 500) * 1e-6, np.arange(1, 400, 2) / 128.0, -np.arange(1, 400, 2) / 128.0,
            np.array([0.0, -0.0, 0.9999995, 0.99999949999, 9.9999995, 99.9999995, -0.9999995, 1e-7, -1e-9, 5e-324, 2.5e-7, 7.5e-7,
                      123456.7890125, -99999.9999995, 1e12, -8.5e12, 9.1e12, 0.5, 1.5e-6, 2.5e-6, 3.5e-6, 1048575.9999995,
                      np.nan, np.inf, -np.inf, 1.0, -1.0, 1e6, 0.1, 0.2, 0.3]),
            rng.uniform(-2000, 2000, 3000), rng.uniform(0, 1, 2000)]
    v = np.concatenate(vals)
    v = np.concatenate([v, np.zeros((-len(v)) % 4)])
    pts = v.reshape(-1, 4)
    path = os.path.join(HERE, '_tmp.pcd')
    LMC.LiDARMotionSimulator().save_pcd(pts, path)
    data = np.fromfile(path, np.uint8)
    os.remove(path)
    np.savez_compressed(os.path.join(HERE, 'pcd_ascii.npz'), pts=pts, file_bytes=data)
    manifest['pcd_ascii'] = dict(points=len(pts), bytes=int(len(data)), sha256=sha(data))


def scan_fixture(LMC, name, manifest):
    """(N4) inputs of the reference's frame loop for config `name`: environment, noisy trajectory, the
    frame grid, the effective config and the GLOBAL NumPy RNG state right before the loop (LMC:802), so
    the scanner + its noise stream can be replayed without the reference.  Expected output = the raw
    sha256 anchor of the same config in MANIFEST['lmc']."""
    sim = LMC.LiDARMotionSimulator(dict(CONFIGS[name]))
    traj = sim.add_sensor_noise(sim.generate_trajectory())
    with contextlib.redirect_stdout(io.StringIO()):
        env = sim.generate_environment_pointcloud()
    state = np.random.get_state()
    # replay the loop with the reference's own scan_environment from this state: must hit the anchor
    lidar_times = np.linspace(0, sim.config['duration'], int(sim.config['duration'] * sim.config['lidar_fps']))
    raws = []
    for t in lidar_times:
        k = max(min(int(np.searchsorted(traj['time'], t)), len(traj['time']) - 1), 0)
        raws.append(sim.scan_environment(env, {'position': traj['position_gps'][k], 'orientation': traj['orientation_imu'][k]}))
    assert sha(np.vstack(raws)) == manifest['lmc'][name]['raw_sha256'], name
    cfg = {k: sim.config[k] for k in ['range_max', 'range_min', 'fov_horizontal', 'fov_vertical', 'points_per_frame',
                                      'lidar_range_noise', 'duration', 'lidar_fps']}
    np.savez_compressed(os.path.join(HERE, f'scan_{name}.npz'), environment=env, traj_time=traj['time'],
                        traj_position_gps=traj['position_gps'], traj_orientation_imu=traj['orientation_imu'],
                        rng_keys=state[1], rng_pos=np.int64(state[2]), rng_has_gauss=np.int64(state[3]),
                        rng_cached=np.float64(state[4]), config_json=np.frombuffer(json.dumps(cfg).encode(), np.uint8))
    manifest.setdefault('scan', {})[name] = dict(env_points=int(len(env)), fixture=f'scan_{name}.npz')


def outputs_fixture(LMC, name, manifest):
    """The reference's whole output directory for config `name` (run_simulation + save_results, LMC:860-930): sha256 and
    size of every hot-path file -- raw_scans_pcd/frame_%04d.pcd, aligned_scans_pcd/aligned_frame_%04d.pcd,
    merged_aligned.pcd (absent when a frame is empty, LMC:887-891), merged_raw_overlapped.pcd, lidar_data.lvx.
    (merged_aligned.las needs laspy; the two CSVs carry generator columns the scan fixtures do not store.)"""
    import tempfile
    import shutil
    sim, res = run_lmc(LMC, CONFIGS[name])
    tmp = tempfile.mkdtemp(prefix='lmc_ref_out_')
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim.save_results(res, tmp)
        names, digests, sizes = [], [], []
        for dp, _, fs in os.walk(tmp):
            for f in fs:
                rel = os.path.relpath(os.path.join(dp, f), tmp).replace(os.sep, '/')
                if not (rel.endswith('.pcd') or rel.endswith('.lvx')):
                    continue
                data = open(os.path.join(dp, f), 'rb').read()
                names.append(rel); digests.append(hashlib.sha256(data).hexdigest()); sizes.append(len(data))
        order = np.argsort(names)
        np.savez_compressed(os.path.join(HERE, f'outputs_{name}.npz'), names=np.array(names)[order],
                            sha256=np.array(digests)[order], sizes=np.array(sizes, np.int64)[order])
        manifest.setdefault('outputs', {})[name] = dict(files=len(names), bytes=int(np.sum(sizes)), fixture=f'outputs_{name}.npz',
                                                        merged_aligned='merged_aligned.pcd' in names)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def modeb_fixture(CS, manifest):
    """Reference MotionCompensator.compensate_point_cloud (CS:1435-1536) + LVX2 packer
    (CS:365-374) on synthetic Mid-70-shaped frames against the reference's own 200 Hz
    IMUSimulator output (CS:1191-1246)."""
    rng = np.random.default_rng(99)
    cfg = {'random_seed': 42, 'duration': 2.0, 'trajectory_type': 'figure_eight', 'max_speed': 12.0}
    tg = CS.TrajectoryGenerator('figure_eight', seed=42)
    traj = tg.generate_trajectory(2.0, dt=0.1, max_speed=12.0)
    imu = CS.IMUSimulator(cfg).simulate_imu_data(traj, 2.0)
    imu_ts = np.array([s.timestamp for s in imu], np.int64)
    # make the motion non-trivial: the simulated gyro of a 2 s figure-eight is small, so add a
    # deterministic swell (still goes through the reference code unchanged)
    swell = 0.8 * np.sin(np.arange(len(imu))[:, None] * np.array([0.05, 0.031, 0.07]))
    for s, d in zip(imu, swell):
        s.gyro_x += d[0]; s.gyro_y += d[1]; s.gyro_z += d[2]
    imu_gyro = np.array([[s.gyro_x, s.gyro_y, s.gyro_z] for s in imu], np.float64)
    mc = CS.MotionCompensator({'enable_motion_compensation': True})
    # frames: one before the first IMU sample (clamp), three inside, one past the end (clamp)
    frame_starts = [-50_000_000, 100_000_000, 700_000_000, 1_234_567_891, 1_990_000_000]
    counts = [300, 700, 701, 1, 400]
    all_pts, all_ts, all_out, all_tag = [], [], [], []
    lvx2 = io.BytesIO()
    w = CS.LivoxLVXWriter('lvx2')
    for fs, n in zip(frame_starts, counts):
        az = np.radians(rng.uniform(-35.2, 35.2, n)); el = np.radians(rng.uniform(-38.6, 38.6, n))
        r = rng.uniform(0.05, 90, n)
        xyz = np.column_stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)])
        inten = rng.integers(0, 256, n)
        step = 100_000_000 // max(n, 1)
        ts = fs + np.arange(n, dtype=np.int64) * step + rng.integers(0, 7, n)
        tag = rng.integers(0, 2, n)
        pts = [CS.LiDARPoint(x=float(xyz[i, 0]), y=float(xyz[i, 1]), z=float(xyz[i, 2]), intensity=int(inten[i]),
                             timestamp=int(ts[i]), ring=i % 16, tag=int(tag[i])) for i in range(n)]
        comp = mc.compensate_point_cloud(pts, imu, fs, 100_000_000)
        out = np.array([[p.x, p.y, p.z, p.intensity] for p in comp], np.float64).reshape(n, 4)
        all_pts.append(np.column_stack([xyz, inten.astype(np.float64)])); all_ts.append(ts)
        all_out.append(out); all_tag.append(tag.astype(np.uint8))
        fbuf = io.BytesIO()
        w._write_frame_lvx2(fbuf, {'points': comp, 'timestamp': fs if fs > 0 else 0}, 0)
        lvx2.write(fbuf.getvalue()[24 + 21:])       # skip 24-B frame header + 21-B package header
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    rec = np.frombuffer(lvx2.getvalue(), np.uint8).reshape(-1, 14)
    np.savez_compressed(os.path.join(HERE, 'modeb.npz'), pts=np.vstack(all_pts), ts=np.concatenate(all_ts),
                        tag=np.concatenate(all_tag), frame_off=off, frame_start=np.array(frame_starts, np.int64),
                        imu_ts=imu_ts, imu_gyro=imu_gyro, compensated=np.vstack(all_out), lvx2_records=rec)
    manifest['modeb'] = dict(points=int(off[-1]), imu_samples=len(imu), compensated_sha256=sha(np.vstack(all_out)),
                             lvx2_sha256=sha(rec))


def modeb_tiers_fixture(CS, manifest):
    """The same reference path (CS:1435-1536 + 365-374) at gyro rates up to ~12 rad/s: per-point angles from 0 to ~1.2 rad,
    so every sin / cos tier of the device code (|a| < 2^-4, <= 2^-3, <= 0.5, library) meets the reference's np.sin / np.cos."""
    rng = np.random.default_rng(1234)
    cfg = {'random_seed': 42, 'duration': 1.0, 'trajectory_type': 'figure_eight', 'max_speed': 12.0}
    tg = CS.TrajectoryGenerator('figure_eight', seed=42)
    traj = tg.generate_trajectory(1.0, dt=0.1, max_speed=12.0)
    imu = CS.IMUSimulator(cfg).simulate_imu_data(traj, 1.0)
    imu_ts = np.array([s.timestamp for s in imu], np.int64)
    amp = np.linspace(0.05, 12.0, len(imu))[:, None]
    swell = amp * np.sin(np.arange(len(imu))[:, None] * np.array([0.11, 0.071, 0.13]) + np.array([0.3, 1.1, 2.0]))
    for s, d in zip(imu, swell):
        s.gyro_x += d[0]; s.gyro_y += d[1]; s.gyro_z += d[2]
    imu_gyro = np.array([[s.gyro_x, s.gyro_y, s.gyro_z] for s in imu], np.float64)
    mc = CS.MotionCompensator({'enable_motion_compensation': True})
    frame_starts = [0, 100_000_000, 300_000_000, 500_000_000, 700_000_000, 880_000_000]
    counts = [400, 401, 399, 400, 402, 400]
    all_pts, all_ts, all_out, all_tag = [], [], [], []
    lvx2 = io.BytesIO()
    w = CS.LivoxLVXWriter('lvx2')
    for fs, n in zip(frame_starts, counts):
        az = np.radians(rng.uniform(-35.2, 35.2, n)); el = np.radians(rng.uniform(-38.6, 38.6, n))
        r = rng.uniform(0.05, 90, n)
        xyz = np.column_stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)])
        inten = rng.integers(0, 256, n)
        ts = fs + np.arange(n, dtype=np.int64) * (100_000_000 // n) + rng.integers(0, 7, n)
        tag = rng.integers(0, 2, n)
        pts = [CS.LiDARPoint(x=float(xyz[i, 0]), y=float(xyz[i, 1]), z=float(xyz[i, 2]), intensity=int(inten[i]),
                             timestamp=int(ts[i]), ring=i % 16, tag=int(tag[i])) for i in range(n)]
        comp = mc.compensate_point_cloud(pts, imu, fs, 100_000_000)
        out = np.array([[p.x, p.y, p.z, p.intensity] for p in comp], np.float64).reshape(n, 4)
        all_pts.append(np.column_stack([xyz, inten.astype(np.float64)])); all_ts.append(ts)
        all_out.append(out); all_tag.append(tag.astype(np.uint8))
        fbuf = io.BytesIO()
        w._write_frame_lvx2(fbuf, {'points': comp, 'timestamp': fs}, 0)
        lvx2.write(fbuf.getvalue()[24 + 21:])       # skip 24-B frame header + 21-B package header
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    rec = np.frombuffer(lvx2.getvalue(), np.uint8).reshape(-1, 14)
    # per-point angle magnitudes (what picks the tier), for the test's coverage assertion
    tsa, fsa = np.concatenate(all_ts), np.repeat(np.array(frame_starts, np.int64), counts)
    g = np.column_stack([np.interp(tsa, imu_ts, imu_gyro[:, c]) for c in range(3)])
    amax = np.abs(g * ((tsa - fsa) * 1e-9)[:, None]).max(axis=1)
    np.savez_compressed(os.path.join(HERE, 'modeb_tiers.npz'), pts=np.vstack(all_pts), ts=tsa,
                        tag=np.concatenate(all_tag), frame_off=off, frame_start=np.array(frame_starts, np.int64),
                        imu_ts=imu_ts, imu_gyro=imu_gyro, compensated=np.vstack(all_out), lvx2_records=rec, angle_max=amax)
    manifest['modeb_tiers'] = dict(points=int(off[-1]), imu_samples=len(imu), compensated_sha256=sha(np.vstack(all_out)),
                                   lvx2_sha256=sha(rec), tier_counts=[int((amax < 0.0625).sum()), int(((amax >= 0.0625) & (amax <= 0.125)).sum()),
                                                                      int(((amax > 0.125) & (amax <= 0.5)).sum()), int((amax > 0.5).sum())])


def modec_lerp_fixture(CS, manifest):
    """The reference has no per-point pose interpolation as a whole (Mode C is the north_star's), but it does interpolate a
    trajectory to arbitrary times: IMUSimulator._interpolate_trajectory (CS:1248-1275, np.interp over trajectory['time']).  That pins
    the bracket search and the POSITION lerp of Mode C: with identity orientations Mode C's output is p + lerp(position)(t)."""
    rng = np.random.default_rng(2024)
    tg = CS.TrajectoryGenerator('figure_eight', seed=7)
    traj = tg.generate_trajectory(3.0, dt=0.005, max_speed=15.0)                 # 200 Hz samples
    t_s = np.asarray(traj['time'], np.float64)
    sample_ts = np.round(t_s * 1e9).astype(np.int64)
    tt = sample_ts.astype(np.float64) * 1e-9                                      # the times both sides interpolate over
    traj = dict(traj); traj['time'] = tt
    n = 6000
    ts = np.sort(rng.integers(sample_ts[0] - 30_000_000, sample_ts[-1] + 30_000_000, n)).astype(np.int64)
    ts[:40] = sample_ts[rng.integers(0, len(sample_ts), 40)]                     # exactly on samples
    ts = np.sort(ts)
    imu = CS.IMUSimulator({'random_seed': 1})
    interp = imu._interpolate_trajectory(traj, ts.astype(np.float64) * 1e-9)
    pts = np.column_stack([rng.uniform(-50, 50, (n, 3)), rng.uniform(0, 1, n)])
    pos = np.column_stack([traj['x'], traj['y'], traj['z']]).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, 'modec_lerp.npz'), pts=pts, ts=ts, sample_ts=sample_ts, sample_pos=pos,
                        ref_position=interp['position'], ref_orientation=interp['orientation'],
                        sample_euler=np.column_stack([traj['roll'], traj['pitch'], traj['yaw']]).astype(np.float64))
    manifest['modec_lerp'] = dict(points=n, samples=len(sample_ts), position_sha256=sha(interp['position']))


def lvx_cs_fixture(CS, manifest):
    """Reference LivoxLVXWriter.write_lvx_file (CS:245-374) for 'lvx2', 'lvx3' and 'lvx' on ragged synthetic
    frames (one empty, one of a single point, one above 1024 points) with a non-trivial DeviceInfo."""
    import tempfile
    rng = np.random.default_rng(2024)
    counts = [5, 0, 1, 1500, 96, 1024, 1025, 0]
    stamps = [0, 100_000_000, 200_000_003, 1_700_000_000_123_456_789, 400_000_000, 500_000_000, 600_000_000, 700_000_000]
    di = CS.DeviceInfo(lidar_sn="3GGDJ6K00200101", device_type=1, firmware_version="03.08.0000", extrinsic_enable=True,
                       roll=0.01, pitch=-0.02, yaw=1.5, x=0.1, y=-0.2, z=1.25)
    frames, all_pts, all_tag = [], [], []
    for n, t in zip(counts, stamps):
        xyz = rng.uniform(-120, 120, (n, 3))
        if n >= 5:
            xyz[0] = [0.0, -0.0004, 0.0009999]; xyz[1] = [-0.0015, 2147483.647, -2147483.648]
            xyz[2] = [1e-320, -1e-320, 0.9999999999999999]
        inten = rng.integers(0, 256, n); tag = rng.integers(0, 256, n)
        frames.append({'points': [CS.LiDARPoint(x=float(a[0]), y=float(a[1]), z=float(a[2]), intensity=int(i), timestamp=int(t) + k,
                                                ring=k % 16, tag=int(g)) for k, (a, i, g) in enumerate(zip(xyz, inten, tag))],
                       'timestamp': int(t)})
        all_pts.append(np.column_stack([xyz, inten.astype(np.float64)]).reshape(n, 4)); all_tag.append(tag.astype(np.uint8))
    off = np.zeros(len(counts) + 1, np.int64); np.cumsum(counts, out=off[1:])
    files = {}
    with tempfile.TemporaryDirectory() as d:
        for ver in ['lvx2', 'lvx3', 'lvx']:
            fn = os.path.join(d, 'x.' + ver)
            CS.LivoxLVXWriter(ver).write_lvx_file(fn, frames, di)
            files[ver] = np.frombuffer(open(fn, 'rb').read(), np.uint8)
    assert np.array_equal(files['lvx2'], files['lvx3'])
    np.savez_compressed(os.path.join(HERE, 'lvx_cs.npz'), pts=np.vstack(all_pts), tag=np.concatenate(all_tag), frame_off=off,
                        frame_ts=np.array(stamps, np.int64), file_lvx2=files['lvx2'], file_legacy=files['lvx'],
                        device_info_json=np.frombuffer(json.dumps(di.__dict__).encode(), np.uint8))
    manifest['lvx_cs'] = dict(frames=len(counts), points=int(off[-1]), lvx2_bytes=int(len(files['lvx2'])), lvx2_sha256=sha(files['lvx2']),
                              legacy_bytes=int(len(files['lvx'])), legacy_sha256=sha(files['lvx']))


def text_rows_fixture(CS, manifest):
    """Reference DataExporter._export_pcd / _export_xyz / _export_csv (CS:1643-1716) on an (N,5) array
    [x y z intensity timestamp] with rounding ties, negative zeros, sub-ulp values and ns timestamps."""
    import tempfile
    rng = np.random.default_rng(31337)
    n = 1500
    a = np.column_stack([rng.uniform(-150, 150, (n, 3)), rng.integers(0, 256, n).astype(np.float64),
                         (rng.integers(0, 3_600_000_000_000, n)).astype(np.float64)])
    a[0] = [0.0, -0.0, 1e-7, 0.0, 0.0]
    a[1] = [0.0000005, -0.0000005, 0.0000015, 0.5, 1.5]                 # ties after binary rounding
    a[2] = [2.5e-7 * 2, 123456.7890125, -999999.9999995, 2.5, 1_700_000_000_123_456_789.0]
    a[3] = [1e-320, -1e-320, 4.9e-324, 254.5, 255.5]
    a[4] = [9.9999995, 99.9999995, -0.9999995, 255.0, 9_007_199_254_740_993.0]
    a[5] = [1234567.0000005, -7654321.1234565, 0.1234565, 17.49999999, 3_599_999_999_999.5]
    ex = CS.DataExporter({})
    with tempfile.TemporaryDirectory() as d:
        ex._export_pcd(a, os.path.join(d, 'a.pcd')); ex._export_xyz(a, os.path.join(d, 'a.xyz')); ex._export_csv(a, os.path.join(d, 'a.csv'))
        pcd = open(os.path.join(d, 'a.pcd'), 'rb').read(); xyz = open(os.path.join(d, 'a.xyz'), 'rb').read(); csv = open(os.path.join(d, 'a.csv'), 'rb').read()
    np.savez_compressed(os.path.join(HERE, 'text_rows.npz'), points=a, pcd=np.frombuffer(pcd, np.uint8), xyz=np.frombuffer(xyz, np.uint8),
                        csv=np.frombuffer(csv, np.uint8))
    manifest['text_rows'] = dict(points=n, pcd_sha256=sha(np.frombuffer(pcd, np.uint8)), xyz_sha256=sha(np.frombuffer(xyz, np.uint8)),
                                 csv_sha256=sha(np.frombuffer(csv, np.uint8)))


def coord_frames_fixture(CS, manifest):
    """(N3) Reference LiDARMotionSimulator._transform_coordinates (CS:2107-2163): every point goes through
    transform_points as a (1,3) array.  Called unbound with a stub self (the method only touches
    self.coordinate_transformer); targets 'vehicle' (default) and 'local' (user-set), ragged frames."""
    rng = np.random.default_rng(2718)
    ct = CS.CoordinateTransformer()
    ct.set_transformation(CS.CoordinateSystem.SENSOR, CS.CoordinateSystem.LOCAL, [105.25, -37.5, 2.125], [0.013, -0.021, 2.3])
    stub = types.SimpleNamespace(coordinate_transformer=ct)
    counts = [40, 0, 1, 300]
    frames, pts = [], []
    for i, n in enumerate(counts):
        xyz = rng.uniform(-90, 90, (n, 3))
        frames.append({'points': [CS.LiDARPoint(x=float(a[0]), y=float(a[1]), z=float(a[2]), intensity=k % 256, timestamp=i * 10 ** 8 + k,
                                                ring=k % 16, tag=k % 3) for k, a in enumerate(xyz)], 'timestamp': i * 10 ** 8})
        pts.append(xyz)
    out = {}
    for target in [CS.CoordinateSystem.VEHICLE, CS.CoordinateSystem.LOCAL]:
        res = CS.LiDARMotionSimulator._transform_coordinates(stub, frames, target, [])
        assert all(r['coordinate_system'] == target for r in res)
        out[target] = np.array([[p.x, p.y, p.z] for r in res for p in r['points']], np.float64)
    off = np.zeros(len(counts) + 1, np.int64); np.cumsum(counts, out=off[1:])
    np.savez_compressed(os.path.join(HERE, 'coord_frames.npz'), pts=np.vstack(pts), frame_off=off, out_vehicle=out['vehicle'], out_local=out['local'],
                        T_local=ct.transformations[(CS.CoordinateSystem.SENSOR, CS.CoordinateSystem.LOCAL)],
                        T_vehicle=ct.transformations[(CS.CoordinateSystem.SENSOR, CS.CoordinateSystem.VEHICLE)])
    manifest['coord_frames'] = dict(points=int(off[-1]), vehicle_sha256=sha(out['vehicle']), local_sha256=sha(out['local']))


def config2_fixture(CS, manifest):
    """BASELINE configs[2] -- parking_detailed frames, per-point timestamps, deskew against a 200 Hz IMU stream:
    raw C3 frames (already the LMC reference's output, lmc_C3.npz) subsampled to 600 points each, timestamps
    frame_ns + i * (frame_period / n_f), the reference's own IMUSimulator on a 30 s circular
    trajectory (6000 samples), pushed through the reference MotionCompensator.compensate_point_cloud."""
    g = np.load(os.path.join(HERE, 'lmc_C3.npz'))
    off = g['frame_off']
    cfg = {'random_seed': 42, 'duration': 30.0, 'trajectory_type': 'circular', 'max_speed': 5.0}
    tg = CS.TrajectoryGenerator('circular', seed=42)
    traj = tg.generate_trajectory(30.0, dt=0.1, max_speed=5.0)
    imu = CS.IMUSimulator(cfg).simulate_imu_data(traj, 30.0)
    imu_ts = np.array([s.timestamp for s in imu], np.int64)
    imu_gyro = np.array([[s.gyro_x, s.gyro_y, s.gyro_z] for s in imu], np.float64)
    mc = CS.MotionCompensator({'enable_motion_compensation': True})
    period = 50_000_000                                     # 20 fps
    pts_all, ts_all, out_all, starts, counts = [], [], [], [], []
    for i in [0, 3, 7]:
        raw = g['raw'][off[i]:off[i + 1]]
        raw = raw[::max(1, len(raw) // 600)][:600]
        n = len(raw)
        fs = int(float(g['frame_t_all'][g['frame_ids'][i]]) * 1e9)
        ts = fs + np.arange(n, dtype=np.int64) * (period // n)
        pts = [CS.LiDARPoint(x=float(p[0]), y=float(p[1]), z=float(p[2]), intensity=float(p[3]), timestamp=int(t), ring=k % 16, tag=0)
               for k, (p, t) in enumerate(zip(raw, ts))]
        comp = mc.compensate_point_cloud(pts, imu, fs, period)
        pts_all.append(raw); ts_all.append(ts); starts.append(fs); counts.append(n)
        out_all.append(np.array([[p.x, p.y, p.z, p.intensity] for p in comp], np.float64))
    o = np.zeros(len(counts) + 1, np.int64); np.cumsum(counts, out=o[1:])
    np.savez_compressed(os.path.join(HERE, 'config2.npz'), pts=np.vstack(pts_all), ts=np.concatenate(ts_all), frame_off=o,
                        frame_start=np.array(starts, np.int64), imu_ts=imu_ts, imu_gyro=imu_gyro, compensated=np.vstack(out_all),
                        traj_time=g['traj_time'], traj_position_gps=g['traj_position_gps'], traj_orientation_imu=g['traj_orientation_imu'])
    manifest['config2'] = dict(points=int(o[-1]), imu_samples=len(imu), compensated_sha256=sha(np.vstack(out_all)))


def cs_run_fixture(CS, manifest):
    """Whole run of the SECOND simulator past its scanner: the reference's own LiDARMotionSimulator object (CS:1884-1946) drives
    _apply_motion_compensation (CS:2086-2105) -> _transform_coordinates (CS:2107-2163, target 'vehicle') ->
    DataExporter.export_point_clouds (CS:1612-1641: .pcd / .xyz / .csv; .las needs laspy) -> LivoxLVXWriter.write_lvx_file
    (CS:245-374, 'lvx2') on 20 ragged synthetic Mid-70-shaped frames (the ray-marching scanner, CS:1108-1146, is out of scope and
    returns a handful of points per minute of CPU) against the reference's own trajectory + 200 Hz IMUSimulator output.  Stored:
    the inputs, the reference's compensated and transformed coordinates, and every output file's bytes + sha256."""
    import tempfile
    rng = np.random.default_rng(4711)
    with tempfile.TemporaryDirectory() as d:
        cfg = {'random_seed': 42, 'duration': 2.0, 'trajectory_type': 'figure_eight', 'environment_complexity': 'simple', 'max_speed': 12.0,
               'enable_motion_compensation': True, 'coordinate_system': 'vehicle', 'lvx_format': 'lvx2', 'log_level': 'ERROR',
               'output_prefix': os.path.join(d, 'run'),
               'device_info': {'lidar_sn': '3GGDJ6K00200101', 'device_type': 1, 'extrinsic_enable': True, 'roll': 0.01, 'pitch': -0.02,
                               'yaw': 0.3, 'x': 0.5, 'y': -0.25, 'z': 1.5}}
        sim = CS.LiDARMotionSimulator(cfg)
        traj = sim.trajectory_generator.generate_trajectory(duration=2.0, max_speed=12.0, max_angular_vel=0.5)
        imu = sim.imu_simulator.simulate_imu_data(traj, 2.0)
        swell = 0.6 * np.sin(np.arange(len(imu))[:, None] * np.array([0.045, 0.027, 0.061]))      # non-trivial rotation rates
        for smp, dl in zip(imu, swell):
            smp.gyro_x += dl[0]; smp.gyro_y += dl[1]; smp.gyro_z += dl[2]
        counts = [400, 0, 1, 513, 96, 97, 300, 1024, 2, 250, 333, 0, 640, 128, 77, 500, 5, 256, 1025, 210]
        frames, all_pts, all_ts, all_tag, all_ring = [], [], [], [], []
        for i, n in enumerate(counts):
            fts = int(i * 0.1 * 1e9)                                                               # CS:2058
            az = np.radians(rng.uniform(-35.2, 35.2, n)); el = np.radians(rng.uniform(-38.6, 38.6, n))
            r = rng.uniform(0.05, 90, n)
            xyz = np.column_stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)]).reshape(n, 3)
            inten = rng.integers(0, 256, n); tag = rng.integers(0, 4, n); ring = np.arange(n) % 16
            ts = fts + np.arange(n, dtype=np.int64) * 1000                                         # CS:1048: 1 us per point
            pts = [CS.LiDARPoint(x=float(xyz[k, 0]), y=float(xyz[k, 1]), z=float(xyz[k, 2]), intensity=int(inten[k]),
                                 timestamp=int(ts[k]), ring=int(ring[k]), tag=int(tag[k])) for k in range(n)]
            frames.append({'frame_id': i, 'timestamp': fts, 'sensor_position': np.zeros(3), 'sensor_orientation': np.zeros(3),
                           'points': pts, 'frame_duration_ns': int(0.1 * 1e9)})
            all_pts.append(np.column_stack([xyz, inten.astype(np.float64)]).reshape(n, 4)); all_ts.append(ts)
            all_tag.append(tag.astype(np.uint8)); all_ring.append(ring.astype(np.int32))
        comp = sim._apply_motion_compensation(frames, imu)
        tr = sim._transform_coordinates(comp, CS.CoordinateSystem.VEHICLE, [])
        assert all(f['motion_compensated'] for f in comp) and all(f['coordinate_system'] == 'vehicle' for f in tr)
        sim.data_exporter.export_point_clouds(tr, cfg['output_prefix'])
        sim.lvx_writer.write_lvx_file(cfg['output_prefix'] + '.lvx2', tr, sim.device_info)
        files = {ext: np.frombuffer(open(cfg['output_prefix'] + '.' + ext, 'rb').read(), np.uint8) for ext in ['pcd', 'xyz', 'csv', 'lvx2']}
        T = sim.coordinate_transformer.transformations[(CS.CoordinateSystem.SENSOR, CS.CoordinateSystem.VEHICLE)]
        di = dict(sim.device_info.__dict__)
    off = np.zeros(len(counts) + 1, np.int64); np.cumsum(counts, out=off[1:])
    comp_xyz = np.array([[p.x, p.y, p.z] for f in comp for p in f['points']], np.float64).reshape(-1, 3)
    tr_xyz = np.array([[p.x, p.y, p.z] for f in tr for p in f['points']], np.float64).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, 'cs_run.npz'), pts=np.vstack(all_pts), ts=np.concatenate(all_ts), tag=np.concatenate(all_tag),
                        ring=np.concatenate(all_ring), frame_off=off, frame_ts=np.array([f['timestamp'] for f in frames], np.int64),
                        imu_ts=np.array([smp.timestamp for smp in imu], np.int64),
                        imu_gyro=np.array([[smp.gyro_x, smp.gyro_y, smp.gyro_z] for smp in imu], np.float64),
                        compensated_xyz=comp_xyz, transformed_xyz=tr_xyz, T_vehicle=T,
                        device_info_json=np.frombuffer(json.dumps(di).encode(), np.uint8),
                        file_pcd=files['pcd'], file_xyz=files['xyz'], file_csv=files['csv'], file_lvx2=files['lvx2'])
    manifest['cs_run'] = dict(frames=len(counts), points=int(off[-1]), imu_samples=len(imu),
                              **{f'{k}_sha256': sha(v) for k, v in files.items()}, **{f'{k}_bytes': int(len(v)) for k, v in files.items()})


def las_fixture(LMC, CS, manifest):
    """LAS parity hook: needs the REAL laspy.  Runs the reference's two LAS call sites --
    LiDARMotionSimulator.save_las (LMC:950-963, laspy header defaults) and DataExporter._export_las (CS:1671-1698, scale 0.001,
    raw intensity, gps_time) -- and stores the files' bytes; tests/test_oracle_golden.py::test_las_parity_pins_itself and
    tests/test_gpu_parity.py::test_las_file_vs_laspy pick the fixture up.  Without laspy this prints why and writes nothing."""
    import tempfile
    try:
        import laspy
        ver = laspy.__version__
        assert hasattr(laspy, 'LasHeader')
    except Exception as e:                   # noqa: BLE001
        print(f"las fixture NOT generated: the real laspy is not importable here ({e!r}); LAS parity stays unpinned")
        return
    rng = np.random.default_rng(1234)
    n = 20_000
    pts = np.column_stack([rng.uniform(-250, 250, (n, 3)), rng.uniform(0, 1, n)])
    pts[:6, :3] = [[0.005, -0.005, 0.015], [0.025, -0.015, 0.0005], [0.0015, -0.0025, 1e-9], [21474.83, -21474.83, 0.0],
                   [1.005, 2.675, -1.005], [123.4565, -0.0049999, 99.995]]        # rounding ties of (v - 0) / 0.01 and / 0.001
    pts5 = np.column_stack([pts[:, :3], np.floor(pts[:, 3] * 255), np.sort(rng.integers(0, 3_600_000_000_000, n)).astype(np.float64)])
    with tempfile.TemporaryDirectory() as d:
        with contextlib.redirect_stdout(io.StringIO()):
            LMC.LiDARMotionSimulator().save_las(pts, os.path.join(d, 'a.las'))
            CS.DataExporter({})._export_las(pts5, os.path.join(d, 'b.las'))
        a = np.frombuffer(open(os.path.join(d, 'a.las'), 'rb').read(), np.uint8)
        b = np.frombuffer(open(os.path.join(d, 'b.las'), 'rb').read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, 'las_ref.npz'), pts=pts, pts5=pts5, file_lmc=a, file_cs=b,
                        laspy_version=np.frombuffer(ver.encode(), np.uint8))
    manifest['las_ref'] = dict(points=n, laspy=ver, lmc_sha256=sha(a), cs_sha256=sha(b))


def main():
    LMC, CS = import_reference()
    if len(sys.argv) > 2 and sys.argv[1] == '--only':          # add / refresh single fixtures, keep the rest of the manifest
        with open(os.path.join(HERE, 'MANIFEST.json')) as f:
            manifest = json.load(f)
        for name in sys.argv[2:]:
            {'lvx_cs': lambda: lvx_cs_fixture(CS, manifest), 'text_rows': lambda: text_rows_fixture(CS, manifest),
             'coord_frames': lambda: coord_frames_fixture(CS, manifest),
             'config2': lambda: config2_fixture(CS, manifest), 'modeb_tiers': lambda: modeb_tiers_fixture(CS, manifest), 'modec_lerp': lambda: modec_lerp_fixture(CS, manifest), 'cs_run': lambda: cs_run_fixture(CS, manifest), 'las': lambda: las_fixture(LMC, CS, manifest),
             'outputs': lambda: [outputs_fixture(LMC, n, manifest) for n in ['C1a', 'C2a', 'C3']],
             'lmc_edge': lambda: lmc_edge_fixture(LMC, manifest), 'lvx_type2': lambda: lvx_type2_fixture(LMC, manifest),
             'lvx_file': lambda: lvx_file_fixture(LMC, manifest), 'modeb': lambda: modeb_fixture(CS, manifest),
             'coord_chain': lambda: coord_chain_fixture(CS, manifest), 'pcd_ascii': lambda: pcd_fixture(LMC, manifest)}[name]()
        with open(os.path.join(HERE, 'MANIFEST.json'), 'w') as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        return
    manifest = {'generator': 'tests/golden/make_golden.py', 'numpy': np.__version__,
                'scipy': __import__('scipy').__version__, 'lmc': {}}
    for name in CONFIGS:
        lmc_fixture(LMC, name, manifest)
        print(name, manifest['lmc'][name]['total_points'], manifest['lmc'][name]['raw_sha256'][:16],
              manifest['lmc'][name]['aligned_sha256'][:16])
    for name in ['C1a', 'C2a', 'C3']:
        scan_fixture(LMC, name, manifest)
    lmc_edge_fixture(LMC, manifest)
    lvx_type2_fixture(LMC, manifest)
    lvx_file_fixture(LMC, manifest)
    modeb_fixture(CS, manifest)
    modeb_tiers_fixture(CS, manifest)
    modec_lerp_fixture(CS, manifest)
    coord_chain_fixture(CS, manifest)
    pcd_fixture(LMC, manifest)
    lvx_cs_fixture(CS, manifest)
    text_rows_fixture(CS, manifest)
    coord_frames_fixture(CS, manifest)
    config2_fixture(CS, manifest)
    cs_run_fixture(CS, manifest)
    las_fixture(LMC, CS, manifest)
    for name in ['C1a', 'C2a', 'C3']:
        outputs_fixture(LMC, name, manifest)
    with open(os.path.join(HERE, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print('wrote', sorted(os.listdir(HERE)))


if __name__ == '__main__':
    main()
