"""
Host-side mirror of the reference's ``LiDARMotionSimulator`` for the motion-compensation hot
path (pose lookup -> per-point transform -> merge -> quantise), B200-native underneath.

Same constructor, config keys, ``ValueError`` messages, ``transform_pointcloud`` operator,
``results`` dict contract and output file names as /root/reference/lidar_motion_compensation.py
(LMC).  What changes is *how* the alignment runs: the reference's per-frame Python loop
(LMC:802-832) becomes ONE batched device call over a frame-major point buffer + CSR offsets.

Out of scope here: the trajectory / environment / scanner
generators and the plots.  ``run_simulation`` therefore takes a ``frame_source`` that supplies the
trajectory and the per-frame raw scans -- in an integration that is the reference's own generator
methods (see INTEGRATION.md for the three-line mix-in).

There is no CPU fallback: a CUDA device and liblmc_b200.so are required.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _trace
from . import _capi as C
from . import frames as FR
from . import ops
from .lvx import frame_layout



class LiDARMotionSimulator:
    # ------------------------------------------------------------------ construction (LMC:275-359)
    def __init__(self, config: Optional[Dict] = None):
        self.config = self.default_config()
        if config:
            self._validate_config(config)
            self.config.update(config)          # unknown keys accepted silently, as in LMC:285
        np.random.seed(self.config['random_seed'])          # LMC:288 (generators upstream rely on it)
        self.device = torch.device(self.config.get('device', 'cuda:0'))
        self._performance_stats = {'scan_times': [], 'transform_times': [], 'total_points_processed': 0}

    def default_config(self):
        """All 22 reference keys with the reference's defaults (LMC:297-330) + B200 knobs."""
        return {
            'duration': 60.0, 'lidar_fps': 10, 'imu_rate': 100, 'gps_rate': 5, 'random_seed': 42,
            'max_speed': 15.0, 'max_angular_vel': 0.5, 'trajectory_type': 'figure_eight',
            'fov_horizontal': 70.0, 'fov_vertical': 77.2, 'range_max': 90.0, 'range_min': 0.05,
            'points_per_frame': 96000, 'angular_resolution': 0.28,
            'gps_noise_std': 0.03, 'imu_accel_noise': 0.1, 'imu_gyro_noise': 0.01, 'lidar_range_noise': 0.02,
            'environment_complexity': 'medium', 'ground_height': 0.0, 'obstacle_density': 0.1,
            # --- additional keys (defaults reproduce the reference's behaviour) ---
            'device': 'cuda:0',
            'io_dtype': 'float64',               # 'float64' = reference-exact layout, 'float32' = throughput layout
            'strict_reference_merge': True,      # LMC:887-891: no merged_aligned if any frame is empty
            'las_scale': (0.01, 0.01, 0.01),     # laspy header default used by LMC:953
            'las_offset': (0.0, 0.0, 0.0),
            # 'hold_next' = the reference's per-frame pose (LMC:802-812); 'slerp' = per-point deskew: every point's
            # timestamp is bracketed in the trajectory samples, orientation SLERPed, position lerped (north_star Mode C)
            'pose_interpolation': 'hold_next',
            # False (default): results are ordinary NumPy arrays -- D2H into one reusable pinned staging buffer, then a threaded
            # copy (lmc_host_copy); nothing stays page-locked behind the caller's back.
            # True: opt-in throughput knob -- results ARE page-locked blocks of torch's caching host allocator (no second copy:
            # align_frames 42 -> 7 ms on the reference's run shapes), which stay locked as long as any per-frame view of them
            # lives and are cached afterwards (~2 x 2 GB for the reference's default run)
            'pinned_results': False,
        }

    def _validate_config(self, config: Dict) -> None:
        """Same checks and messages as LMC:332-359."""
        for key in ['duration', 'lidar_fps', 'max_speed', 'range_max', 'range_min', 'points_per_frame']:
            if key in config and not isinstance(config[key], (int, float)):
                raise ValueError(f"Configuration '{key}' must be numeric")
        if 'lidar_fps' in config and config['lidar_fps'] <= 0:
            raise ValueError("LiDAR frame rate must be positive")
        if 'duration' in config and config['duration'] <= 0:
            raise ValueError("Simulation duration must be positive")
        if 'max_speed' in config and config['max_speed'] < 0:
            raise ValueError("Maximum speed cannot be negative")
        if 'range_max' in config and 'range_min' in config:
            if config['range_max'] <= config['range_min']:
                raise ValueError("Maximum range must be greater than minimum range")

    # ------------------------------------------------------------------ device helpers
    def _np_dtype(self):
        return np.float64 if self.config.get('io_dtype', 'float64') == 'float64' else np.float32

    def _pinned_stage(self, n: int, np_dtype) -> torch.Tensor:
        tdt = torch.float64 if np_dtype == np.float64 else torch.float32
        st = getattr(self, '_stage', None)
        if st is None or st.dtype != tdt or st.shape[0] < n:
            cap = max(n, 1 << 16)
            self._stage = st = torch.empty((cap, 4), dtype=tdt, pin_memory=self.device.type == 'cuda')
        return st[:n]

    def _to_dev(self, a: np.ndarray, dtype=None) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(a if dtype is None else a.astype(dtype, copy=False)))
        return t.to(self.device, non_blocking=False)

    def _to_host(self, t: torch.Tensor) -> np.ndarray:
        """Device result -> host array the caller owns."""
        if t.device.type != 'cuda' or t.numel() == 0:
            return t.cpu().numpy()
        if self.config.get('pinned_results', False):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)      # cached block after the first call of a size
            h.copy_(t)
            return h.numpy()                                              # keeps the pinned block alive while referenced
        flat = t.reshape(-1)
        n8 = (flat.numel() * flat.element_size() + 7) // 8
        st = getattr(self, '_stage_out', None)
        if st is None or st.shape[0] < n8:
            self._stage_out = st = torch.empty(max(n8, 1 << 16), dtype=torch.int64, pin_memory=True)
        hb = st.view(torch.uint8)[:flat.numel() * flat.element_size()]
        hb.copy_(flat.view(torch.uint8))
        out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
        C.check(C.lib().lmc_host_copy(out.ctypes.data, hb.data_ptr(), out.nbytes, FR._host_threads()))
        return out

    # ------------------------------------------------------------------ (a1) LMC:802-812
    def lookup_frame_poses(self, trajectory: Dict, lidar_times: np.ndarray):
        """Hold-next pose per frame on the device. Returns (pose_Rt device (F,12), pose_idx np int32)."""
        traj_Rt = FR.pose_table(trajectory['position_gps'], trajectory['orientation_imu'])
        pose, idx = ops.pose_lookup_hold_next(self._to_dev(np.asarray(trajectory['time'], np.float64)),
                                              self._to_dev(traj_Rt),
                                              self._to_dev(np.asarray(lidar_times, np.float64)))
        return pose, idx.cpu().numpy()

    # ------------------------------------------------------------------ (a2) LMC:772-776
    def transform_pointcloud(self, points, transformation):
        """Apply transformation to point cloud -- same contract as LMC:772-776: (n,4) f64 in, new
        (n,4) f64 out, input untouched, empty in -> empty out.  Runs as a one-frame device batch."""
        points = np.asarray(points, np.float64)
        n = len(points)
        if n == 0:
            return np.column_stack([np.zeros((0, 3)), np.zeros(0)])
        pose = FR.pose_table(np.asarray(transformation['translation'], np.float64),
                             np.asarray(transformation['rotation'], np.float64))
        off = torch.tensor([0, n], dtype=torch.int64, device=self.device)
        out, _ = ops.align_rigid(self._to_dev(points), off, self._to_dev(pose))
        return out.cpu().numpy()

    # ------------------------------------------------------------------ (N4) LMC:701-770
    @_trace.traced("LiDARMotionSimulator.scan_all")
    def scan_all(self, environment, positions, eulers, keep_device: bool = False) -> List[np.ndarray]:
        """scan_environment for every frame in ONE device pass (range / FOV cull, compaction, subsample);
        the noise is drawn on the host from the global NumPy RNG exactly as LMC:767 does, so a seeded
        run reproduces the reference's raw scans.  Returns the per-frame (n_f,4) arrays."""
        from scipy.spatial.transform import Rotation
        positions = np.asarray(positions, np.float64).reshape(-1, 3)
        Rm = Rotation.from_euler('xyz', np.asarray(eulers, np.float64).reshape(-1, 3)).as_matrix().reshape(-1, 9)   # LMC:726
        c = self.config
        raw, off = ops.scan_frames(self._to_dev(np.asarray(environment, np.float64)), self._to_dev(positions), self._to_dev(Rm),
                                   range_max=c['range_max'], range_min=c['range_min'], fov_horizontal=c['fov_horizontal'],
                                   fov_vertical=c['fov_vertical'], points_per_frame=c['points_per_frame'],
                                   noise_std=c['lidar_range_noise'])
        self._dev_scan = (raw, off) if keep_device else None      # run_simulation aligns the device copy directly
        host = self._to_host(raw)
        return [host[off[i]:off[i + 1]] if off[i + 1] > off[i] else np.array([]).reshape(0, 4) for i in range(len(off) - 1)]

    def scan_environment(self, environment, sensor_pose):
        """Same contract as LMC:701-770 for one pose (one-frame device batch)."""
        return self.scan_all(environment, [sensor_pose['position']], [sensor_pose['orientation']])[0]

    # ------------------------------------------------------------------ (a2)+(a3) batched
    @_trace.traced("LiDARMotionSimulator.align_frames")
    def align_frames(self, frames: Sequence[np.ndarray], positions: np.ndarray, eulers: np.ndarray,
                     export: Optional[ops.ExportSpec] = None):
        """The reference's frame loop body LMC:826-832 for every frame at once.

        Returns (merged (N,4) host array, frame_off, ExportBuffers|None).  The merged array is
        frame-major == np.vstack(aligned) (LMC:888); per-frame results are views of it."""
        dt = self._np_dtype()
        off = FR.frame_offsets(frames)
        n = int(off[-1])
        pose = FR.pose_table(positions, eulers)
        if n == 0:
            return np.zeros((0, 4), dt), off, None
        # frames are concatenated straight into a pinned staging buffer that is reused across calls
        # (no page faults on a fresh 100 MB array, and the H2D copy runs at the pinned rate)
        stage = self._pinned_stage(n, dt)
        FR.flatten_frames_into(frames, stage.numpy())
        pts_d = stage.to(self.device, non_blocking=True)
        out, bufs = ops.align_rigid(pts_d, self._to_dev(off), self._to_dev(pose), export=export)
        self._performance_stats['total_points_processed'] += n
        return self._to_host(out), off, bufs

    # ------------------------------------------------------------------ LMC:778-858
    @_trace.traced("LiDARMotionSimulator.run_simulation")
    def run_simulation(self, frame_source=None):
        """Run the simulation with the alignment step on the B200.

        ``frame_source`` supplies what the (out-of-scope) generators produce:
          .trajectory                -> dict with 'time','position','velocity','orientation',
                                        'position_gps','orientation_imu'  (LMC:396-428)
          .scan(i, t, sensor_pose)   -> (n,4) f64 raw sensor-frame points   (LMC:701-770); if absent,
                                        .environment is scanned on the device for all frames at once (N4)
          .environment               -> (M,4) f64 world cloud (needed when .scan is absent)
        Returns the reference's results dict (LMC:852-858)."""
        if frame_source is None:
            raise NotImplementedError(
                "trajectory / environment / scanner generators are outside the accelerated hot path "
                "(SURVEY.md section 2 rows 8-11); pass a frame_source, or mix this class into the "
                "reference's LiDARMotionSimulator as shown in INTEGRATION.md")
        trajectory = frame_source.trajectory
        lidar_times = FR.lidar_frame_times(self.config['duration'], self.config['lidar_fps'])
        _, pose_idx = self.lookup_frame_poses(trajectory, lidar_times)

        all_scans, motion_data = [], []
        device_scans = None
        if not hasattr(frame_source, 'scan'):           # no host scanner supplied: scan every frame on the device (N4)
            device_scans = self.scan_all(frame_source.environment, trajectory['position_gps'][pose_idx],
                                         trajectory['orientation_imu'][pose_idx], keep_device=True)
        for i, t in enumerate(lidar_times):
            k = int(pose_idx[i])
            sensor_pose = {'position': trajectory['position_gps'][k],
                           'orientation': trajectory['orientation_imu'][k],
                           'velocity': trajectory['velocity'][k]}
            scan = device_scans[i] if device_scans is not None else frame_source.scan(i, t, sensor_pose)
            all_scans.append({'frame_id': i, 'timestamp': t, 'points_local': scan, 'sensor_pose': sensor_pose})
            motion_data.append(self._motion_row(i, t, sensor_pose))
        dev_scan = getattr(self, '_dev_scan', None) if device_scans is not None else None
        self._dev_scan = None
        if self.config.get('pose_interpolation', 'hold_next') == 'slerp':
            aligned = self.deskew_scans(all_scans, trajectory, _device_raw=dev_scan)
        elif dev_scan is not None and self._np_dtype() == np.float64 and dev_scan[0].shape[0] > 0:
            # the scans are still resident: align them where they are (no flatten, no second upload)
            raw_d, off = dev_scan
            pose = FR.pose_table(trajectory['position_gps'][pose_idx], trajectory['orientation_imu'][pose_idx])
            out, _ = ops.align_rigid(raw_d, self._to_dev(off), self._to_dev(pose))
            self.last_merged, self.last_frame_off, self.last_export = self._to_host(out), off, None
            self._performance_stats['total_points_processed'] += int(off[-1])
            aligned = FR.split_frames(self.last_merged, off)
        else:
            aligned = self.align_scans(all_scans)
        return {'raw_scans': all_scans, 'aligned_pointclouds': aligned, 'motion_data': motion_data,
                'trajectory': trajectory, 'environment': getattr(frame_source, 'environment', None)}

    @staticmethod
    def _motion_row(i, t, sensor_pose):
        """LMC:835-847."""
        p, o, v = sensor_pose['position'], sensor_pose['orientation'], sensor_pose['velocity']
        return {'frame_id': i, 'timestamp': t,
                'gps_lat': p[1] / 111320.0 + 40.0,
                'gps_lon': p[0] / (111320.0 * np.cos(np.radians(40.0))) - 74.0,
                'gps_alt': p[2], 'imu_roll': o[0], 'imu_pitch': o[1], 'imu_yaw': o[2],
                'vel_x': v[0], 'vel_y': v[1], 'vel_z': v[2]}

    @_trace.traced("LiDARMotionSimulator.align_scans")
    def align_scans(self, raw_scans: List[dict], export: Optional[ops.ExportSpec] = None) -> List[np.ndarray]:
        """Batched LMC:826-832 over the reference's raw_scans list; keeps the merged buffer and the
        export buffers on ``self`` (last_merged / last_frame_off / last_export)."""
        frames = [s['points_local'] for s in raw_scans]
        pos = np.array([s['sensor_pose']['position'] for s in raw_scans], np.float64).reshape(-1, 3)
        eul = np.array([s['sensor_pose']['orientation'] for s in raw_scans], np.float64).reshape(-1, 3)
        merged, off, bufs = self.align_frames(frames, pos, eul, export=export)
        self.last_merged, self.last_frame_off, self.last_export = merged, off, bufs
        return FR.split_frames(merged, off)

    @_trace.traced("LiDARMotionSimulator.deskew_scans")
    def deskew_scans(self, raw_scans: List[dict], trajectory: Dict, export: Optional[ops.ExportSpec] = None,
                     _device_raw=None) -> List[np.ndarray]:
        """Per-point deskew + alignment (north_star Mode C; the reference has no such step -- parity is against the
        builder's SciPy Slerp + lerp oracle).  Pose samples = the trajectory's GPS positions / IMU orientations at
        their own times (LMC:396-428); a point's time is ``scan['point_times']`` (int64 ns) when the scan dict has
        it, else frame time + i * (frame period / n_f).  Points outside the sample span hold the
        end pose.  Returns the per-frame world clouds; merged buffer kept as in align_scans."""
        from scipy.spatial.transform import Rotation
        if _device_raw is not None:                                 # scans still resident on the device (run_simulation)
            flat_d, off = _device_raw
        else:
            flat, off = FR.flatten_frames([s['points_local'] for s in raw_scans], np.float64)
            flat_d = None
        n = int(off[-1])
        if n == 0:
            self.last_merged, self.last_frame_off, self.last_export = np.zeros((0, 4)), off, None
            return FR.split_frames(self.last_merged, off)
        if flat_d is None:
            flat_d = self._to_dev(flat)
        period_ns = int(round(1e9 / float(self.config['lidar_fps'])))
        fstart = np.round(np.array([s['timestamp'] for s in raw_scans], np.float64) * 1e9).astype(np.int64)   # same conversion as the sample times below
        ts = np.empty(n, np.int64)
        for i, s in enumerate(raw_scans):
            m = int(off[i + 1] - off[i])
            if m:
                pt = s.get('point_times')
                ts[off[i]:off[i + 1]] = pt if pt is not None else fstart[i] + (np.arange(m, dtype=np.int64) * period_ns) // m
        s_ts = np.round(np.asarray(trajectory['time'], np.float64) * 1e9).astype(np.int64)
        quat = Rotation.from_euler('xyz', np.asarray(trajectory['orientation_imu'], np.float64)).as_quat()
        seg = ops.build_slerp_table(self._to_dev(quat), self._to_dev(np.asarray(trajectory['position_gps'], np.float64)), self._to_dev(s_ts))
        out, bufs = ops.deskew_slerp(flat_d, self._to_dev(ts), self._to_dev(off), self._to_dev(fstart), self._to_dev(s_ts), seg,
                                     export=export)
        self.last_merged, self.last_frame_off, self.last_export = self._to_host(out), off, bufs
        return FR.split_frames(self.last_merged, off)

    # ------------------------------------------------------------------ (a3) LMC:886-899
    def _merge_flags(self, results):
        """The reference's two merge guards (LMC:887-899) with its warnings: (merged_aligned exists, merged_raw exists)."""
        aligned = results['aligned_pointclouds']
        strict = self.config.get('strict_reference_merge', True)
        has_aligned = bool(aligned) and (all(len(pc) > 0 for pc in aligned) or not strict)
        if not has_aligned:
            print("Warning: No aligned point clouds to merge")
        has_raw = any(len(s['points_local']) > 0 for s in results['raw_scans'])
        if not has_raw:
            print("Warning: No raw point clouds to merge")
        return has_aligned, has_raw

    def merge_results(self, results) -> Dict[str, Optional[np.ndarray]]:
        """merged_aligned / merged_raw with the reference's guards."""
        aligned = results['aligned_pointclouds']
        out: Dict[str, Optional[np.ndarray]] = {'merged_aligned': None, 'merged_raw': None}
        has_aligned, has_raw = self._merge_flags(results)
        if has_aligned:
            base = getattr(self, 'last_merged', None)
            same = base is not None and len(aligned) and aligned[0].base is base
            out['merged_aligned'] = base if same else np.vstack(aligned)
        if has_raw:
            out['merged_raw'] = np.vstack([s['points_local'] for s in results['raw_scans'] if len(s['points_local']) > 0])
        return out

    # ------------------------------------------------------------------ (a4) LMC:252-272 / 965-990
    def quantize_lvx(self, results) -> tuple:
        """int32-millimetre LVX type-2 records of the RAW scans (LMC:973-978), on the device.
        Returns ((N,14) uint8 host array, frame_off)."""
        flat, off = FR.flatten_frames([s['points_local'] for s in results['raw_scans']], np.float64)
        if len(flat) == 0:
            return np.zeros((0, 14), np.uint8), off
        bufs = ops.quantize(self._to_dev(flat), ops.ExportSpec(lvx=True, lvx_mode=C.LVX_TYPE2_OF_INPUT))
        bufs.raise_for_flags()
        return self._to_host(bufs.lvx14), off

    # ------------------------------------------------------------------ (a5) LMC:950-963
    def quantize_las(self, points: np.ndarray):
        """LAS integer X/Y/Z + uint16 intensity of the merged cloud (parity unpinned: laspy)."""
        bufs = ops.quantize(self._to_dev(np.asarray(points, np.float64)),
                            ops.ExportSpec(las=True, las_scale=self.config['las_scale'],
                                           las_offset=self.config['las_offset'],
                                           las_intensity_mode=C.LAS_INTENSITY_UNIT))
        bufs.raise_for_flags()
        return (bufs.las_x.cpu().numpy(), bufs.las_y.cpu().numpy(), bufs.las_z.cpu().numpy(),
                bufs.las_intensity.cpu().numpy())

    # ------------------------------------------------------------------ LMC:860-930 (hot-path outputs)
    @_trace.traced("LiDARMotionSimulator.save_results")
    def save_results(self, results, output_dir='lidar_simulation_output'):
        """Writes the reference's hot-path outputs under the reference's names:
        aligned_scans_pcd/aligned_frame_%04d.pcd, raw_scans_pcd/frame_%04d.pcd, merged_aligned.pcd,
        merged_raw_overlapped.pcd, merged_aligned.las, lidar_data.lvx -- plus the two small tables
        motion_data.csv / trajectory.csv (LMC:866-868, 916-927; plain pandas, as in the reference)."""
        os.makedirs(output_dir, exist_ok=True)
        print(f"Saving results to {output_dir}...")
        import pandas as pd
        if results.get('motion_data') is not None:
            pd.DataFrame(results['motion_data']).to_csv(os.path.join(output_dir, 'motion_data.csv'), index=False)
        # LMC:870-899: every per-frame PCD and the two merged PCDs.  One formatting pass per cloud family: the text
        # of the frame-major buffer IS the merged file's body, and each per-frame body is a slice of it
        has_aligned, has_raw = self._merge_flags(results)
        pcd_dir = os.path.join(output_dir, 'raw_scans_pcd'); os.makedirs(pcd_dir, exist_ok=True)
        self.save_pcd_frames([scan['points_local'] for scan in results['raw_scans']],
                             [os.path.join(pcd_dir, f'frame_{scan["frame_id"]:04d}.pcd') for scan in results['raw_scans']],
                             os.path.join(output_dir, 'merged_raw_overlapped.pcd') if has_raw else None)
        aligned_dir = os.path.join(output_dir, 'aligned_scans_pcd'); os.makedirs(aligned_dir, exist_ok=True)
        aligned = results['aligned_pointclouds']
        aligned_d = self.save_pcd_frames(aligned, [os.path.join(aligned_dir, f'aligned_frame_{i:04d}.pcd') for i in range(len(aligned))],
                                         os.path.join(output_dir, 'merged_aligned.pcd') if has_aligned else None)
        try:
            if not has_aligned:
                raise UnboundLocalError("merged_aligned")           # what LMC:903 hits when the merge was skipped
            self.save_las(np.zeros((0, 4)) if aligned_d is None else None, os.path.join(output_dir, 'merged_aligned.las'),
                          _device_pts=aligned_d)                                    # the aligned cloud is still resident
            print("LAS format saved successfully")
        except Exception as e:                                       # LMC:905-906 swallows and prints
            print(f"Could not save LAS format: {e}")
        del aligned_d
        try:
            self.save_lvx(results, os.path.join(output_dir, 'lidar_data'))
            print("LVX formats saved successfully")
        except Exception as e:                                       # LMC:913-914
            print(f"Could not save LVX formats: {e}")
        traj = results.get('trajectory')
        if traj is not None and all(k in traj for k in ('time', 'position', 'position_gps')):
            pd.DataFrame({'time': traj['time'], 'x': traj['position'][:, 0], 'y': traj['position'][:, 1], 'z': traj['position'][:, 2],
                          'x_gps': traj['position_gps'][:, 0], 'y_gps': traj['position_gps'][:, 1],
                          'z_gps': traj['position_gps'][:, 2]}).to_csv(os.path.join(output_dir, 'trajectory.csv'), index=False)
        print("Results saved successfully!")
        return output_dir

    @staticmethod
    def _pcd_header(n: int) -> bytes:
        """LMC:934-945."""
        return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\n"
                "SIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
                f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ascii\n").encode('ascii')

    def save_pcd(self, points, filename):
        """ASCII PCD, byte-identical to LMC:932-948 ('%.6f' per field); the point lines are formatted on the GPU."""
        points = np.asarray(points, np.float64).reshape(-1, 4)
        n = len(points)
        with open(filename, 'wb') as f:
            f.write(self._pcd_header(n))
            if n:
                body, status = ops.pcd_ascii_body(self._to_dev(points))       # (N2) '%.6f' formatting on the device
                if int(status.item()):
                    raise OverflowError("save_pcd: |value| >= 9.2e12 is outside the device formatter's range")
                f.write(body.cpu().numpy().tobytes())

    @_trace.traced("LiDARMotionSimulator.save_pcd_frames")
    def save_pcd_frames(self, frames: Sequence[np.ndarray], filenames: Sequence[str], merged_filename: Optional[str] = None):
        """save_pcd for a whole list of frames (LMC:870-884) plus their np.vstack (LMC:893-899) with ONE upload and
        ONE formatting pass: the frame-major buffer is formatted once, frame f's file body is the byte range between
        the offsets of its first and one-past-last row (``lmc_pcd_ascii_row_offsets``), the merged file's body is the
        whole text.  Every file is byte-identical to what save_pcd writes.  Returns the device copy of the
        frame-major (N,4) f64 buffer (None when there are no points)."""
        if len(frames) != len(filenames):
            raise ValueError("one filename per frame")
        off = FR.frame_offsets(frames)
        n = int(off[-1])
        pts_d = None
        if n:
            stage = self._pinned_stage(n, np.float64)
            FR.flatten_frames_into(frames, stage.numpy())
            pts_d = stage.to(self.device, non_blocking=True)
            text_d, boff_d, status = ops.pcd_ascii_frames(pts_d, self._to_dev(off))   # (syncs: the staging buffer is free again)
            if int(status.item()):
                raise OverflowError("save_pcd: |value| >= 9.2e12 is outside the device formatter's range")
            text = memoryview(self._to_host(text_d))
            boff = boff_d.cpu().numpy()
        for i, fn in enumerate(filenames):
            m = int(off[i + 1] - off[i])
            with open(fn, 'wb') as f:
                f.write(self._pcd_header(m))
                if m:
                    f.write(text[int(boff[i]):int(boff[i + 1])])
        if merged_filename is not None:
            with open(merged_filename, 'wb') as f:
                f.write(self._pcd_header(n))
                if n:
                    f.write(text)
        return pts_d

    @_trace.traced("LiDARMotionSimulator.save_las")
    def save_las(self, points, filename, _device_pts: Optional[torch.Tensor] = None):
        """merged_aligned.las (LMC:950-963): LAS 1.2 / point format 3 with the laspy header defaults the
        reference relies on (scale 0.01, offset 0; config 'las_scale' / 'las_offset'), intensity scaled to
        16 bits.  The whole file image -- header extremes included -- is built on the device.  Parity
        with laspy's bytes is unpinned (laspy is not available); the file follows the LAS 1.2 spec."""
        import datetime
        today = datetime.date.today()
        pts_d = _device_pts if _device_pts is not None else self._to_dev(np.asarray(points, np.float64))
        data, status = ops.build_las_pf3(pts_d, scale=self.config['las_scale'],
                                         offset=self.config['las_offset'], intensity_mode=C.LAS_INTENSITY_UNIT,
                                         year=today.year, day_of_year=today.timetuple().tm_yday)
        ops.ExportBuffers(status=status).raise_for_flags()
        with open(filename, 'wb') as f:
            f.write(memoryview(self._to_host(data)))

    @_trace.traced("LiDARMotionSimulator.save_lvx")
    def save_lvx(self, results, base_filename):
        """lidar_data.lvx with the LVX v1.1 container of LMC:58-250 around device-quantised records."""
        data = self.build_lvx_bytes(results)
        with open(f"{base_filename}.lvx", 'wb') as f:
            f.write(data.tobytes())
        return True

    def build_lvx_bytes(self, results) -> np.ndarray:
        """The whole LVX v1.1 file image (LMC:58-250), quantised and laid out on the device."""
        scans = results['raw_scans']
        if not scans:
            raise ValueError("No frame data provided")                      # LMC:75-76
        flat, off = FR.flatten_frames([s['points_local'] for s in scans], np.float64)
        ts = np.array([s['timestamp'] for s in scans], np.float64)
        ids = np.array([s['frame_id'] for s in scans], np.int64)
        _, fpos = frame_layout(off)
        data, status = ops.build_lvx_v11(self._to_dev(flat), self._to_dev(off), self._to_dev(fpos), self._to_dev(ts),
                                         self._to_dev(ids), int(np.diff(off).max()), size=int(fpos[-1]))
        bufs = ops.ExportBuffers(status=status)
        bufs.raise_for_flags()
        return self._to_host(data)
