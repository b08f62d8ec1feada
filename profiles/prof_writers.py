#!/usr/bin/env python
"""Launch every SURVEY 8(f) kernel once on a bounded sample, for ncu (profiles/README.md):

    ncu --set full --clock-control none --import-source on \
        -k regex:'k_lvx|k_las_records|k_pcd|k_text|k_scan_mark|k_scan_emit|k_homog' -o gpurun_out/prof_writers \
        python profiles/prof_writers.py

Sample: the first 3 600 frames x 10 000 points of the M-1H stream (3.6e7 points, 576 MB of float4 input,
well above the 126 MB L2).  Prints the algorithmic bytes of every launch, which the summary compares with
dram__bytes_read + dram__bytes_write.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from livox_motion_compensation_sim_b200 import _capi as C, ops, synth  # noqa: E402
from livox_motion_compensation_sim_b200.lvx import frame_layout  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    F, P = 3600, 10_000
    st = synth.make_stream(F, P, 4242, device=dev, dtype=torch.float32)
    n = st.n_points
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    _, fpos = frame_layout(st.frame_off)
    fpos_d, off_d, ft_d, id_d = d(fpos), d(st.frame_off), d(st.frame_t), d(np.arange(F, dtype=np.int64))
    ts_d = d((st.frame_t * 1e9).astype(np.int64))
    raw = st.pts
    alg = {}
    for _ in range(reps):
        out, _s = ops.build_lvx_v11(raw, off_d, fpos_d, ft_d, id_d, P)
        alg["k_lvx_v11"] = n * 16 + out.numel()
        out, _s = ops.build_lvx_cs(raw, None, off_d, ts_d, bytes(88), C.LVXCS_LVX2, P)
        alg["k_lvx_cs"] = n * 16 + out.numel()
        out, _s = ops.build_las_pf3(raw, scale=(0.001,) * 3, offset=(0.0,) * 3)
        alg["k_las_records"] = n * 16 + out.numel()
        out, _s = ops.pcd_ascii_body(raw)
        alg["k_pcd_len"] = n * 16
        alg["k_pcd_write"] = n * 16 + out.numel()
        rows5 = torch.cat([raw.double(), torch.arange(n, device=dev, dtype=torch.float64).unsqueeze(1) * 1000.0], dim=1).contiguous()
        rows5[:, 3] = torch.floor(rows5[:, 3] * 255.0)
        out, _s = ops.text_rows(rows5, (0, 1, 2, 3, 4), (6, 6, 6, 0, 0), " ")
        alg["k_text_len"] = n * 40
        alg["k_text_write"] = n * 40 + out.numel()
        del rows5, out
        T = np.eye(4)
        T[:3, :3] = [[0.36, -0.8, 0.48], [0.48, 0.6, 0.64], [-0.8, 0.0, 0.6]]
        T[:3, 3] = [1.0, 2.0, 3.0]
        o = ops.transform_homog(raw, T)
        alg["k_homog<f32>"] = n * 32
        del o
        # scanner: the C2a-sized problem (600 frames x ~90 k environment points)
        rng = np.random.default_rng(7)
        M, Fs = 90_000, 600
        env = d(np.column_stack([rng.uniform(-80, 80, (M, 2)), rng.uniform(0, 12, M), rng.uniform(0.1, 0.9, M)]))
        pos = d(np.column_stack([30 * np.sin(np.linspace(0, 6, Fs)), 30 * np.sin(np.linspace(0, 12, Fs)), np.full(Fs, 1.5)]))
        from scipy.spatial.transform import Rotation
        Rm = d(Rotation.from_euler('xyz', np.column_stack([np.zeros(Fs), np.zeros(Fs), np.linspace(0, 6, Fs)])).as_matrix().reshape(Fs, 9))
        rs, fo = ops.scan_frames(env, pos, Rm, range_max=100.0, range_min=0.05, fov_horizontal=70.4, fov_vertical=77.2,
                                 points_per_frame=96_000, noise_std=0.0)
        alg["k_scan_mark"] = Fs * M * (32 + 1)
        alg["k_scan_emit"] = Fs * M * 1 + int(fo[-1]) * (32 + 32)
        torch.cuda.synchronize()
    print(json.dumps({"points": n, "algorithmic_bytes": alg}))


if __name__ == "__main__":
    main()
