#!/usr/bin/env python
"""CUDA-event timing of the SURVEY 8(f) writer kernels on the first 3 600 frames of the M-1H stream (3.6e7 points):
the same measurement as bench.py's `writers` key, stand-alone for A/B runs.  Prints one JSON line."""
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
    raw64 = raw.double()
    peak = 6550.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:                              # noqa: BLE001
        pass

    def t_ms(fn, reps=10):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    res = {"points": n, "peak_GBps": peak}

    def add(name, ms, nbytes):
        res[name] = {"ms": ms, "Gpts_per_s": n / ms / 1e6, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    txt = ops.pcd_ascii_body(raw)[0].numel()
    add("pcd_ascii_f32", t_ms(lambda: ops.pcd_ascii_body(raw)), 2 * n * 16 + txt)                 # the points are read twice: size pass, write pass
    txt64 = ops.pcd_ascii_body(raw64)[0].numel()
    add("pcd_ascii_f64", t_ms(lambda: ops.pcd_ascii_body(raw64)), 2 * n * 32 + txt64)
    add("pcd_ascii_frames_f32", t_ms(lambda: ops.pcd_ascii_frames(raw, off_d)), 2 * n * 16 + txt)
    if len(sys.argv) > 1 and sys.argv[1] == "pcd":                 # A/B runs of the formatter alone
        print(json.dumps(res))
        return
    add("lvx_v11", t_ms(lambda: ops.build_lvx_v11(raw, off_d, fpos_d, ft_d, id_d, P, size=int(fpos[-1]))), n * 16 + int(fpos[-1]))
    add("lvx2", t_ms(lambda: ops.build_lvx_cs(raw, None, off_d, ts_d, bytes(88), C.LVXCS_LVX2, P)), n * 16 + 88 + 45 * F + 14 * n)
    add("las_pf3", t_ms(lambda: ops.build_las_pf3(raw, scale=(0.001,) * 3)), n * 16 + 227 + 34 * n)
    rows5 = torch.cat([raw64, torch.arange(n, device=dev, dtype=torch.float64).unsqueeze(1) * 1000.0], dim=1).contiguous()
    rows5[:, 3] = torch.floor(rows5[:, 3] * 255.0)
    t5 = ops.text_rows(rows5, (0, 1, 2, 3, 4), (6, 6, 6, 0, 0), " ")[0].numel()
    add("cs_pcd_text", t_ms(lambda: ops.text_rows(rows5, (0, 1, 2, 3, 4), (6, 6, 6, 0, 0), " "), reps=5), 2 * n * 40 + t5)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
