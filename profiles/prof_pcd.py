#!/usr/bin/env python
"""Launch the ASCII PCD formatter (size pass + write pass) on the 3.6e7-point sample, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'k_pcd' -o gpurun_out/prof_pcd python profiles/prof_pcd.py [f32|f64]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from livox_motion_compensation_sim_b200 import ops, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = synth.make_stream(3600, 10_000, 4242, device=dev, dtype=torch.float32)
    raw = st.pts
    which = sys.argv[1] if len(sys.argv) > 1 else "f32"
    src = raw.double() if which == "f64" else raw
    out, _ = ops.pcd_ascii_body(src)
    torch.cuda.synchronize()
    print(json.dumps({"points": st.n_points, "text_bytes": out.numel(), "algorithmic_bytes": {"k_pcd_len": st.n_points * (32 if which == "f64" else 16), "k_pcd_write": st.n_points * (32 if which == "f64" else 16) + out.numel()}}))


if __name__ == "__main__":
    main()
