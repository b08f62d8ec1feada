#!/usr/bin/env python
"""Recipe for oracle/_ref/ -- the UNMODIFIED reference, staged so that it can travel to the GPU box.
TEST / BENCH INFRASTRUCTURE ONLY (same rule as the rest of oracle/: nothing under the package imports it).

    python oracle/make_ref.py            # in the build container, where /root/reference exists

The reference is two plain Python files with no packaging; `bench.py --impl reference` needs its stock code
path (`LiDARMotionSimulator.transform_pointcloud`, `LivoxLVXWriter._write_frame/_write_point_data_type2`,
`MotionCompensator.compensate_point_cloud`) on the GPU box's host cores, and /root/reference does not exist
there.  This script copies the two files byte for byte into the git-ignored oracle/_ref/ (they never enter
the history; `.gitignore` lists the directory, `.gpurunignore` does not), records their sha256, and writes
the import stubs for the third-party modules the reference imports unconditionally but this image lacks
(laspy for LMC:12; matplotlib / mpl_toolkits for CS:26-28).  The stubs are only put on sys.path for modules
that are really absent.

`load()` returns (LMC module, CS module) from oracle/_ref, or None when the directory was never made.
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib
import io
import json
import logging
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("LMC_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("lidar_motion_compensation.py", "livox_mid70_complete_simulator.py")

STUBS = {
    "laspy.py": '"""stub: laspy is not installed in this image (LMC:12 imports it unconditionally)."""\n',
    "matplotlib/__init__.py": '"""stub: matplotlib is not installed in this image (CS:26-28)."""\n',
    "matplotlib/pyplot.py": "",
    "matplotlib/animation.py": "FuncAnimation = object\n",
    "mpl_toolkits/__init__.py": "",
    "mpl_toolkits/mplot3d.py": "Axes3D = object\n",
}


def make(force: bool = False) -> str:
    """Stage the reference under oracle/_ref/ (no-op when already staged from the same sources)."""
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} does not exist: oracle/_ref can only be made in the build container")
    os.makedirs(REF_DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(REF_SRC, name), os.path.join(REF_DST, name)
        data = open(src, "rb").read()
        manifest[name] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    for rel, text in STUBS.items():
        p = os.path.join(REF_DST, "_stubs", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": manifest, "note": "byte-for-byte copies; git-ignored"}, f, indent=1)
    return REF_DST


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, n)) for n in FILES)


def load():
    """(LMC, CS) imported from oracle/_ref with stubs for absent third-party modules; None if not staged."""
    if not available():
        return None
    stubs = os.path.join(REF_DST, "_stubs")
    need_stub = False
    for mod in ("laspy", "matplotlib"):
        try:
            importlib.import_module(mod)
        except ImportError:
            need_stub = True
    if need_stub and stubs not in sys.path:
        sys.path.append(stubs)                       # appended: a real package always wins
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    prev = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            LMC = importlib.import_module("lidar_motion_compensation")
            CS = importlib.import_module("livox_mid70_complete_simulator")
    finally:
        logging.disable(prev)
    return LMC, CS


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
