"""livox_motion_compensation_sim_b200 -- B200-native (sm_100a) motion-compensation hot path of
manishborikar92/livox-motion-compensation-sim: GPS/IMU pose lookup -> per-point rigid transform /
deskew -> fused LVX/LAS integer quantisation -> merged-cloud assembly.

Importing the package does not need a GPU; calling any operator does (no CPU fallback)."""
from . import _capi, frames
from ._build import build_library, LIB_PATH

__all__ = ["_capi", "frames", "build_library", "LIB_PATH", "LiDARMotionSimulator", "MotionCompensator",
           "LiDARPoint", "IMUData", "ops", "DataExporter", "LivoxLVXWriter", "DeviceInfo"]


def __getattr__(name):          # torch-dependent modules load lazily
    import importlib
    if name in ("ops", "sharding", "simulator", "compensator", "lvx", "exporter", "coords", "pipeline", "synth"):
        return importlib.import_module(f"{__name__}.{name}")
    if name == "LiDARMotionSimulator":
        return importlib.import_module(f"{__name__}.simulator").LiDARMotionSimulator
    if name in ("MotionCompensator", "LiDARPoint", "IMUData"):
        return getattr(importlib.import_module(f"{__name__}.compensator"), name)
    if name == "DataExporter":
        return importlib.import_module(f"{__name__}.exporter").DataExporter
    if name in ("LivoxLVXWriter", "DeviceInfo"):          # the complete simulator's writer (CS:235-374)
        return getattr(importlib.import_module(f"{__name__}.lvx"), name)
    raise AttributeError(name)
