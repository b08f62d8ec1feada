"""LVX v1.1 container around device-quantised point records.

The 14-byte records (the arithmetic of LMC:252-272) come from the GPU; this module only lays out
the container of LivoxLVXWriter.write_compatible_lvx (LMC:58-250): 24-B public header, 5-B
private header, 59-B device block, then per frame a 24-B frame header and ceil(n/96) packages of
22-B header + 96 x 14-B records (tail zero-padded, LMC:246-250).  All offsets are closed-form, so
the layout is a handful of vectorised NumPy scatters (SURVEY.md section 8f row N1 moves this onto
the device).

LMC = /root/reference/lidar_motion_compensation.py
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

from . import _trace

POINTS_PER_PACKAGE = 96          # LMC:48
RECORD = 14
PKG_HEADER = 22
PKG_BYTES = PKG_HEADER + POINTS_PER_PACKAGE * RECORD
FRAME_HEADER = 24
FILE_HEADER = 24 + 5 + 59


def lvx_v11_preamble() -> np.ndarray:
    """Public header + private header + device info (LMC:85-115, 147-172)."""
    h = np.zeros(FILE_HEADER, np.uint8)
    h[0:10] = np.frombuffer(b'livox_tech', np.uint8)               # 16-byte signature, NUL padded
    h[16:20] = [1, 1, 0, 0]                                        # version 1.1.0.0
    h[20:24] = np.frombuffer(np.uint32(0xAC0EA767).tobytes(), np.uint8)
    h[24:28] = np.frombuffer(np.uint32(50).tobytes(), np.uint8)    # frame duration (ms)
    h[28] = 1                                                      # device count
    d = 29
    h[d:d + 15] = np.frombuffer(b'3GGDJ6K00200101', np.uint8)      # LiDAR SN, NUL terminated
    h[d + 33] = 1                                                  # device type (LMC:43)
    return h


def frame_layout(frame_off: np.ndarray):
    """Closed-form byte layout: (packages per frame, frame byte offsets[F+1])."""
    counts = np.diff(np.asarray(frame_off, np.int64))
    pkgs = (counts + POINTS_PER_PACKAGE - 1) // POINTS_PER_PACKAGE
    sizes = FRAME_HEADER + pkgs * PKG_BYTES
    pos = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(sizes, out=pos[1:])
    return pkgs, pos + FILE_HEADER


def build_lvx_v11_file(records: np.ndarray, frame_off: np.ndarray, timestamps: np.ndarray,
                       frame_ids: np.ndarray) -> np.ndarray:
    """Assemble the whole file as one uint8 array (one write())."""
    records = np.ascontiguousarray(records, np.uint8).reshape(-1, RECORD)
    frame_off = np.asarray(frame_off, np.int64)
    F = len(frame_off) - 1
    if F == 0:
        raise ValueError("No frame data provided")                 # LMC:75-76
    pkgs, fpos = frame_layout(frame_off)
    out = np.zeros(int(fpos[-1]), np.uint8)
    out[:FILE_HEADER] = lvx_v11_preamble()
    # frame headers: current offset, next offset (0 for the last frame), frame index  (LMC:179-193)
    fh = np.zeros((F, 3), '<u8')
    fh[:, 0] = fpos[:-1]
    fh[:-1, 1] = fpos[1:-1]
    fh[:, 2] = np.asarray(frame_ids, np.int64)
    idx = fpos[:-1, None] + np.arange(FRAME_HEADER)
    out[idx] = fh.view(np.uint8).reshape(F, FRAME_HEADER)
    # package headers (LMC:206-237)
    P = int(pkgs.sum())
    if P:
        pf = np.repeat(np.arange(F), pkgs)                          # frame of each package
        first = np.zeros(F, np.int64); np.cumsum(pkgs[:-1], out=first[1:])
        pk_local = np.arange(P) - first[pf]
        ppos = fpos[pf] + FRAME_HEADER + pk_local * PKG_BYTES
        hdr = np.zeros((P, PKG_HEADER), np.uint8)
        hdr[:, 1] = 5; hdr[:, 3] = 1; hdr[:, 9] = 1; hdr[:, 10] = 2
        ts_ns = (np.asarray(timestamps, np.float64) * 1e9).astype(np.int64)     # int(t * 1e9), LMC:177
        hdr[:, 14:22] = ts_ns[pf].astype('<u8').view(np.uint8).reshape(P, 8)
        out[ppos[:, None] + np.arange(PKG_HEADER)] = hdr
        # point records
        N = len(records)
        if N:
            counts = np.diff(frame_off)
            pfp = np.repeat(np.arange(F), counts)                  # frame of each point
            j = np.arange(N) - frame_off[pfp]
            dst = fpos[pfp] + FRAME_HEADER + (j // POINTS_PER_PACKAGE) * PKG_BYTES + PKG_HEADER + (j % POINTS_PER_PACKAGE) * RECORD
            out[dst[:, None] + np.arange(RECORD)] = records
    return out


# ---------------------------------------------------------------------------------------------------
# The complete simulator's LivoxLVXWriter (CS = /root/reference/livox_mid70_complete_simulator.py,
# lines 235-374): LVX2 / LVX3 / legacy containers.  The host only builds the few leading bytes that
# depend on DeviceInfo; frames, package headers and records are laid out by the device
# (csrc/lmc_lvx2.cu).
# ---------------------------------------------------------------------------------------------------

@dataclass
class DeviceInfo:                    # CS:131-143
    lidar_sn: str
    device_type: int
    firmware_version: str
    extrinsic_enable: bool
    roll: float
    pitch: float
    yaw: float
    x: float
    y: float
    z: float


def lvx_cs_prefix(format_version: str, device_info, n_frames: int) -> bytes:
    """Leading bytes of the file: lvx2 / lvx3 = 24-B file header + 64-B private header (CS:272-283,
    323-341); lvx = 28-B header + 32-B device block (CS:295-306)."""
    sn = device_info.lidar_sn.encode('ascii').ljust(16, b'\x00')
    if len(sn) != 16:
        raise ValueError("lidar_sn longer than 16 bytes")
    if format_version in ("lvx2", "lvx3"):
        head = b"livox_tech".ljust(10, b'\x00') + b"2.0.0".ljust(6, b'\x00') + struct.pack('<I', 0xAC0EA767) + b'\x00' * 4
        priv = (struct.pack('<II', 50, 1) + sn + struct.pack('<BB', device_info.device_type, 1 if device_info.extrinsic_enable else 0)
                + struct.pack('<ffffff', device_info.roll, device_info.pitch, device_info.yaw, device_info.x, device_info.y, device_info.z)
                + b'\x00' * 14)
        return head + priv
    if format_version == "lvx":
        return b"livox_file" + struct.pack('<II', 1, n_frames) + b'\x00' * 10 + sn + struct.pack('<B', device_info.device_type) + b'\x00' * 15
    raise ValueError(f"Unsupported format version: {format_version}")      # CS:242-243


def frames_to_arrays(frames_data):
    """frames_data (CS:2195: list of {'points': List[LiDARPoint] | (n,4+) ndarray [x y z intensity (tag)], 'timestamp': ns})
    -> (pts (N,4) f64, tag u8[N], frame_off int64[F+1], frame_ts int64[F])."""
    counts = np.array([len(f['points']) for f in frames_data], np.int64)
    off = np.zeros(len(frames_data) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    pts = np.zeros((int(off[-1]), 4), np.float64)
    tag = np.zeros(int(off[-1]), np.uint8)
    for i, f in enumerate(frames_data):
        p = f['points']
        if len(p) == 0:
            continue
        if isinstance(p, np.ndarray):
            pts[off[i]:off[i + 1]] = p[:, :4]
            if p.shape[1] > 4:
                tag[off[i]:off[i + 1]] = p[:, 4].astype(np.uint8)
        else:
            for q in p:
                if not 0 <= q.tag <= 255:
                    raise struct.error("ubyte format requires 0 <= number <= 255")       # what CS:374 raises
            pts[off[i]:off[i + 1]] = [(q.x, q.y, q.z, q.intensity) for q in p]
            tag[off[i]:off[i + 1]] = [q.tag for q in p]
    ts = np.array([int(f['timestamp']) for f in frames_data], np.int64)
    if (ts < 0).any():
        raise struct.error("argument out of range")                                      # '<Q' of a negative timestamp, CS:350
    return pts, tag, off, ts


class LivoxLVXWriter:
    """CS:235-374 -- same constructor and ``write_lvx_file(filename, frames_data, device_info)``; the bytes are
    produced by one device launch (``build_bytes``) and written with one write().  No CPU fallback."""

    supported_versions = ["lvx", "lvx2", "lvx3"]

    def __init__(self, format_version: str = "lvx2", device: str = "cuda:0"):
        if format_version not in self.supported_versions:
            raise ValueError(f"Unsupported format version: {format_version}")          # CS:242-243
        self.format_version = format_version
        self.device = device

    @_trace.traced("LivoxLVXWriter.build_bytes")
    def build_bytes(self, frames_data, device_info) -> np.ndarray:
        import torch
        from . import _capi as C
        from . import ops
        pts, tag, off, ts = frames_to_arrays(frames_data)
        prefix = lvx_cs_prefix(self.format_version, device_info, len(frames_data))
        fmt = C.LVXCS_LEGACY if self.format_version == "lvx" else C.LVXCS_LVX2
        dev = torch.device(self.device)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        data, status = ops.build_lvx_cs(d(pts), d(tag), d(off), d(ts), prefix, fmt, int(np.diff(off).max()) if len(off) > 1 else 0)
        fl = int(status.item())
        if fl & C.FLAG_NAN:
            raise ValueError("cannot convert float NaN to integer")                     # int(nan), CS:368
        if fl & C.FLAG_OVERFLOW:
            raise struct.error("argument out of range")                                 # struct.pack beyond the field, CS:372-373 / 319-320
        return data.cpu().numpy()

    @_trace.traced("LivoxLVXWriter.write_lvx_file")
    def write_lvx_file(self, filename: str, frames_data, device_info) -> None:
        data = self.build_bytes(frames_data, device_info)
        with open(filename, 'wb') as f:
            f.write(data.tobytes())
