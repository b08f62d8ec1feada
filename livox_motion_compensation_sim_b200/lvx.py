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

import numpy as np

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
