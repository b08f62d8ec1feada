"""
Host-side mirror of the complete simulator's ``DataExporter`` point-cloud exports
(/root/reference/livox_mid70_complete_simulator.py = CS, lines 1603-1716), B200-native underneath:
the per-row Python formatting loops (PCD, CS:1663-1664), ``np.savetxt`` (XYZ, CS:1703), the pandas
CSV body (CS:1711-1712) and laspy's packing (LAS, CS:1675-1693; parity unpinned -- laspy is absent)
are device kernels (csrc/lmc_pcd.cu k_text_*, csrc/lmc_las.cu); the host writes the few header
bytes and does one write() per file.  No CPU fallback.

Same method names and file naming as the reference: ``export_point_clouds(frames_data, output_prefix)``
writes ``{prefix}.pcd``, ``{prefix}.las``, ``{prefix}.xyz``, ``{prefix}.csv`` from the merged
``[x, y, z, intensity, timestamp]`` rows (CS:1618-1628).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import _trace
from . import _capi as C
from . import ops

PCD_HEADER = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity timestamp\nSIZE 4 4 4 4 8\n"
              "TYPE F F F F F\nCOUNT 1 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ascii\n")   # CS:1648-1659
CSV_HEADER = "x,y,z,intensity,timestamp\n"                                                                            # CS:1711


def merge_frames(frames_data: List[Dict]) -> np.ndarray:
    """CS:1618-1628: all frames' points as one (N,5) f64 array [x y z intensity timestamp]."""
    parts = []
    for f in frames_data:
        p = f['points']
        if len(p) == 0:
            continue
        if isinstance(p, np.ndarray):
            parts.append(np.asarray(p[:, :5], np.float64))
        else:
            parts.append(np.array([[q.x, q.y, q.z, q.intensity, q.timestamp] for q in p], np.float64))
    return np.vstack(parts) if parts else np.zeros((0, 5), np.float64)


class DataExporter:
    """Multi-format data export (CS:1603-1716) fed from device buffers."""

    def __init__(self, config: Dict):
        self.config = config
        self.coordinate_system = config.get('coordinate_system', 'sensor')
        self.device = torch.device(config.get('device', 'cuda:0'))

    def _dev(self, points) -> torch.Tensor:
        if isinstance(points, torch.Tensor):
            return points.to(self.device, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(self.device)

    @staticmethod
    def _text(rows: torch.Tensor, cols, decimals, sep) -> bytes:
        text, status = ops.text_rows(rows, cols, decimals, sep)
        if int(status.item()) & C.FLAG_OVERFLOW:
            raise OverflowError("value beyond the supported text range (|v| >= 2**64)")
        return text.cpu().numpy().tobytes()

    # -- CS:1612-1641 ----------------------------------------------------------------------------
    @_trace.traced("DataExporter.export_point_clouds")
    def export_point_clouds(self, frames_data: List[Dict], output_prefix: str = "lidar_data"):
        points = merge_frames(frames_data)
        if len(points) == 0:                                       # CS:1622-1624
            return
        d = self._dev(points)
        self._export_pcd(d, f"{output_prefix}.pcd")
        self._export_las(d, f"{output_prefix}.las")
        self._export_xyz(d, f"{output_prefix}.xyz")
        self._export_csv(d, f"{output_prefix}.csv")

    def pcd_bytes(self, points) -> bytes:
        d = self._dev(points)
        return PCD_HEADER.format(n=d.shape[0]).encode() + self._text(d, (0, 1, 2, 3, 4), (6, 6, 6, 0, 0), " ")     # CS:1664

    def xyz_bytes(self, points) -> bytes:
        return self._text(self._dev(points), (0, 1, 2), (6, 6, 6), " ")                                            # CS:1703

    def csv_bytes(self, points) -> bytes:
        d = self._dev(points)
        if bool(torch.isnan(d).any().item()):
            raise ValueError("NaN in the point array: pandas writes empty fields there (CS:1712), not supported")
        return CSV_HEADER.encode() + self._text(d, (0, 1, 2, 3, 4), (6, 6, 6, 6, 6), ",")                          # CS:1711-1712

    def las_bytes(self, points) -> bytes:
        """CS:1675-1693 -- scale 0.001 x3, default offsets, intensity.astype(uint16), gps_time = ts * 1e-9."""
        d = self._dev(points)
        gps = (d[:, 4] * 1e-9).contiguous()
        data, status = ops.build_las_pf3(d[:, :4].contiguous(), scale=(0.001,) * 3, offset=(0.0,) * 3,
                                         intensity_mode=C.LAS_INTENSITY_RAW, gps_time=gps)
        if int(status.item()) & C.FLAG_OVERFLOW:
            raise OverflowError("LAS integer field out of range")
        return data.cpu().numpy().tobytes()

    def _export_pcd(self, points, filename: str):
        with open(filename, 'wb') as f:
            f.write(self.pcd_bytes(points))

    def _export_xyz(self, points, filename: str):
        with open(filename, 'wb') as f:
            f.write(self.xyz_bytes(points))

    def _export_csv(self, points, filename: str):
        with open(filename, 'wb') as f:
            f.write(self.csv_bytes(points))

    def _export_las(self, points, filename: str):
        with open(filename, 'wb') as f:
            f.write(self.las_bytes(points))
