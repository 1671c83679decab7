"""Turns the reference's geometry (module globals of config/kitti_config.py:23-47, or any object
with the same attribute names) into the SfaBevParams struct + density table the C ABI takes.

Every float is rounded to float32 exactly where numpy rounds it when a python scalar meets the
float32 sweep: bounds, DISCRETIZATION, (W+1)/2, max_height."""
import math

import numpy as np

from . import _lib


def density_lut64():
    """float64 the reference stores for a cell with `count` points: min(1, log(count+1)/log(64))
    (data_process/kitti_bev_utils.py:46).  Saturates to exactly 1.0 at count 63."""
    counts = np.arange(64, dtype=np.int64)
    lut = np.minimum(1.0, np.log(counts + 1) / np.log(64))
    lut[0] = 0.0
    return lut


class BevGeometry:
    """One raster configuration.  `cnf` is a module/object with BEV_HEIGHT, BEV_WIDTH, DISCRETIZATION;
    `boundary` is the dict the reference passes around (minX..maxZ)."""

    def __init__(self, boundary, cnf, apply_filter=True, algorithm=_lib.BEV_AUTO):
        self.boundary = dict(boundary)
        self.height = int(cnf.BEV_HEIGHT)
        self.width = int(cnf.BEV_WIDTH)
        self.discretization = float(cnf.DISCRETIZATION)
        self.apply_filter = bool(apply_filter)
        b = self.boundary
        p = _lib.SfaBevParams()
        p.min_x, p.max_x = np.float32(b["minX"]), np.float32(b["maxX"])
        p.min_y, p.max_y = np.float32(b["minY"]), np.float32(b["maxY"])
        p.min_z, p.max_z = np.float32(b["minZ"]), np.float32(b["maxZ"])
        p.discretization = np.float32(self.discretization)
        p.y_offset = np.float32((self.width + 1) / 2)                       # kitti_bev_utils.py:24,29
        p.max_height = np.float32(float(np.abs(b["maxZ"] - b["minZ"])))    # kitti_bev_utils.py:43
        p.height, p.width = self.height, self.width
        p.apply_filter = 1 if apply_filter else 0
        p.algorithm = int(algorithm)
        self.algorithm = int(algorithm)
        self.params = p
        self.lut64 = density_lut64()
        self.lut32 = self.lut64.astype(np.float32)
        assert np.all(np.diff(self.lut32[:64]) > 0), "float32 density table must stay invertible"

    def key(self):
        b = self.boundary
        return (tuple(sorted(b.items())), self.height, self.width, self.discretization, self.apply_filter,
                self.algorithm)

    @property
    def cells(self):
        return self.height * self.width


def from_config(cnf, boundary=None, apply_filter=True, algorithm=_lib.BEV_AUTO):
    return BevGeometry(cnf.boundary if boundary is None else boundary, cnf, apply_filter, algorithm)
