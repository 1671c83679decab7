"""Geometry constants of the KITTI configuration — the values of the reference's
config/kitti_config.py:23-47 that the hot path reads (boundary, boundary_back, BEV size,
DISCRETIZATION, bound_size_*).  Same names, so `import config.kitti_config as cnf` call sites work
against this module unchanged.  Class maps, calibration matrices and voxel constants are not part
of the hot path and are not mirrored."""

boundary = {"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}
boundary_back = {"minX": -50, "maxX": 0, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}

bound_size_x = boundary["maxX"] - boundary["minX"]
bound_size_y = boundary["maxY"] - boundary["minY"]
bound_size_z = boundary["maxZ"] - boundary["minZ"]

BEV_WIDTH = 608   # across the y axis, -25 m .. 25 m
BEV_HEIGHT = 608  # across the x axis, 0 m .. 50 m
DISCRETIZATION = (boundary["maxX"] - boundary["minX"]) / BEV_HEIGHT
