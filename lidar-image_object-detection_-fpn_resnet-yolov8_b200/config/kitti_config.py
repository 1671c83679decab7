"""Geometry constants of the KITTI configuration — the values of the reference's
config/kitti_config.py:23-47 that the hot path reads (boundary, boundary_back, BEV size,
DISCRETIZATION, bound_size_*).  Same names, so `import config.kitti_config as cnf` call sites work
against this module unchanged, plus the dataset-average calibration (:64-83) that
data_process.transformation falls back on when a caller passes none.  Class maps and voxel constants
are not part of the hot path and are not mirrored."""
import numpy as np

boundary = {"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}
boundary_back = {"minX": -50, "maxX": 0, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}

bound_size_x = boundary["maxX"] - boundary["minX"]
bound_size_y = boundary["maxY"] - boundary["minY"]
bound_size_z = boundary["maxZ"] - boundary["minZ"]

BEV_WIDTH = 608   # across the y axis, -25 m .. 25 m
BEV_HEIGHT = 608  # across the x axis, 0 m .. 50 m
DISCRETIZATION = (boundary["maxX"] - boundary["minX"]) / BEV_HEIGHT

# dataset-average calibration (reference config/kitti_config.py:64-83), homogeneous 4x4 forms
Tr_velo_to_cam = np.array([
    [7.49916597e-03, -9.99971248e-01, -8.65110297e-04, -6.71807577e-03],
    [1.18652889e-02, 9.54520517e-04, -9.99910318e-01, -7.33152811e-02],
    [9.99882833e-01, 7.49141178e-03, 1.18719929e-02, -2.78557062e-01],
    [0, 0, 0, 1]])
R0 = np.array([
    [0.99992475, 0.00975976, -0.00734152, 0],
    [-0.0097913, 0.99994262, -0.00430371, 0],
    [0.00729911, 0.0043753, 0.99996319, 0],
    [0, 0, 0, 1]])
P2 = np.array([
    [719.787081, 0., 608.463003, 44.9538775],
    [0., 719.787081, 174.545111, 0.1066855],
    [0., 0., 1., 3.0106472e-03],
    [0., 0., 0., 0]])
