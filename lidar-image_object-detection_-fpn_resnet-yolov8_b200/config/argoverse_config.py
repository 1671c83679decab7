"""Argoverse-range geometry: the boundary and grid size of the reference's
config/argoverse_config.py:8-9,16-27.  DISCRETIZATION follows the rule of config/kitti_config.py:47
(bound_size_x / BEV_HEIGHT = 100/608); the literal 0.1 at argoverse_config.py:10 addresses 1000
cells on a 608-cell map and makes the reference's makeBEVMap raise IndexError."""

boundary = {"minX": -50, "maxX": 50, "minY": -50, "maxY": 50, "minZ": -3, "maxZ": 5}

bound_size_x = boundary["maxX"] - boundary["minX"]
bound_size_y = boundary["maxY"] - boundary["minY"]
bound_size_z = boundary["maxZ"] - boundary["minZ"]

BEV_WIDTH = 608
BEV_HEIGHT = 608
DISCRETIZATION = bound_size_x / BEV_HEIGHT
DISCRETIZATION_LITERAL = 0.1  # the value written in the reference file; unusable with a 608 grid
