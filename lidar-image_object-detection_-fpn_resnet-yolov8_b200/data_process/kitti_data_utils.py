"""Drop-in for `get_filtered_lidar` of the reference's data_process/kitti_data_utils.py:228-251,
computed on the B200 (order-preserving compaction kernel).  Label parsing, calibration and the
heat-map target helpers of that file are training-side and out of scope."""
import numpy as np
import torch

from ..config import kitti_config as cnf
from .. import geometry as _geometry


def get_filtered_lidar(lidar, boundary, labels=None):
    """Points inside the inclusive boundary box, in their original order, with z -= minZ; the input
    array is left untouched.  With `labels`, also returns the labels inside the half-open box
    (kitti_data_utils.py:243-249 — a handful of rows, filtered on the host)."""
    from ..fast import filter_lidar_device
    if not torch.cuda.is_available():
        raise RuntimeError("get_filtered_lidar needs a CUDA device (there is no CPU fallback)")
    pts = np.ascontiguousarray(lidar, dtype=np.float32)
    extra = None
    if pts.shape[1] != 4:
        raise ValueError("lidar must be [N,4] float32 (x, y, z, intensity)")
    geom = _geometry.BevGeometry(boundary, cnf, apply_filter=True)
    dev = filter_lidar_device(torch.from_numpy(pts).cuda(), geom)
    out = dev.cpu().numpy()
    if labels is not None:
        minX, maxX, minY, maxY = boundary["minX"], boundary["maxX"], boundary["minY"], boundary["maxY"]
        minZ, maxZ = boundary["minZ"], boundary["maxZ"]
        keep = ((labels[:, 1] >= minX) & (labels[:, 1] < maxX) & (labels[:, 2] >= minY) & (labels[:, 2] < maxY) &
                (labels[:, 3] >= minZ) & (labels[:, 3] < maxZ))
        return out, labels[keep]
    return out
