"""Drop-in for `lidar_to_camera_box` of the reference's data_process/transformation.py:99-107 (and the
`lidar_to_camera` :50-60 it loops over): boxes x, y, z, h, w, l, rz in the lidar frame -> x, y, z in the
rect camera frame, h, w, l, ry = -rz - pi/2.  Computed by libsfa_b200.so (`sfa_project_boxes`) in
float64; the batched device form is fast.project_boxes_dense.  The augmentation classes of the same
reference module are training-side and out of scope."""
import numpy as np
import torch

from .. import fast
from ..config import kitti_config as cnf


def _calibration(V2C, R0, P2):
    # transformation.py:52-58: the dataset-average matrices unless BOTH V2C and R0 are given
    if V2C is None or R0 is None:
        V2C, R0 = cnf.Tr_velo_to_cam, cnf.R0
    return V2C, R0, (cnf.P2 if P2 is None else P2)


def lidar_to_camera_box(boxes, V2C=None, R0=None, P2=None):
    """(N, 7) -> (N, 7) float64, like the reference (an empty input gives an empty (0, 7) array)."""
    boxes = np.asarray(boxes, dtype=np.float64)
    if boxes.size == 0:
        return np.array([]).reshape(-1, 7)
    boxes = boxes.reshape(-1, 7)
    if not torch.cuda.is_available():
        raise RuntimeError("libsfa_b200 needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    real = torch.zeros((1, boxes.shape[0], 8), dtype=torch.float64)
    real[0, :, 1:] = torch.from_numpy(boxes)
    calib = fast.pack_calibration(*_calibration(V2C, R0, P2), device=dev)
    _, _, cam = fast.project_boxes_dense(real.to(dev), calib, (1, 1), want_cam=True)
    return cam[0].cpu().numpy()


def lidar_to_camera(x, y, z, V2C=None, R0=None, P2=None):
    """One point; transformation.py:50-60."""
    return tuple(lidar_to_camera_box(np.array([[x, y, z, 0, 0, 0, 0]], dtype=np.float64), V2C, R0, P2)[0, :3])
