"""Drop-in for `lidar_to_camera_box` of the reference's data_process/transformation.py:99-107 (and the
`lidar_to_camera` :50-60 it loops over): boxes x, y, z, h, w, l, rz in the lidar frame -> x, y, z in the
rect camera frame, h, w, l, ry = -rz - pi/2.  Computed by libsfa_b200.so (`sfa_project_boxes`) in
float64; the batched device form is fast.project_boxes_dense.

Also the sweep side of the training augmentation: `point_transform` (:242-285), `Random_Rotation`
(:338-354) and `Random_Scaling` (:357-371) with the reference's np.random call sequence; the rigid
transform runs in `sfa_transform_points` (float64 FMA chains like numpy's dgemm, bit-identical).  The
LABEL side of Random_Rotation (`box_transform`: <= 50 boxes through corner conversions, transformation.py:288-305)
is host code outside the hot path and is not re-implemented here: pass the reference's own `box_transform`
as `box_transform=`.  Without it Random_Rotation REFUSES to rotate a sweep that comes with labels (it raises
instead of silently leaving the targets unrotated); a sweep without labels (inference-time augmentation) needs none."""
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


def _rotation(angle, a, b):
    mat = np.zeros((4, 4))
    for k in range(4):
        if k not in (a, b):
            mat[k, k] = 1
    mat[a, a] = mat[b, b] = np.cos(angle)
    mat[a, b] = -np.sin(angle)
    mat[b, a] = np.sin(angle)
    return mat


def _matrices(tx, ty, tz, rx=0, ry=0, rz=0):
    """The matmul chain of transformation.py:250-283: translation, then rx, ry, rz when non-zero."""
    mat1 = np.eye(4)
    mat1[3, 0:3] = tx, ty, tz
    mats = [mat1]
    for angle, (a, b) in ((rx, (1, 2)), (ry, (2, 0)), (rz, (0, 1))):
        if angle != 0:
            mats.append(_rotation(angle, a, b))
    return np.stack(mats)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("libsfa_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def point_transform(points, tx, ty, tz, rx=0, ry=0, rz=0):
    """(N, 3) -> (N, 3) float64, like the reference."""
    points = np.asarray(points)
    if points.dtype not in (np.float32, np.float64):
        points = points.astype(np.float64)
    if points.shape[0] == 0:
        return np.zeros((0, 3))
    dev = _device()
    mats = torch.from_numpy(_matrices(tx, ty, tz, rx, ry, rz)).to(dev)
    out = fast.transform_points_device(torch.from_numpy(np.ascontiguousarray(points[:, 0:3])).to(dev), mats=mats,
                                       out_dtype=torch.float64)
    return out.cpu().numpy()


class Random_Rotation(object):
    def __init__(self, limit_angle=np.pi / 4, p=0.5, box_transform=None):
        self.limit_angle = limit_angle
        self.p = p
        self.box_transform = box_transform

    def __call__(self, lidar, labels):
        if np.random.random() <= self.p:
            angle = np.random.uniform(-self.limit_angle, self.limit_angle)
            if self.box_transform is None and labels is not None and np.size(labels) > 0:
                # the reference always rotates the labels with the sweep (transformation.py:351)
                raise ValueError("Random_Rotation got labels but no box_transform: pass the reference's "
                                 "data_process.transformation.box_transform (the sweep would be rotated, the labels not)")
            dev = _device()
            pts = torch.from_numpy(lidar).to(dev)          # float32 [N, >= 3], transformed in place like the reference
            mats = torch.from_numpy(_matrices(0, 0, 0, rz=angle)).to(dev)
            lidar[:] = fast.transform_points_device(pts, mats=mats, out=pts).cpu().numpy()
            if self.box_transform is not None:
                labels = self.box_transform(labels, 0, 0, 0, r=angle, coordinate='lidar')
        return lidar, labels


class Random_Scaling(object):
    def __init__(self, scaling_range=(0.95, 1.05), p=0.5):
        self.scaling_range = scaling_range
        self.p = p

    def __call__(self, lidar, labels):
        if np.random.random() <= self.p:
            # the reference draws from (range[0], range[0]) — i.e. always range[0] (transformation.py:366)
            factor = np.random.uniform(self.scaling_range[0], self.scaling_range[0])
            dev = _device()
            pts = torch.from_numpy(lidar).to(dev)
            scales = torch.tensor([factor], dtype=torch.float32, device=dev)
            lidar[:] = fast.transform_points_device(pts, scales=scales, out=pts).cpu().numpy()
            labels[:, 0:6] = labels[:, 0:6] * factor
        return lidar, labels
