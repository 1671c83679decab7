"""CPU ORACLE for the SFA3D point-cloud hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this file.  The product (the package next to this directory) never imports it and has no
CPU fallback: without the CUDA library it raises.

Parity status: PINNED.  Every function below is checked against the reference's own functions
executed in the build container (tests/golden/make_golden.py imports /root/reference in place and
writes tests/golden/*.npz; tests/test_oracle_golden.py replays them, and
tests/test_oracle_vs_reference_live.py re-runs the comparison live whenever /root/reference exists).
The reference ships no golden vectors or tests of its own, so those generated
fixtures are the pin.

Each function restates one reference function with the same third-party calls (numpy lexsort /
unique, torch max_pool2d / topk), so that timing it is a fair "port" CPU baseline:

  get_filtered_lidar   <- data_process/kitti_data_utils.py:228-251
  makeBEVMap           <- data_process/kitti_bev_utils.py:22-55   (lexsort + unique formulation)
  make_bev_scatter     <- same result, independent scatter-max formulation; this is
                          the algorithm the CUDA kernel uses, restated in numpy as a second opinion
  _nms / _topk / decode <- utils/evaluation_utils.py:21-26, :47-62, :77-105
  post_processing      <- utils/evaluation_utils copy.py:112-143 (per-sample append; the live
                          utils/evaluation_utils.py:112-163 appends outside the loop and so returns
                          only the last sample — see post_processing_live_semantics)
  _sigmoid             <- utils/torch_utils.py:44-45
  convert_det_to_real_values <- utils/evaluation_utils.py:177-193

Geometry is an explicit argument here (the reference reads module globals,
config/kitti_config.py:23-47).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

try:  # torch is only needed for the decode half
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None
    F = None


# ----------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class Geometry:
    """config/kitti_config.py:23-47 as a value instead of module globals."""
    boundary: dict = field(default_factory=lambda: {
        "minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27})
    BEV_HEIGHT: int = 608
    BEV_WIDTH: int = 608
    DISCRETIZATION: float = None  # default: (maxX - minX) / BEV_HEIGHT   (kitti_config.py:47)

    def __post_init__(self):
        if self.DISCRETIZATION is None:
            object.__setattr__(self, "DISCRETIZATION",
                               (self.boundary["maxX"] - self.boundary["minX"]) / self.BEV_HEIGHT)

    @property
    def bound_size_x(self):
        return self.boundary["maxX"] - self.boundary["minX"]

    @property
    def bound_size_y(self):
        return self.boundary["maxY"] - self.boundary["minY"]

    @property
    def bound_size_z(self):
        return self.boundary["maxZ"] - self.boundary["minZ"]


KITTI = Geometry()
KITTI_BACK = Geometry(boundary={"minX": -50, "maxX": 0, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27},
                      DISCRETIZATION=50 / 608)                      # kitti_config.py:36-43
# config/argoverse_config.py:16-23 range, 608x608 grid (:8-9), D by the rule of kitti_config.py:47
# (the literal DISCRETIZATION=0.1 at argoverse_config.py:10 makes the reference raise IndexError).
ARGOVERSE = Geometry(boundary={"minX": -50, "maxX": 50, "minY": -50, "maxY": 50, "minZ": -3, "maxZ": 5})


# ----------------------------------------------------------------------------- stage A
def get_filtered_lidar(lidar, boundary, labels=None):
    """data_process/kitti_data_utils.py:228-251 (point filter; label filter kept for signature parity)."""
    minX, maxX = boundary["minX"], boundary["maxX"]
    minY, maxY = boundary["minY"], boundary["maxY"]
    minZ, maxZ = boundary["minZ"], boundary["maxZ"]
    mask = np.where((lidar[:, 0] >= minX) & (lidar[:, 0] <= maxX) &
                    (lidar[:, 1] >= minY) & (lidar[:, 1] <= maxY) &
                    (lidar[:, 2] >= minZ) & (lidar[:, 2] <= maxZ))
    lidar = lidar[mask]
    lidar[:, 2] = lidar[:, 2] - minZ
    if labels is not None:
        keep = ((labels[:, 1] >= minX) & (labels[:, 1] < maxX) &
                (labels[:, 2] >= minY) & (labels[:, 2] < maxY) &
                (labels[:, 3] >= minZ) & (labels[:, 3] < maxZ))
        return lidar, labels[keep]
    return lidar


def makeBEVMap(PointCloud_, boundary, geom: Geometry = KITTI):
    """data_process/kitti_bev_utils.py:22-55, same numpy calls, geometry passed explicitly."""
    Height = geom.BEV_HEIGHT + 1
    Width = geom.BEV_WIDTH + 1
    pc = np.copy(PointCloud_)
    pc[:, 0] = np.int_(np.floor(pc[:, 0] / geom.DISCRETIZATION))                       # :28
    pc[:, 1] = np.int_(np.floor(pc[:, 1] / geom.DISCRETIZATION) + Width / 2)           # :29
    order = np.lexsort((-pc[:, 2], pc[:, 1], pc[:, 0]))                                # :32
    pc = pc[order]
    _, first, counts = np.unique(pc[:, 0:2], axis=0, return_index=True, return_counts=True)  # :34
    top = pc[first]
    heightMap = np.zeros((Height, Width))
    intensityMap = np.zeros((Height, Width))
    densityMap = np.zeros((Height, Width))
    max_height = float(np.abs(boundary["maxZ"] - boundary["minZ"]))                    # :43
    r, c = np.int_(top[:, 0]), np.int_(top[:, 1])
    heightMap[r, c] = top[:, 2] / max_height                                           # :44
    intensityMap[r, c] = top[:, 3]                                                     # :47
    densityMap[r, c] = np.minimum(1.0, np.log(counts + 1) / np.log(64))                # :46,48
    out = np.zeros((3, Height - 1, Width - 1))
    out[2] = densityMap[:geom.BEV_HEIGHT, :geom.BEV_WIDTH]
    out[1] = heightMap[:geom.BEV_HEIGHT, :geom.BEV_WIDTH]
    out[0] = intensityMap[:geom.BEV_HEIGHT, :geom.BEV_WIDTH]
    return out


def density_lut64():
    """float64 value the reference stores for a cell holding `count` points (kitti_bev_utils.py:46);
    exactly 1.0 from count 63 on, so 64 entries cover every count."""
    counts = np.arange(64, dtype=np.int64)
    lut = np.minimum(1.0, np.log(counts + 1) / np.log(64))
    lut[0] = 0.0
    return lut


def bev_cell_selection(points, geom: Geometry = KITTI, apply_filter=True):
    """Scatter-max restatement: per occupied (row, col) the winning ORIGINAL point
    index and the point count — the integer quantities that must be bit-exact.

    Returns (rows, cols, winner_index, counts, zs) for cells inside the cropped H x W map.
    Follows kitti_data_utils.py:237-241 (filter, z shift), kitti_bev_utils.py:28-29 (discretise),
    :32-35 (highest z, stable tie order = lowest original index), :50-53 (crop of row/col H, W)."""
    pts = np.asarray(points, dtype=np.float32)
    H, W = geom.BEV_HEIGHT, geom.BEV_WIDTH
    Hm, Wm = H + 1, W + 1
    b = geom.boundary
    idx = np.arange(pts.shape[0], dtype=np.int64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    if apply_filter:
        f = np.float32
        keep = ((x >= f(b["minX"])) & (x <= f(b["maxX"])) & (y >= f(b["minY"])) & (y <= f(b["maxY"])) &
                (z >= f(b["minZ"])) & (z <= f(b["maxZ"])))
        x, y, z, idx = x[keep], y[keep], z[keep], idx[keep]
        z = z - np.float32(b["minZ"])
    d32 = np.float32(geom.DISCRETIZATION)
    ix = np.floor(x / d32).astype(np.int64)
    iy = (np.floor(y / d32) + np.float32(Wm / 2)).astype(np.int64)     # astype truncates toward zero
    if ix.size and (ix.min() < -Hm or ix.max() >= Hm or iy.min() < -Wm or iy.max() >= Wm):
        raise IndexError("point outside the %dx%d map (the reference raises here too)" % (Hm, Wm))
    row = np.where(ix < 0, ix + Hm, ix)
    col = np.where(iy < 0, iy + Wm, iy)
    inside = (row < H) & (col < W)
    row, col, z, idx = row[inside], col[inside], z[inside], idx[inside]
    cell = row * W + col
    # highest z first, then lowest original index (stable lexsort tie order); NaN z sorts last
    zkey = np.where(np.isnan(z), -np.inf, z.astype(np.float64))
    order = np.lexsort((idx, -zkey, cell))
    cell_s = cell[order]
    first = np.ones(cell_s.shape[0], dtype=bool)
    first[1:] = cell_s[1:] != cell_s[:-1]
    starts = np.flatnonzero(first)
    counts = np.diff(np.append(starts, cell_s.shape[0]))
    win = order[starts]
    return row[win], col[win], idx[win], counts, z[win]


def make_bev_scatter(points, geom: Geometry = KITTI, apply_filter=True, dtype=np.float32):
    """[3,H,W] map built from bev_cell_selection: ch0 intensity, ch1 height, ch2 density
    (kitti_bev_utils.py:50-53).  float32 output equals `reference.astype(float32)` bit for bit."""
    pts = np.asarray(points, dtype=np.float32)
    H, W = geom.BEV_HEIGHT, geom.BEV_WIDTH
    rows, cols, win, counts, zs = bev_cell_selection(pts, geom, apply_filter)
    out = np.zeros((3, H, W), dtype=np.float64)
    max_height = np.float32(float(np.abs(geom.boundary["maxZ"] - geom.boundary["minZ"])))
    out[1, rows, cols] = zs / max_height
    out[0, rows, cols] = pts[win, 3]
    out[2, rows, cols] = density_lut64()[np.minimum(counts, 63)]
    return out.astype(dtype)


# ----------------------------------------------------------------------------- stage B
def _sigmoid(x):
    """utils/torch_utils.py:44-45 (in-place sigmoid, then clamp)."""
    return torch.clamp(x.sigmoid_(), min=1e-4, max=1 - 1e-4)


def _nms(heat, kernel=3):
    """utils/evaluation_utils.py:21-26."""
    pad = (kernel - 1) // 2
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=pad)
    keep = (hmax == heat).float()
    return heat * keep


def _gather_feat(feat, ind):
    """utils/evaluation_utils.py:29-37 (mask branch is unused on this path)."""
    dim = feat.size(2)
    ind = ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim)
    return feat.gather(1, ind)


def _transpose_and_gather_feat(feat, ind):
    """utils/evaluation_utils.py:40-44."""
    feat = feat.permute(0, 2, 3, 1).contiguous()
    feat = feat.view(feat.size(0), -1, feat.size(3))
    return _gather_feat(feat, ind)


def _topk(scores, K=40):
    """utils/evaluation_utils.py:47-62 (two-stage: per class, then across the C*K candidates)."""
    batch, cat, height, width = scores.size()
    topk_scores, topk_inds = torch.topk(scores.view(batch, cat, -1), K)
    topk_inds = topk_inds % (height * width)
    topk_ys = (torch.floor_divide(topk_inds, width)).float()
    topk_xs = (topk_inds % width).int().float()
    topk_score, topk_ind = torch.topk(topk_scores.view(batch, -1), K)
    topk_clses = (torch.floor_divide(topk_ind, K)).int()
    topk_inds = _gather_feat(topk_inds.view(batch, -1, 1), topk_ind).view(batch, K)
    topk_ys = _gather_feat(topk_ys.view(batch, -1, 1), topk_ind).view(batch, K)
    topk_xs = _gather_feat(topk_xs.view(batch, -1, 1), topk_ind).view(batch, K)
    return topk_score, topk_inds, topk_clses, topk_ys, topk_xs


def decode(hm_cen, cen_offset, direction, z_coor, dim, K=40, return_inds=False):
    """utils/evaluation_utils.py:77-105.  Columns: score, x, y, z, dim(h,w,l), dir(im,re), cls."""
    batch_size = hm_cen.size(0)
    hm_cen = _nms(hm_cen)
    scores, inds, clses, ys, xs = _topk(hm_cen, K=K)
    if cen_offset is not None:
        off = _transpose_and_gather_feat(cen_offset, inds).view(batch_size, K, 2)
        xs = xs.view(batch_size, K, 1) + off[:, :, 0:1]
        ys = ys.view(batch_size, K, 1) + off[:, :, 1:2]
    else:
        xs = xs.view(batch_size, K, 1) + 0.5
        ys = ys.view(batch_size, K, 1) + 0.5
    direction = _transpose_and_gather_feat(direction, inds).view(batch_size, K, 2)
    z_coor = _transpose_and_gather_feat(z_coor, inds).view(batch_size, K, 1)
    dim = _transpose_and_gather_feat(dim, inds).view(batch_size, K, 3)
    det = torch.cat([scores.view(batch_size, K, 1), xs, ys, z_coor, dim, direction,
                     clses.view(batch_size, K, 1).float()], dim=2)
    if return_inds:
        return det, inds, clses
    return det


def get_yaw(direction):
    """utils/evaluation_utils.py:108-109."""
    return np.arctan2(direction[:, 0:1], direction[:, 1:2])


def post_processing(detections, num_classes=3, down_ratio=4, peak_thresh=0.2, geom: Geometry = KITTI):
    """Per-sample semantics of `utils/evaluation_utils copy.py:112-143` (ret.append inside the batch
    loop), without the prints the live copy adds."""
    ret = []
    for i in range(detections.shape[0]):
        top_preds = {}
        classes = detections[i, :, -1]
        for j in range(num_classes):
            inds = (classes == j)
            top_preds[j] = np.concatenate([
                detections[i, inds, 0:1],
                detections[i, inds, 1:2] * down_ratio,
                detections[i, inds, 2:3] * down_ratio,
                detections[i, inds, 3:4],
                detections[i, inds, 4:5],
                detections[i, inds, 5:6] / geom.bound_size_y * geom.BEV_WIDTH,
                detections[i, inds, 6:7] / geom.bound_size_x * geom.BEV_HEIGHT,
                get_yaw(detections[i, inds, 7:9]).astype(np.float32)], axis=1)
            if len(top_preds[j]) > 0:
                keep_inds = (top_preds[j][:, 0] > peak_thresh)
                top_preds[j] = top_preds[j][keep_inds]
        ret.append(top_preds)
    return ret


def post_processing_live_semantics(detections, **kw):
    """What the live utils/evaluation_utils.py:112-163 returns: [] for an empty batch (:124-126),
    otherwise a 1-element list holding the LAST sample only (append at :158 is outside the loop)."""
    if detections.shape[0] == 0:
        return []
    return post_processing(detections, **kw)[-1:]


def convert_det_to_real_values(detections, num_classes=3, geom: Geometry = KITTI):
    """utils/evaluation_utils.py:177-193."""
    out = []
    for cls_id in range(num_classes):
        if len(detections[cls_id]) > 0:
            for det in detections[cls_id]:
                _score, _x, _y, _z, _h, _w, _l, _yaw = det
                _yaw = -_yaw
                x = _y / geom.BEV_HEIGHT * geom.bound_size_x + geom.boundary["minX"]
                y = _x / geom.BEV_WIDTH * geom.bound_size_y + geom.boundary["minY"]
                z = _z + geom.boundary["minZ"]
                w = _w / geom.BEV_WIDTH * geom.bound_size_y
                l = _l / geom.BEV_HEIGHT * geom.bound_size_x
                out.append([cls_id, x, y, z, _h, w, l, _yaw])
    return np.array(out)


# ----------------------------------------------------------------------------- training-side point augmentation
def point_transform(points, tx, ty, tz, rx=0, ry=0, rz=0):
    """data_process/transformation.py:242-285: homogeneous row vectors times a translation matrix and
    up to three rotation matrices, each a separate float64 matmul.  (numpy's dgemm accumulates the four
    products of a row in order with fused multiply-adds; the CUDA kernel does the same.)"""
    N = points.shape[0]
    points = np.hstack([points, np.ones((N, 1))])
    mat1 = np.eye(4)
    mat1[3, 0:3] = tx, ty, tz
    points = np.matmul(points, mat1)
    for angle, (a, b) in ((rx, (1, 2)), (ry, (2, 0)), (rz, (0, 1))):
        if angle != 0:
            points = np.matmul(points, rotation_matrix(angle, a, b))
    return points[:, 0:3]


def rotation_matrix(angle, a, b):
    """The 4x4 matrices of transformation.py:256-283: rotation in the (a, b) plane,
    mat[a,a] = mat[b,b] = cos, mat[a,b] = -sin, mat[b,a] = sin."""
    mat = np.zeros((4, 4))
    for k in range(4):
        if k not in (a, b):
            mat[k, k] = 1
    mat[a, a] = np.cos(angle)
    mat[a, b] = -np.sin(angle)
    mat[b, a] = np.sin(angle)
    mat[b, b] = np.cos(angle)
    return mat


def transform_matrices(tx, ty, tz, rx=0, ry=0, rz=0):
    """The matmul chain of point_transform as a list of 4x4 float64 matrices."""
    mat1 = np.eye(4)
    mat1[3, 0:3] = tx, ty, tz
    mats = [mat1]
    for angle, (a, b) in ((rx, (1, 2)), (ry, (2, 0)), (rz, (0, 1))):
        if angle != 0:
            mats.append(rotation_matrix(angle, a, b))
    return mats


def random_rotation_points(lidar, angle):
    """Random_Rotation.__call__ on the sweep (transformation.py:349-352): float32 in place."""
    lidar = lidar.copy()
    lidar[:, 0:3] = point_transform(lidar[:, 0:3], 0, 0, 0, rz=angle)
    return lidar


def random_scaling_points(lidar, factor):
    """Random_Scaling.__call__ on the sweep (transformation.py:366-368): float32 times a Python float."""
    lidar = lidar.copy()
    lidar[:, 0:3] = lidar[:, 0:3] * factor
    return lidar


# ----------------------------------------------------------------------------- Argoverse raster
ARGO_BV_BOUNDARY = {"minX": -40.0, "maxX": 40.0, "minY": -40.0, "maxY": 40.0, "minZ": -3.0, "maxZ": 1.0}
ARGO_BV_DISCRETIZATION = 0.1   # argoverse_test.py:37-47
# (discretization, boundary) sets the fixtures and parity tests run on: the scripts' own, a coarse
# asymmetric one (400 x 300) and one whose cell count is not a multiple of 4 (150 x 109 -> scalar kernel paths)
BV_TEST_GEOMS = {
    "argo": (ARGO_BV_DISCRETIZATION, ARGO_BV_BOUNDARY),
    "coarse": (0.2, {"minX": -20.0, "maxX": 60.0, "minY": -30.0, "maxY": 30.0, "minZ": -2.5, "maxZ": 3.5}),
    "odd": (0.3, {"minX": -10.0, "maxX": 35.0, "minY": -12.5, "maxY": 20.2, "minZ": -1.7, "maxZ": 2.2})}


def bv_feature_shape(discretization, boundary):
    """argoverse_test.py:224-225."""
    return (int((boundary["maxX"] - boundary["minX"]) / discretization),
            int((boundary["maxY"] - boundary["minY"]) / discretization))


def makeBVFeature(points, discretization, boundary):
    """argoverse_test.py:199-254 restated without the per-point Python loop: max / max / count are
    commutative, so ufunc.at gives the loop's result.  float32 sweeps [N, 3 or >= 4] ->
    float32 [3, H, W] = (density, height, intensity)."""
    if points.shape[1] == 3:
        x, y, z = points[:, 0], points[:, 1], points[:, 2]
        i = np.ones_like(x) * 0.5
    elif points.shape[1] >= 4:
        x, y, z, i = points[:, 0], points[:, 1], points[:, 2], points[:, 3]
    else:
        raise ValueError(f"Invalid point cloud shape: {points.shape}")
    mask = ((x >= boundary["minX"]) & (x <= boundary["maxX"]) & (y >= boundary["minY"]) & (y <= boundary["maxY"]) &
            (z >= boundary["minZ"]) & (z <= boundary["maxZ"]))
    H, W = bv_feature_shape(discretization, boundary)
    if not np.any(mask):
        return np.zeros((3, H, W), dtype=np.float32)
    x, y, z, i = x[mask], y[mask], z[mask], i[mask]
    x_idx = np.clip(((boundary["maxX"] - x) / discretization).astype(np.int32), 0, H - 1)
    y_idx = np.clip(((y - boundary["minY"]) / discretization).astype(np.int32), 0, W - 1)
    h_map = np.zeros((H, W), dtype=np.float32)
    d_map = np.zeros((H, W), dtype=np.float32)
    i_map = np.zeros((H, W), dtype=np.float32)
    z_rel = z - boundary["minZ"]
    # the loop's `max(map[cell], v)` keeps the map value unless v > it: starting from +0.0, only
    # values > 0 ever enter (no NaN, no -0.0)
    zs, si = z_rel > 0, i > 0
    np.maximum.at(h_map, (x_idx[zs], y_idx[zs]), z_rel[zs])
    np.maximum.at(i_map, (x_idx[si], y_idx[si]), i[si])
    np.add.at(d_map, (x_idx, y_idx), np.float32(1))
    d_map = np.clip(d_map / 10.0, 0, 1)
    with np.errstate(all="ignore"):
        if h_map.max() > 0:
            h_map = h_map / (boundary["maxZ"] - boundary["minZ"])
        if i_map.max() > 0:
            i_map = i_map / i_map.max()
    return np.stack([d_map, h_map, i_map], axis=0).astype(np.float32)


def synth_argoverse_sweep(seed, n=250_000, kind="uniform", boundary=ARGO_BV_BOUNDARY):
    """float32 [n,4] sweeps for makeBVFeature: Argoverse-like (intensity 0..255), spilling 10 % over
    each bound; `adversarial` adds grid-aligned coordinates, points on the bounds, negative / NaN /
    zero intensities, NaN coordinates and a crowded cell."""
    rng = np.random.default_rng(seed)
    b = boundary
    sx, sy, sz = b["maxX"] - b["minX"], b["maxY"] - b["minY"], b["maxZ"] - b["minZ"]
    pts = np.empty((n, 4), dtype=np.float64)
    pts[:, 0] = rng.uniform(b["minX"] - 0.1 * sx, b["maxX"] + 0.1 * sx, n)
    pts[:, 1] = rng.uniform(b["minY"] - 0.1 * sy, b["maxY"] + 0.1 * sy, n)
    pts[:, 2] = rng.uniform(b["minZ"] - 0.1 * sz, b["maxZ"] + 0.1 * sz, n)
    pts[:, 3] = rng.integers(0, 256, n)
    if kind == "adversarial" and n >= 64:
        q = n // 8
        pts[:q, 0] = np.round(pts[:q, 0] / 0.1) * 0.1                    # on cell edges
        pts[:q, 1] = np.round(pts[:q, 1] / 0.1) * 0.1
        pts[q:q + 6, 0] = [b["minX"], b["maxX"], b["minX"], b["maxX"], 0.0, -0.0]
        pts[q:q + 6, 1] = [b["minY"], b["maxY"], b["maxY"], b["minY"], 0.0, -0.0]
        pts[q:q + 6, 2] = [b["minZ"], b["maxZ"], b["minZ"], b["maxZ"], b["minZ"], b["maxZ"]]
        pts[2 * q:2 * q + q // 2, 3] = -pts[2 * q:2 * q + q // 2, 3]     # negative intensity never wins over 0
        pts[3 * q:3 * q + 50, 3] = np.nan
        pts[3 * q + 50:3 * q + 100, 0] = np.nan
        pts[3 * q + 100:3 * q + 150, 2] = np.inf
        pts[4 * q:4 * q + q // 2, 3] = 0.0
        pts[5 * q:5 * q + 40, :2] = [12.34, -7.89]                        # > 10 points in one cell: density saturates
        pts[6 * q:7 * q, 2] = np.round(pts[6 * q:7 * q, 2] * 4) / 4       # z ties
    elif kind == "xyz_only":
        return pts[:, :3].astype(np.float32)
    elif kind == "dark":
        pts[:, 3] = 0.0                                                   # i_map.max() == 0: no normalisation
    elif kind == "outside":
        pts[:, 2] = b["maxZ"] + 5.0                                       # nothing passes the mask
    elif kind == "wide":
        return np.concatenate([pts, rng.uniform(0, 1, (n, 2))], axis=1).astype(np.float32)   # 6 columns
    return pts.astype(np.float32)


# ----------------------------------------------------------------------------- lidar boxes -> camera frame -> image
def lidar_to_camera(x, y, z, V2C, R0):
    """data_process/transformation.py:50-60 with explicit calibration (V2C 3x4, R0 3x3)."""
    p = np.array([x, y, z, 1])
    p = np.matmul(V2C, p)
    p = np.matmul(R0, p)
    return tuple(p[0:3])


def lidar_to_camera_box(boxes, V2C, R0, P2=None):
    """data_process/transformation.py:99-107: (N,7) x,y,z,h,w,l,rz -> x,y,z (rect camera frame),h,w,l,ry."""
    ret = []
    for box in boxes:
        x, y, z, h, w, l, rz = box
        (x, y, z), ry = lidar_to_camera(x, y, z, V2C, R0), -rz - np.pi / 2
        ret.append([x, y, z, h, w, l, ry])
    return np.array(ret).reshape(-1, 7)


def project_box_to_image(box_3d_cam, P2, img_shape):
    """test6.py:147-185 (same body in test4.py:146-184, msac.py:161-199, slam.py:161-199): the 8
    corners of the camera-frame box through P2 -> axis-aligned image box, clipped with Python's
    max()/min() (a NaN bound falls back to the image border).  Returns (min_x, min_y, max_x, max_y)
    as floats and whether the reference would emit the box."""
    x, y, z, h, w, l, ry = box_3d_cam
    corners_3d = np.array([
        [-l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2],
        [0, 0, 0, 0, -h, -h, -h, -h],
        [-w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2]])
    R = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    corners_3d = np.dot(R, corners_3d)
    corners_3d[0, :] += x
    corners_3d[1, :] += y
    corners_3d[2, :] += z
    with np.errstate(all="ignore"):
        corners_2d = P2.dot(np.vstack((corners_3d, np.ones((1, 8)))))
        corners_2d = corners_2d[:2] / corners_2d[2]
    min_x, max_x = np.min(corners_2d[0]), np.max(corners_2d[0])
    min_y, max_y = np.min(corners_2d[1]), np.max(corners_2d[1])
    min_x = max(0, min_x)
    min_y = max(0, min_y)
    max_x = min(img_shape[1], max_x)
    max_y = min(img_shape[0], max_y)
    return (float(min_x), float(min_y), float(max_x), float(max_y)), bool(max_x > min_x and max_y > min_y)


def convert_sfa3d_to_2d_boxes(sfa_detections, V2C, R0, P2, img_shape, min_confidence=0.3, geom: Geometry = KITTI):
    """test6.py:129-187.  Literal behaviour kept: `confidence` is column 0 of the real-value row, which
    convert_det_to_real_values fills with the CLASS ID (evaluation_utils.py:191), so class 0 is always
    skipped and the returned confidences are class ids.  min_confidence is 0.3 in test4/test6 and 0.2
    in msac/slam."""
    boxes_2d, confidences = [], []
    if len(sfa_detections) > 0:
        for detection in convert_det_to_real_values(sfa_detections, geom=geom):
            confidence = detection[0]
            if confidence < min_confidence:
                continue
            cam = lidar_to_camera_box(detection[1:].reshape(1, -1), V2C, R0, P2)[0]
            (min_x, min_y, max_x, max_y), ok = project_box_to_image(cam, P2, img_shape)
            if ok:
                boxes_2d.append([int(min_x), int(min_y), int(max_x - min_x), int(max_y - min_y)])
                confidences.append(confidence)
    return boxes_2d, confidences


def synth_calibration(seed=0):
    """A KITTI-like calibration (float32, as Calibration.read_calib_file parses it,
    data_process/kitti_data_utils.py:149-170): the dataset-average matrices of
    config/kitti_config.py:64-83 with a small seeded perturbation."""
    rng = np.random.default_rng(seed)
    V2C = np.array([[7.49916597e-03, -9.99971248e-01, -8.65110297e-04, -6.71807577e-03],
                    [1.18652889e-02, 9.54520517e-04, -9.99910318e-01, -7.33152811e-02],
                    [9.99882833e-01, 7.49141178e-03, 1.18719929e-02, -2.78557062e-01]])
    R0 = np.array([[0.99992475, 0.00975976, -0.00734152],
                   [-0.0097913, 0.99994262, -0.00430371],
                   [0.00729911, 0.0043753, 0.99996319]])
    P2 = np.array([[719.787081, 0., 608.463003, 44.9538775],
                   [0., 719.787081, 174.545111, 0.1066855],
                   [0., 0., 1., 3.0106472e-03]])
    if seed:
        V2C = V2C + rng.normal(0, 1e-3, V2C.shape)
        R0 = R0 + rng.normal(0, 1e-4, R0.shape)
        P2 = P2 * (1 + rng.normal(0, 1e-2)) * np.array([[1, 0, 1, 1], [0, 1, 1, 1], [0, 0, 1, 1]])
    return V2C.astype(np.float32), R0.astype(np.float32), P2.astype(np.float32)


# ----------------------------------------------------------------------------- tie-insensitive decode comparator
def canonical_detections(det):
    """torch.topk's order among EQUAL scores is implementation-defined (SURVEY.md §7 "top-K ties"),
    so detections are compared after sorting rows by (score desc, cls asc, y asc, x asc ... all cols)."""
    det = np.asarray(det)
    out = np.empty_like(det)
    for b in range(det.shape[0]):
        d = det[b]
        keys = [d[:, c] for c in range(d.shape[1] - 1, 0, -1)] + [-d[:, 0]]
        out[b] = d[np.lexsort(keys)]
    return out


# ----------------------------------------------------------------------------- synthetic inputs
def synth_sweep(seed, n=120_000, geom: Geometry = KITTI, kind="uniform"):
    """Seeded synthetic sweeps.  `uniform` is the timed KITTI-shaped workload; the others are the
    adversarial parity cases of SURVEY.md §8d."""
    rng = np.random.default_rng(seed)
    b = geom.boundary
    sx, sy, sz = b["maxX"] - b["minX"], b["maxY"] - b["minY"], b["maxZ"] - b["minZ"]

    def uni(lo, hi, m=n):
        return rng.uniform(lo, hi, m)

    if kind == "uniform":
        cols = [uni(b["minX"], b["maxX"]), uni(b["minY"], b["maxY"]), uni(b["minZ"], b["maxZ"]), uni(0, 1)]
    elif kind == "outside":        # ~10 % beyond each bound
        cols = [uni(b["minX"] - 0.1 * sx, b["maxX"] + 0.1 * sx), uni(b["minY"] - 0.1 * sy, b["maxY"] + 0.1 * sy),
                uni(b["minZ"] - 0.1 * sz, b["maxZ"] + 0.1 * sz), uni(0, 1)]
    elif kind == "zties":          # z quantised to 0.25 m -> many equal-z points per cell
        z = np.round(uni(b["minZ"], b["maxZ"]) * 4) / 4
        cols = [uni(b["minX"], b["maxX"]), uni(b["minY"], b["maxY"]), z, uni(0, 1)]
    elif kind == "gridaligned":    # coordinates at exact multiples of the cell size (fp32 division edge)
        d32 = np.float32(geom.DISCRETIZATION)
        kx = rng.integers(int(b["minX"] / geom.DISCRETIZATION) - 2, int(b["maxX"] / geom.DISCRETIZATION) + 3, n)
        ky = rng.integers(int(b["minY"] / geom.DISCRETIZATION) - 2, int(b["maxY"] / geom.DISCRETIZATION) + 3, n)
        jit = rng.integers(-1, 2, (2, n))
        x = (kx * d32).astype(np.float32)
        y = (ky * d32).astype(np.float32)
        toward = lambda v, j: np.where(j > 0, np.float32(np.inf), np.where(j < 0, np.float32(-np.inf), v))
        x = np.nextafter(x, toward(x, jit[0]).astype(np.float32))   # one ulp below / at / above the cell edge
        y = np.nextafter(y, toward(y, jit[1]).astype(np.float32))
        cols = [x, y, uni(b["minZ"], b["maxZ"]), uni(0, 1)]
    elif kind == "bounds":         # points exactly on each bound (inclusive filter, cropped last row/col)
        pick = lambda lo, hi: rng.choice(np.array([lo, hi, (lo + hi) / 2], dtype=np.float64), n)
        cols = [pick(b["minX"], b["maxX"]), pick(b["minY"], b["maxY"]), pick(b["minZ"], b["maxZ"]), uni(0, 1)]
    elif kind == "nonfinite":      # NaN / inf rows sprinkled in
        cols = [uni(b["minX"], b["maxX"]), uni(b["minY"], b["maxY"]), uni(b["minZ"], b["maxZ"]), uni(0, 1)]
        bad = rng.integers(0, n, max(1, n // 50))
        which = rng.integers(0, 3, bad.shape[0])
        val = rng.choice(np.array([np.nan, np.inf, -np.inf]), bad.shape[0])
        for c in range(3):
            cols[c][bad[which == c]] = val[which == c]
    elif kind == "onecell":        # everything in one cell: count saturation + long tie chains
        cx = b["minX"] + 0.37 * sx
        cy = b["minY"] + 0.61 * sy
        d = geom.DISCRETIZATION
        z = np.round(uni(b["minZ"], b["maxZ"]) * 8) / 8
        cols = [cx + uni(0, 0.2 * d), cy + uni(0, 0.2 * d), z, uni(0, 1)]
    elif kind == "clustered":      # scan-line-like locality: consecutive points fall in nearby cells
        t = np.linspace(0, 1, n)
        cols = [b["minX"] + sx * (0.5 + 0.49 * np.sin(37 * t) * t), b["minY"] + sy * (0.5 + 0.49 * np.cos(37 * t) * t),
                uni(b["minZ"], b["maxZ"]), uni(0, 1)]
    else:
        raise ValueError(kind)
    return np.stack(cols, axis=1).astype(np.float32)


def synth_heads(seed, B=1, C=3, h=152, w=152, tie_free=False):
    """Heads(seed) of SURVEY.md §8d, in the value range every reference caller feeds decode with:
    hm / cen_offset are post-_sigmoid, i.e. inside [1e-4, 1-1e-4] (test.py:150,167); dir/z/dim raw.
    Only exactly-rounded operations are used (uniform draw, one multiply, one add) so that the same
    seed gives the same bits on every host — torch's CPU sigmoid differs in the last ulp with the
    thread count, which would unpin the committed fixtures.
    tie_free maps a random permutation onto distinct float32 values in (1e-4, 1-1e-4) so that
    top-K indices are uniquely defined (SURVEY.md §8c parity rules)."""
    g = torch.Generator().manual_seed(int(seed))

    def unit(*shape):
        return torch.rand(*shape, generator=g) * (1 - 2e-4) + 1e-4

    if tie_free:
        n = C * h * w
        base = (torch.arange(n, dtype=torch.float64) * ((1 - 3e-4) / max(n - 1, 1)) + 1.5e-4).float()
        assert torch.unique(base).numel() == n
        hm = torch.stack([base[torch.randperm(n, generator=g)] for _ in range(B)]).view(B, C, h, w).contiguous()
    else:
        hm = unit(B, C, h, w)
    off = unit(B, 2, h, w)
    direction = torch.randn(B, 2, h, w, generator=g)
    z = torch.randn(B, 1, h, w, generator=g)
    dim = torch.randn(B, 3, h, w, generator=g)
    return hm, off, direction, z, dim
