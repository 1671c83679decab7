"""Drop-in for the reference's data_process/kitti_bev_utils.py on the hot path:
`makeBEVMap(PointCloud_, boundary)` with the same signature, dtype and error behaviour
(reference: data_process/kitti_bev_utils.py:22-55), computed on the B200 by libsfa_b200.so.
The drawing helpers of that file (get_corners, drawRotatedBox; cv2) are out of scope.

Geometry comes from `cnf` (module global, like the reference); assign `cnf = other_module` to
rasterise another range (the reference is monkey-patched the same way for Argoverse ranges)."""
import threading

import numpy as np

from ..config import kitti_config as cnf  # noqa: F401  (module global on purpose, see above)
from .. import geometry as _geometry

_pipelines = {}
_pipelines_lock = threading.Lock()


def _pipeline(boundary, apply_filter, n_points):
    """One cached HostPipeline per (CUDA device, geometry, filter) — the calling thread's CURRENT device, like any
    torch op — grown when a larger sweep arrives.  The cache is guarded by a lock; calls on one pipeline serialise
    inside the library (SfaPipeline holds a mutex)."""
    import torch
    from ..fast import HostPipeline
    geom = _geometry.BevGeometry(boundary, cnf, apply_filter=apply_filter)
    if not torch.cuda.is_available():
        raise RuntimeError("libsfa_b200 needs a CUDA device (there is no CPU fallback)")
    device = torch.cuda.current_device()
    key = (device,) + tuple(geom.key())
    with _pipelines_lock:
        pl = _pipelines.get(key)
        if pl is None or pl.max_points < n_points:
            if pl is not None:
                pl.close()
            cap = max(131072, 1 << int(np.ceil(np.log2(max(n_points, 1)))))
            pl = HostPipeline(geom, max_frames=1, max_points=cap, C=0, h=1, w=1, K=1, device=device)
            _pipelines[key] = pl
    return pl


def _widen_like_reference(maps32, geom):
    """float32 planes -> the reference's float64 map.  Height and intensity are float32 values the
    reference widens on store; density is a float64 the reference computes directly, recovered
    exactly from the (invertible) float32 table."""
    out = maps32.astype(np.float64)
    cnt = np.searchsorted(geom.lut32, maps32[2])
    out[2] = geom.lut64[cnt]
    return out


def makeBEVMap(PointCloud_, boundary):
    """[N,4] float32 (x, y, z, intensity), already filtered -> float64 [3, BEV_HEIGHT, BEV_WIDTH]
    (ch0 intensity, ch1 height, ch2 density).  Raises IndexError like the reference when a point
    indexes outside the (H+1)x(W+1) map (kitti_bev_utils.py:44)."""
    pts = np.ascontiguousarray(PointCloud_, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 4:
        raise IndexError("PointCloud_ must be [N, >=4]")
    if pts.shape[1] != 4:
        pts = np.ascontiguousarray(pts[:, :4])
    pl = _pipeline(boundary, False, pts.shape[0])
    offsets = np.array([0, pts.shape[0]], dtype=np.int64)
    maps, n_bad = pl.bev(pts, offsets)
    if n_bad:
        raise IndexError("%d point(s) index outside the %dx%d BEV map" % (n_bad, pl.geom.height + 1, pl.geom.width + 1))
    return _widen_like_reference(maps[0], pl.geom)


def makeBEVMap_from_raw(lidar, boundary):
    """get_filtered_lidar + makeBEVMap in one fused pass over the raw sweep (what
    data_process/kitti_dataset.py:64-66 does with two calls); float32 [3,H,W]."""
    pts = np.ascontiguousarray(lidar, dtype=np.float32)
    pl = _pipeline(boundary, True, pts.shape[0])
    maps, n_bad = pl.bev(pts, np.array([0, pts.shape[0]], dtype=np.int64))
    if n_bad:
        raise IndexError("%d point(s) index outside the BEV map" % n_bad)
    return maps[0]
