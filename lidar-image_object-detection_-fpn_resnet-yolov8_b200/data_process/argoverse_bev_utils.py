"""Drop-in for `makeBVFeature` of the reference's Argoverse scripts (argoverse_test.py:199-254, same
body in argoverse_test2.py): float32 sweep [N, 3 or >= 4] -> float32 [3, H, W] = (density, height,
intensity), computed by libsfa_b200.so (`sfa_bvfeature_rasterize`), bit-exact with the reference.
The batched, device-resident form is fast.BvFeatureRasterizer."""
import numpy as np
import torch

from .. import fast

_rasterizers = {}


def makeBVFeature(points, discretization, boundary):
    points = np.asarray(points)
    if points.ndim != 2 or points.shape[1] < 3:
        raise ValueError(f"Invalid point cloud shape: {points.shape}")       # argoverse_test.py:207
    if points.dtype != np.float32:
        # the reference's arithmetic follows the sweep's dtype; the PLY loader (:182-197) yields float32
        raise TypeError("makeBVFeature on the B200 takes float32 sweeps (got %s)" % points.dtype)
    if not torch.cuda.is_available():
        raise RuntimeError("libsfa_b200 needs a CUDA device (there is no CPU fallback)")
    cap = max(1 << 17, 1 << int(points.shape[0] - 1).bit_length()) if points.shape[0] > 1 else 1 << 17
    key = (float(discretization), tuple(sorted(boundary.items())), points.shape[1], cap, torch.cuda.current_device())
    rast = _rasterizers.get(key)
    if rast is None:
        rast = _rasterizers[key] = fast.BvFeatureRasterizer(discretization, boundary, point_floats=points.shape[1],
                                                            max_batch=1, max_points=cap)
    pts = torch.from_numpy(np.ascontiguousarray(points)).to(rast.device)
    return rast(pts[None]).cpu().numpy()[0]
