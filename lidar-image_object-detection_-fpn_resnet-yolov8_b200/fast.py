"""Batched, device-resident API over libsfa_b200.so — the fast path a B200 inference loop uses.

    rast = BevRasterizer(geometry.from_config(kitti_config), max_batch=64)
    bev  = rast(points, offsets, max_points)        # cuda f32 [B,3,608,608]   (stage A)
    det  = decode(hm, off, dir, z, dim, K=50)       # utils.evaluation_utils   (stage B)
    rows, cls, keep = post_process_dense(det)

PyTorch is used for device memory and streams only; all arithmetic happens in the CUDA library.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .geometry import BevGeometry


def _require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("%s must be a CUDA tensor" % name)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class BevRasterizer:
    """Stage A for batches of sweeps resident in HBM (replaces get_filtered_lidar + makeBEVMap,
    data_process/kitti_data_utils.py:228-241 and data_process/kitti_bev_utils.py:22-55)."""

    def __init__(self, geom: BevGeometry, max_batch: int = 64, max_points: int = 131072, device=None, share_with=None):
        """share_with: another BevRasterizer of the same map size, batch and sweep length whose workspace, density
        table and status word this one reuses (calls on the two must be ordered on one stream)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BevRasterizer needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.geom = geom
        self.max_batch = int(max_batch)
        self.max_points = int(max_points)
        nbytes = self.lib.sfa_bev_workspace_bytes(self.max_batch, self.max_points, ctypes.byref(geom.params))
        if nbytes == 0:
            raise _lib.SfaError(-1, _lib.last_error() or "bad geometry")
        if share_with is not None:
            if (share_with._ws_bytes < nbytes or share_with.device != self.device or
                    not np.array_equal(share_with.geom.lut32, geom.lut32) or share_with.geom.algorithm != geom.algorithm or
                    (share_with.geom.height, share_with.geom.width) != (geom.height, geom.width)):
                raise ValueError("share_with must be a rasterizer on the same device with the same map size, algorithm and "
                                 "density table and a workspace at least as large")
            self.workspace, self._ws_ptr, self._ws_bytes = share_with.workspace, share_with._ws_ptr, share_with._ws_bytes
            self.lut, self.status = share_with.lut, share_with.status
            self._owns_ws = False
            return
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_ptr = self.workspace.data_ptr() + off
            self._ws_bytes = nbytes
            self._owns_ws = True
            self.lut = torch.from_numpy(geom.lut32.copy()).to(self.device)
            self.status = torch.zeros(2, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.sfa_bev_workspace_init(ctypes.c_void_p(self._ws_ptr), self._ws_bytes,
                                                       _stream_ptr(self.device)))

    def __del__(self):
        # the library keeps two side streams + events per workspace (chunk-parallel lanes): hand them back with the workspace
        try:
            if getattr(self, "_owns_ws", False) and self.lib is not None:
                self.lib.sfa_bev_workspace_release(ctypes.c_void_p(self._ws_ptr))
        except Exception:
            pass

    def __call__(self, points, offsets, max_points, out=None, mats=None, scales=None, hflip=None, second=None, out_second=None):
        """points [total,4] f32 cuda, offsets [B+1] i64 cuda (or None for a uniform batch of sweeps with
        exactly max_points points each), max_points: python int upper bound on the points of any sweep.
        Returns out [B,3,H,W] f32 (channel 0 intensity, 1 height, 2 density).
        Sweep-side extras, applied to each point while it is in registers (sfa_bev_rasterize_ex; one read of the sweep):
          mats   CUDA float64 [B, m, 4, 4], scales CUDA float32 [B]: the training augmentation in front of the filter
                 (Random_Rotation / Random_Scaling, data_process/transformation.py:349-352, :366-368)
          hflip  CUDA uint8 / bool [B]: mirror that frame's map left-right (torch.flip(bev_map, [-1]), kitti_dataset.py:93-97)
          second a BevGeometry differing in its boundary only, rasterised from the same read into out_second
                 (front + back maps, demo_dataset.py:70-88); the call then returns (out, out_second)."""
        _require_cuda(points, "points")
        if points.dtype != torch.float32 or not points.is_contiguous():
            raise TypeError("points must be contiguous float32")
        if points.device != self.device or (offsets is not None and offsets.device != self.device) or \
                (out is not None and out.device != self.device):
            raise ValueError("points / offsets / out must live on the rasteriser's device %s" % self.device)
        if offsets is None:      # uniform batch: every sweep holds exactly max_points points
            if int(max_points) <= 0 or points.numel() % (4 * int(max_points)):
                raise ValueError("a uniform batch needs total_points to be a multiple of max_points")
            B = points.numel() // (4 * int(max_points))
        else:
            _require_cuda(offsets, "offsets")
            if offsets.dtype != torch.int64 or not offsets.is_contiguous():
                raise TypeError("offsets must be contiguous int64")
            B = offsets.numel() - 1
        g = self.geom
        if int(max_points) > self.max_points:
            raise ValueError("max_points %d exceeds the %d this rasteriser's workspace was sized for"
                             % (int(max_points), self.max_points))
        if out is None:
            out = torch.empty((B, 3, g.height, g.width), dtype=torch.float32, device=self.device)
        elif out.shape != (B, 3, g.height, g.width) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be contiguous float32 [B,3,H,W]")
        if mats is None and scales is None and hflip is None and second is None:
            with torch.cuda.device(self.device):   # (any B: the frames stream through the workspace's ring)
                _lib.check(self.lib.sfa_bev_rasterize(_ptr(points), _ptr(offsets), B, int(max_points),
                                                      ctypes.byref(g.params), _ptr(self.lut), _ptr(out), _ptr(self.status),
                                                      ctypes.c_void_p(self._ws_ptr), self._ws_bytes, _stream_ptr(self.device)))
            return out
        ex = _lib.SfaBevExtras()
        keep = []   # tensors the struct points to
        if mats is not None:
            if mats.dim() == 3:
                mats = mats[None]
            if (not mats.is_cuda or mats.device != self.device or mats.dtype != torch.float64 or mats.dim() != 4 or
                    mats.shape[0] != B or tuple(mats.shape[2:]) != (4, 4) or mats.shape[1] > 4):
                raise ValueError("mats must be a CUDA float64 [B, m <= 4, 4, 4] tensor on %s" % self.device)
            mats = mats.contiguous()
            keep.append(mats)
            ex.mats, ex.n_mats = mats.data_ptr(), int(mats.shape[1])
        if scales is not None:
            if not scales.is_cuda or scales.device != self.device or scales.dtype != torch.float32 or scales.numel() != B:
                raise ValueError("scales must be a CUDA float32 [B] tensor on %s" % self.device)
            scales = scales.contiguous()
            keep.append(scales)
            ex.scales = scales.data_ptr()
        if hflip is not None:
            if not hflip.is_cuda or hflip.device != self.device or hflip.numel() != B:
                raise ValueError("hflip must be a CUDA [B] tensor on %s" % self.device)
            hflip = hflip.to(torch.uint8).contiguous()
            keep.append(hflip)
            ex.hflip = hflip.data_ptr()
        if second is not None:
            if out_second is None:
                out_second = torch.empty((B, 3, g.height, g.width), dtype=torch.float32, device=self.device)
            elif (out_second.shape != (B, 3, g.height, g.width) or out_second.dtype != torch.float32 or
                  not out_second.is_contiguous() or out_second.device != self.device):
                raise ValueError("out_second must be contiguous float32 [B,3,H,W] on %s" % self.device)
            if self.max_batch < 2:
                raise ValueError("two maps per sweep need a workspace sized for max_batch >= 2")
            ex.second = ctypes.pointer(second.params)
            ex.out_second = out_second.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sfa_bev_rasterize_ex(_ptr(points), _ptr(offsets), B, int(max_points), ctypes.byref(g.params),
                                                     ctypes.byref(ex), _ptr(self.lut), _ptr(out), _ptr(self.status),
                                                     ctypes.c_void_p(self._ws_ptr), self._ws_bytes, _stream_ptr(self.device)))
        return out if second is None else (out, out_second)

    def rasterize_uniform(self, points, out=None):
        """points [B,N,4]: every sweep has N points (no offsets array, no dependent load in the kernels)."""
        N = points.shape[1]
        return self(points.reshape(-1, 4), None, N, out=out)

    def out_of_map_points(self, reset=True):
        """Points seen since the last reset whose cell index fell outside the (H+1)x(W+1) map (the
        reference raises IndexError for those).  Synchronises."""
        n = int(self.status[0].item())
        if reset:
            self.status.zero_()
        return n


class FrontBackRasterizer:
    """The front + back pair of the 2-sides demo (data_process/demo_dataset.py:70-88): the same sweeps filtered and
    rasterised with cnf.boundary and with cnf.boundary_back (negative x rows wrap like numpy's negative indices), from
    ONE read of each sweep: every point is mapped with both geometries while it sits in registers and its records go
    to the buckets of the front or the back map (sfa_bev_rasterize_ex).  Returns (front [B,3,H,W], back [B,3,H,W])."""

    def __init__(self, cnf=None, max_batch: int = 64, max_points: int = 131072, device=None):
        from .config import kitti_config
        from . import geometry
        cnf = kitti_config if cnf is None else cnf
        self.front = BevRasterizer(geometry.from_config(cnf, cnf.boundary), max(2, max_batch), max_points, device)
        self.back_geom = geometry.from_config(cnf, cnf.boundary_back)

    def __call__(self, points, offsets, max_points, out=None):
        front_out, back_out = (None, None) if out is None else out
        return self.front(points, offsets, max_points, out=front_out, second=self.back_geom, out_second=back_out)


def filter_lidar_device(points, geom: BevGeometry):
    """get_filtered_lidar (data_process/kitti_data_utils.py:228-241) on a CUDA [N,4] sweep:
    returns the filtered, z-shifted sweep as a new CUDA tensor (one host sync to size it)."""
    lib = _lib.load()
    _require_cuda(points, "points")
    points = points.contiguous()
    n = points.shape[0]
    out = torch.empty_like(points)
    count = torch.zeros(1, dtype=torch.int64, device=points.device)
    ws_bytes = lib.sfa_filter_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=points.device)
    _lib.check(lib.sfa_filter_lidar(_ptr(points), n, ctypes.byref(geom.params), _ptr(out), _ptr(count), _ptr(ws),
                                    ws_bytes, _stream_ptr(points.device)))
    return out[: int(count.item())]


_decode_workspaces = {}


def decode_workspace(device, B, C, h, w, K):
    """(pointer, bytes) of a cached, initialised sfa_decode workspace for up to B frames of [C,h,w]
    heads on `device` (one per device and head shape, grown on demand)."""
    lib = _lib.load()
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), C, h, w)
    ent = _decode_workspaces.get(key)
    if ent is None or ent[2] < B:
        cap = max(B, 2 * ent[2] if ent else B)
        nbytes = lib.sfa_decode_workspace_bytes(cap, C, h, w, K)
        if nbytes == 0:
            raise _lib.SfaError(-1, _lib.last_error())
        with torch.cuda.device(device):
            buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            ptr = buf.data_ptr() + (-buf.data_ptr()) % 256
            _lib.check(lib.sfa_decode_workspace_init(ctypes.c_void_p(ptr), nbytes, _stream_ptr(device)))
        ent = (buf, ptr, cap, nbytes)
        _decode_workspaces[key] = ent
    return ent[1], ent[3]


class DecodeWorkspace:
    """An explicit sfa_decode workspace for up to B frames of [C,h,w] heads.  decode_device uses a
    cached per-(device, shape) workspace by default; calls that may run CONCURRENTLY (different
    streams) must each bring their own."""

    def __init__(self, device, B, C, h, w, K):
        lib = _lib.load()
        self.device, self.shape, self.frames = torch.device(device), (C, h, w), int(B)
        self.nbytes = lib.sfa_decode_workspace_bytes(self.frames, C, h, w, K)
        if self.nbytes == 0:
            raise _lib.SfaError(-1, _lib.last_error())
        with torch.cuda.device(self.device):
            self.buf = torch.empty(self.nbytes + 256, dtype=torch.uint8, device=self.device)
            self.ptr = self.buf.data_ptr() + (-self.buf.data_ptr()) % 256
            _lib.check(lib.sfa_decode_workspace_init(ctypes.c_void_p(self.ptr), self.nbytes, _stream_ptr(self.device)))


def decode_device(hm_cen, cen_offset, direction, z_coor, dim, K=40, out=None, inds=None, workspace=None,
                  apply_sigmoid=False, post=None, post_params=None):
    """decode (utils/evaluation_utils.py:77-105) on CUDA float32 NCHW-contiguous heads, writing into
    `out` [B,K,10] (allocated when None) — no allocation, copy or sync when `out` is given, so the
    call can be captured in a CUDA graph.  `inds` optional int64 [B,K] receives the spatial indices.
    apply_sigmoid=True takes the backbone's RAW hm_cen / cen_offset logits and applies `_sigmoid`
    (utils/torch_utils.py:44-45) while loading them (tolerance-level parity, see include/sfa_b200.h).
    post=(rows [B,K,8] f32, cls [B,K] i32, keep [B,K] u8[, real [B,K,8] f32]): also writes the dense post_processing of
    the detections in the same launches (sfa_decode_post); post_params = dict(num_classes, down_ratio, peak_thresh, cnf)."""
    lib = _lib.load()
    for name, t in (("hm_cen", hm_cen), ("direction", direction), ("z_coor", z_coor), ("dim", dim)):
        _require_cuda(t, name)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("%s must be contiguous float32" % name)
    if cen_offset is not None and (not cen_offset.is_cuda or cen_offset.dtype != torch.float32 or
                                   not cen_offset.is_contiguous()):
        raise TypeError("cen_offset must be a contiguous float32 CUDA tensor or None")
    B, C, h, w = hm_cen.shape
    if out is None:
        out = torch.empty((B, K, 10), dtype=torch.float32, device=hm_cen.device)
    elif out.shape != (B, K, 10) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be contiguous float32 [B,K,10]")
    if workspace is not None:
        if workspace.shape != (C, h, w) or workspace.frames < B:
            raise ValueError("workspace was sized for %d frames of %s heads" % (workspace.frames, workspace.shape))
        ws_ptr, ws_bytes = workspace.ptr, workspace.nbytes
    else:
        ws_ptr, ws_bytes = decode_workspace(hm_cen.device, B, C, h, w, K)
    with torch.cuda.device(hm_cen.device):
        if post is None:
            rc = lib.sfa_decode(_ptr(hm_cen), _ptr(cen_offset), _ptr(direction), _ptr(z_coor), _ptr(dim), B, C, h, w, K,
                                _ptr(out), _ptr(inds), 1 if apply_sigmoid else 0, ctypes.c_void_p(ws_ptr), ws_bytes,
                                _stream_ptr(hm_cen.device))
        else:
            from .config import kitti_config
            pp = dict(post_params or {})
            cnf = pp.get("cnf") or kitti_config
            rows, cls, keep = post[0], post[1], post[2]
            real = post[3] if len(post) > 3 else None
            if (rows.shape != (B, K, 8) or rows.dtype != torch.float32 or cls.shape != (B, K) or cls.dtype != torch.int32 or
                    keep.shape != (B, K) or keep.dtype != torch.uint8 or not (rows.is_contiguous() and cls.is_contiguous() and keep.is_contiguous())):
                raise ValueError("post must be (rows [B,K,8] f32, cls [B,K] i32, keep [B,K] u8[, real [B,K,8] f32]), contiguous")
            bd = cnf.boundary
            rc = lib.sfa_decode_post(_ptr(hm_cen), _ptr(cen_offset), _ptr(direction), _ptr(z_coor), _ptr(dim), B, C, h, w, K,
                                     _ptr(out), _ptr(inds), 1 if apply_sigmoid else 0, int(pp.get("num_classes", 3)),
                                     float(pp.get("down_ratio", 4)), float(cnf.bound_size_y), float(cnf.BEV_WIDTH),
                                     float(cnf.bound_size_x), float(cnf.BEV_HEIGHT), float(pp.get("peak_thresh", 0.2)),
                                     float(bd["minX"]), float(bd["minY"]), float(bd["minZ"]), _ptr(rows), _ptr(cls), _ptr(keep),
                                     _ptr(real), ctypes.c_void_p(ws_ptr), ws_bytes, _stream_ptr(hm_cen.device))
    if rc != 0:
        # torch.topk raises RuntimeError when K > h*w (evaluation_utils.py:50): same exception type
        raise RuntimeError("decode: " + _lib.last_error())
    return out


def post_process_dense(detections, num_classes=3, down_ratio=4, peak_thresh=0.2, cnf=None, out=None, real=None):
    """Dense post_processing (utils/evaluation_utils.py:112-163) on CUDA detections [B,K,10]:
    returns (rows [B,K,8] f32, cls [B,K] i32, keep [B,K] u8/bool), all on the device.  With
    `out=(rows, cls, keep_u8)` nothing is allocated (graph-capturable) and keep stays uint8.
    `real`: optional CUDA float32 [B,K,8] that receives convert_det_to_real_values (:177-193) of every
    row (cls, x, y, z, h, w, l, yaw in metres), or True to allocate it; it is then returned as a 4th item."""
    from .config import kitti_config
    cnf = kitti_config if cnf is None else cnf
    lib = _lib.load()
    _require_cuda(detections, "detections")
    det = detections if (detections.dtype == torch.float32 and detections.is_contiguous()) else detections.contiguous().float()
    B, K = det.shape[0], det.shape[1]
    if out is None:
        rows = torch.empty((B, K, 8), dtype=torch.float32, device=det.device)
        cls = torch.empty((B, K), dtype=torch.int32, device=det.device)
        keep = torch.empty((B, K), dtype=torch.uint8, device=det.device)
    else:
        rows, cls, keep = out
    if real is True:
        real = torch.empty((B, K, 8), dtype=torch.float32, device=det.device)
    b = cnf.boundary
    with torch.cuda.device(det.device):
        _lib.check(lib.sfa_post_process(_ptr(det), B, K, int(num_classes), float(down_ratio), float(cnf.bound_size_y),
                                        float(cnf.BEV_WIDTH), float(cnf.bound_size_x), float(cnf.BEV_HEIGHT),
                                        float(peak_thresh), float(b["minX"]), float(b["minY"]), float(b["minZ"]),
                                        _ptr(rows), _ptr(cls), _ptr(keep), _ptr(real), _stream_ptr(det.device)))
    res = (rows, cls, keep.bool()) if out is None else (rows, cls, keep)
    return res if real is None else res + (real,)


def transform_points_device(points, mats=None, scales=None, offsets=None, max_points=None, out=None,
                            out_dtype=torch.float32):
    """point_transform / Random_Rotation / Random_Scaling on sweeps resident in HBM
    (data_process/transformation.py:242-285, :349-352, :366-368).
      points  CUDA float32/float64 [B, N, S] (uniform batch) or [sum N, S] with int64 offsets [B+1]; S >= 3
      mats    CUDA float64 [B, m, 4, 4] (or [m, 4, 4] for one sweep): applied in order as [x y z 1] @ M
      scales  CUDA float32 [B]: float32 x, y, z times the factor, after the matrices
      out     result tensor (same layout; pass `points` itself for the in-place form) or None
    Returns float32 [.., S] (the sweep as the reference leaves it) or, with out_dtype=torch.float64 and
    no out, float64 [.., 3] (what point_transform returns)."""
    lib = _lib.load()
    _require_cuda(points, "points")
    if points.dtype not in (torch.float32, torch.float64) or not points.is_contiguous():
        raise TypeError("points must be a contiguous float32 / float64 tensor")
    if offsets is None:
        if points.dim() == 2:
            points_b = points[None]
        elif points.dim() == 3:
            points_b = points
        else:
            raise ValueError("points must be [N, S] or [B, N, S]")
        B, max_points = points_b.shape[0], points_b.shape[1]
    else:
        _require_cuda(offsets, "offsets")
        if offsets.dtype != torch.int64 or points.dim() != 2:
            raise TypeError("ragged batches are [sum N, S] points with int64 offsets")
        B = offsets.numel() - 1
        if max_points is None:
            raise ValueError("max_points (an upper bound of the longest sweep) is needed with offsets")
    S = points.shape[-1]
    n_mats = 0
    if mats is not None:
        _require_cuda(mats, "mats")
        if mats.dim() == 3:
            mats = mats[None]
        if mats.dtype != torch.float64 or mats.dim() != 4 or mats.shape[0] != B or tuple(mats.shape[2:]) != (4, 4):
            raise ValueError("mats must be float64 [B, m, 4, 4]")
        mats = mats.contiguous()
        n_mats = mats.shape[1]
    if scales is not None:
        _require_cuda(scales, "scales")
        if scales.dtype != torch.float32 or scales.numel() != B:
            raise ValueError("scales must be float32 [B]")
        scales = scales.contiguous()
    if out is None:
        if out_dtype == torch.float64:
            out = torch.empty(points.shape[:-1] + (3,), dtype=torch.float64, device=points.device)
        else:
            out = torch.empty(points.shape, dtype=torch.float32, device=points.device)
    elif out.shape[:-1] != points.shape[:-1] or not out.is_contiguous() or out.dtype not in (torch.float32, torch.float64):
        raise ValueError("out must be a contiguous float tensor with the points' leading shape")
    with torch.cuda.device(points.device):
        _lib.check(lib.sfa_transform_points(_ptr(points), int(points.dtype == torch.float64), S, _ptr(offsets), B,
                                            int(max_points), _ptr(mats), n_mats, _ptr(scales), _ptr(out),
                                            int(out.dtype == torch.float64), out.shape[-1], _stream_ptr(points.device)))
    return out


def bv_params(discretization, boundary, point_floats=4):
    """SfaBvParams the way makeBVFeature derives its geometry (argoverse_test.py:211-229, :248):
    H, W in Python doubles; bounds, cell size and height range as the float32 values numpy uses when
    the Python scalars meet the float32 sweep."""
    f = np.float32
    H = int((boundary["maxX"] - boundary["minX"]) / discretization)
    W = int((boundary["maxY"] - boundary["minY"]) / discretization)
    return _lib.SfaBvParams(f(boundary["minX"]), f(boundary["maxX"]), f(boundary["minY"]), f(boundary["maxY"]),
                            f(boundary["minZ"]), f(boundary["maxZ"]), f(discretization),
                            f(boundary["maxZ"] - boundary["minZ"]), H, W, int(point_floats))


class BvFeatureRasterizer:
    """makeBVFeature (argoverse_test.py:199-254) for batches of sweeps resident in HBM:
    points float32 [sum N, point_floats] + int64 offsets [B+1] (or None for a uniform batch) ->
    float32 [B, 3, H, W] = (density, height, intensity)."""

    def __init__(self, discretization, boundary, point_floats=4, max_batch=64, max_points=262144, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BvFeatureRasterizer needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.params = bv_params(discretization, boundary, point_floats)
        self.H, self.W, self.point_floats = self.params.height, self.params.width, int(point_floats)
        if self.H <= 0 or self.W <= 0:
            raise ValueError("empty map: %d x %d" % (self.H, self.W))
        self.max_batch, self.max_points = int(max_batch), int(max_points)
        self.ws_bytes = int(self.lib.sfa_bvfeature_workspace_bytes(self.max_batch, self.max_points,
                                                                   ctypes.byref(self.params)))
        if self.ws_bytes == 0:
            raise _lib.SfaError(-1, _lib.last_error())
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def __call__(self, points, offsets=None, max_points=None, out=None):
        _require_cuda(points, "points")
        if points.dtype != torch.float32 or not points.is_contiguous():
            raise TypeError("points must be a contiguous float32 tensor")
        if offsets is None:
            if points.dim() != 3 or points.shape[2] != self.point_floats:
                raise ValueError("a uniform batch is [B, N, %d]" % self.point_floats)
            B, max_points = points.shape[0], points.shape[1]
        else:
            _require_cuda(offsets, "offsets")
            if offsets.dtype != torch.int64:
                raise TypeError("offsets must be int64")
            B = offsets.numel() - 1
            if max_points is None:
                raise ValueError("max_points (an upper bound of the longest sweep) is needed with offsets")
        if B > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (B, self.max_batch))
        if int(max_points) > self.max_points:
            raise ValueError("max_points %d exceeds the %d the workspace was sized for" % (max_points, self.max_points))
        if out is None:
            out = torch.empty((B, 3, self.H, self.W), dtype=torch.float32, device=self.device)
        elif (not out.is_cuda or out.device != self.device or out.dtype != torch.float32 or not out.is_contiguous() or
              tuple(out.shape) != (B, 3, self.H, self.W)):
            raise ValueError("out must be a contiguous float32 CUDA tensor [B,3,%d,%d] on %s" % (self.H, self.W, self.device))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sfa_bvfeature_rasterize(_ptr(points), _ptr(offsets), B, int(max_points),
                                                        ctypes.byref(self.params), _ptr(out), _ptr(self.ws),
                                                        self.ws_bytes, _stream_ptr(self.device)))
        return out


def pack_calibration(V2C, R0, P2, device=None):
    """Calibration matrices (data_process/kitti_data_utils.py:127-139: float32 V2C [3,4], R0 [3,3],
    P2 [3,4]; leading batch dimensions allowed, 4x4 homogeneous forms are cut down) -> float64
    [..., 33] in the layout sfa_project_boxes reads."""
    def cut(m, r, c):
        m = np.asarray(m, dtype=np.float64)
        return m[..., :r, :c].reshape(m.shape[:-2] + (r * c,))
    packed = np.concatenate([cut(V2C, 3, 4), cut(R0, 3, 3), cut(P2, 3, 4)], axis=-1)
    t = torch.from_numpy(np.ascontiguousarray(packed))
    return t if device is None else t.to(device)


def project_boxes_dense(real, calib, img_shape, keep=None, min_confidence=0.3, want_cam=False, want_float=False,
                        out=None):
    """Dense lidar_to_camera_box + convert_sfa3d_to_2d_boxes (data_process/transformation.py:99-107,
    test6.py:129-187) on the device.  real: CUDA [B,K,8] float32/float64 rows as produced by
    post_process_dense(real=True); calib: pack_calibration(...) on the device, [33] or [B,33];
    img_shape = (height, width).  Returns (box [B,K,4] i32 = x, y, w, h; valid [B,K] bool) and, if
    requested, cam [B,K,7] f64 and the clipped float bounds [B,K,4] f64.  With
    `out=(box, valid_u8)` nothing is allocated."""
    lib = _lib.load()
    _require_cuda(real, "real")
    if real.dtype not in (torch.float32, torch.float64):
        real = real.float()
    real = real.contiguous()
    if real.dim() != 3 or real.shape[2] != 8:
        raise ValueError("real must be [B, K, 8]")
    B, K = real.shape[0], real.shape[1]
    _require_cuda(calib, "calib")
    if calib.dtype != torch.float64 or calib.shape[-1] != 33 or not calib.is_contiguous():
        raise ValueError("calib must be a contiguous float64 [33] or [B,33] tensor (see pack_calibration)")
    per_frame = calib.dim() == 2
    if per_frame and calib.shape[0] != B:
        raise ValueError("calib has %d rows for %d frames" % (calib.shape[0], B))
    dev = real.device
    if out is None:
        box = torch.empty((B, K, 4), dtype=torch.int32, device=dev)
        valid = torch.empty((B, K), dtype=torch.uint8, device=dev)
    else:
        box, valid = out
    cam = torch.empty((B, K, 7), dtype=torch.float64, device=dev) if want_cam else None
    box_f = torch.empty((B, K, 4), dtype=torch.float64, device=dev) if want_float else None
    if keep is not None:
        keep = keep.to(torch.uint8).contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_project_boxes(_ptr(real), int(real.dtype == torch.float64), _ptr(keep), B, K, _ptr(calib),
                                         int(per_frame), int(img_shape[0]), int(img_shape[1]), float(min_confidence),
                                         _ptr(cam), _ptr(box_f), _ptr(box), _ptr(valid), _stream_ptr(dev)))
    res = (box, valid.bool() if out is None else valid)
    if want_cam:
        res += (cam,)
    if want_float:
        res += (box_f,)
    return res


class HostPipeline:
    """Host-buffer entry points (sfa_pipeline_*): numpy / CPU tensors in, numpy out, with chunked,
    overlapped H2D -> kernels -> D2H inside the library.  This is what the drop-in makeBEVMap /
    decode wrappers and bench.py's `e2e` leg call."""

    def __init__(self, geom: BevGeometry, max_frames=64, max_points=131072, C=3, h=152, w=152, K=50, device=None):
        """device: CUDA device index (default: the calling thread's current device, like every torch op)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (there is no CPU fallback)")
        if device is None:
            device = torch.cuda.current_device()
        elif not isinstance(device, int):
            device = torch.device(device).index
            device = torch.cuda.current_device() if device is None else device
        self.device = int(device)
        self.geom, self.max_frames, self.max_points = geom, int(max_frames), int(max_points)
        self.C, self.h, self.w, self.K = C, h, w, K
        lut = np.ascontiguousarray(geom.lut32)
        self._h = self.lib.sfa_pipeline_create(self.device, self.max_frames, self.max_points, ctypes.byref(geom.params),
                                               lut.ctypes.data_as(ctypes.c_void_p), C, h, w, K)
        if not self._h:
            raise _lib.SfaError(-3, _lib.last_error())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sfa_pipeline_destroy(ctypes.c_void_p(self._h))
            self._h = None

    __del__ = close

    def bev(self, points, offsets, out=None):
        """points [total,4] f32 (numpy or CPU tensor, ideally pinned), offsets [B+1] i64 -> [B,3,H,W] f32
        numpy (or writes into `out`).  Returns (out, n_out_of_map)."""
        pts = _as_host(points, np.float32)
        offs = _as_host(offsets, np.int64)
        B = offs.shape[0] - 1
        g = self.geom
        if pts.ndim != 2 or pts.shape[1] != 4:
            raise ValueError("points must be [total, 4] float32")
        if B < 0 or (B > 0 and (offs[0] < 0 or np.any(np.diff(offs) < 0) or offs[-1] > pts.shape[0])):
            raise ValueError("offsets must be non-decreasing, start at >= 0 and end within the %d points given" % pts.shape[0])
        if out is None:
            out = np.empty((B, 3, g.height, g.width), dtype=np.float32)
        o = _as_out(out, (B, 3, g.height, g.width), "out")
        status = np.zeros(2, dtype=np.uint32)
        _lib.check(self.lib.sfa_pipeline_bev_host(ctypes.c_void_p(self._h), pts.ctypes.data_as(ctypes.c_void_p),
                                                  offs.ctypes.data_as(ctypes.c_void_p), B,
                                                  o.ctypes.data_as(ctypes.c_void_p),
                                                  status.ctypes.data_as(ctypes.c_void_p)))
        return out, int(status[0])

    def decode(self, hm, cen_offset, direction, z_coor, dim, out=None):
        """CPU heads [B,c,h,w] f32 -> detections [B,K,10] f32 numpy."""
        hm_, dir_, z_, dim_ = (_as_host(t, np.float32) for t in (hm, direction, z_coor, dim))
        off_ = _as_host(cen_offset, np.float32) if cen_offset is not None else None
        B = hm_.shape[0]
        if hm_.shape[1:] != (self.C, self.h, self.w):
            raise ValueError("pipeline was created for heads %s, got %s" % ((self.C, self.h, self.w), hm_.shape[1:]))
        for name, t, ch in (("cen_offset", off_, 2), ("direction", dir_, 2), ("z_coor", z_, 1), ("dim", dim_, 3)):
            if t is not None and t.shape != (B, ch, self.h, self.w):
                raise ValueError("%s must be [%d,%d,%d,%d], got %s" % (name, B, ch, self.h, self.w, t.shape))
        if out is None:
            out = np.empty((B, self.K, 10), dtype=np.float32)
        o = _as_out(out, (B, self.K, 10), "out")
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)
        _lib.check(self.lib.sfa_pipeline_decode_host(ctypes.c_void_p(self._h), vp(hm_), vp(off_), vp(dir_), vp(z_),
                                                     vp(dim_), B, vp(o)))
        return out


def _as_out(a, shape, name):
    """A caller-supplied output buffer is written IN PLACE by the C side: it must already be what the C side expects
    (a silently made copy would swallow the results; a smaller buffer would be overrun)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise TypeError("%s must be a host tensor" % name)
        a = a.detach().numpy()
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"] or tuple(a.shape) != tuple(shape):
        raise ValueError("%s must be a C-contiguous float32 array of shape %s" % (name, tuple(shape)))
    return a


def _as_host(a, dtype):
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise TypeError("expected a host tensor")
        a = a.detach().numpy()
    a = np.asarray(a)
    if a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=dtype)
    return a
