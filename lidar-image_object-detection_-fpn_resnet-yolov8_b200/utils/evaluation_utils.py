"""Drop-in for the hot-path functions of the reference's utils/evaluation_utils.py:
`_nms` (:21-26), `_topk` (:47-62), `decode` (:77-105), `post_processing` (:112-163) and
`convert_det_to_real_values` (:177-193) — same names, arguments, return types and error behaviour,
computed by libsfa_b200.so on the B200 — plus `convert_sfa3d_to_2d_boxes`, which the reference keeps
inside its fusion scripts (test6.py:129-187).  `draw_predictions` (cv2) is out of scope.

CUDA tensors are processed in place on the current stream and the result stays on the device; CPU
tensors (most reference scripts pin the model to the CPU, test.py:50) go through the library's
host-buffer pipeline and come back as CPU tensors.  There is no CPU compute path.

Tie rule: among EQUAL scores torch.topk's order is implementation-defined; here it is lower class
first, then lower y*w+x."""
import ctypes
import threading

import numpy as np
import torch

from .. import _lib
from ..config import kitti_config as cnf

_host_pipelines = {}
_host_pipelines_lock = threading.Lock()


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _need_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("libsfa_b200 needs a CUDA device (there is no CPU fallback)")


def _raise_like_topk(rc):
    """torch.topk raises RuntimeError('selected index k out of range') when K > h*w
    (evaluation_utils.py:50); keep the exception type."""
    if rc != 0:
        raise RuntimeError("decode/_topk: " + _lib.last_error())


def _nms(heat, kernel=3):
    """heat * (max_pool2d(heat, 3, stride=1, padding=1) == heat)  — evaluation_utils.py:21-26."""
    if kernel != 3:
        raise NotImplementedError("the reference only ever uses kernel=3 (evaluation_utils.py:21,80)")
    lib = _lib.load()
    _need_cuda()
    was_cpu = not heat.is_cuda
    x = _f32c(heat.cuda() if was_cpu else heat)
    out = torch.empty_like(x)
    h, w = x.shape[-2], x.shape[-1]
    planes = x.numel() // (h * w) if h * w else 0
    with torch.cuda.device(x.device):
        _lib.check(lib.sfa_nms(_p(x), planes, h, w, _p(out), _stream(x)))
    return out.cpu() if was_cpu else out


def _topk(scores, K=40):
    """(topk_score, topk_inds, topk_clses, topk_ys, topk_xs), each [B,K] — evaluation_utils.py:47-62."""
    lib = _lib.load()
    _need_cuda()
    was_cpu = not scores.is_cuda
    x = _f32c(scores.cuda() if was_cpu else scores)
    B, C, h, w = x.shape
    dev = x.device
    score = torch.empty((B, K), dtype=torch.float32, device=dev)
    inds = torch.empty((B, K), dtype=torch.int64, device=dev)
    clses = torch.empty((B, K), dtype=torch.int32, device=dev)
    ys = torch.empty((B, K), dtype=torch.float32, device=dev)
    xs = torch.empty((B, K), dtype=torch.float32, device=dev)
    from ..fast import decode_workspace
    ws_ptr, ws_bytes = decode_workspace(dev, B, C, h, w, K)
    with torch.cuda.device(dev):
        _raise_like_topk(lib.sfa_topk(_p(x), B, C, h, w, K, _p(score), _p(inds), _p(clses), _p(ys), _p(xs),
                                      ctypes.c_void_p(ws_ptr), ws_bytes, _stream(x)))
    res = (score, inds, clses, ys, xs)
    return tuple(t.cpu() for t in res) if was_cpu else res


def decode(hm_cen, cen_offset, direction, z_coor, dim, K=40):
    """detections [B,K,10] = (score, x, y, z, dim_h, dim_w, dim_l, dir_im, dir_re, cls) on the input
    device — evaluation_utils.py:77-105.  hm_cen / cen_offset are expected post-_sigmoid."""
    lib = _lib.load()
    _need_cuda()
    B, C, h, w = hm_cen.shape
    if not hm_cen.is_cuda:
        from ..fast import HostPipeline
        from .. import geometry as _geometry
        device = torch.cuda.current_device()   # CPU tensors: the calling thread's current device, like any torch op
        key = (device, C, h, w, K)
        with _host_pipelines_lock:
            pl = _host_pipelines.get(key)
            if pl is None or pl.max_frames < B:
                if pl is not None:
                    pl.close()
                pl = HostPipeline(_geometry.from_config(cnf), max_frames=max(B, 8), max_points=0, C=C, h=h, w=w, K=K,
                                  device=device)
                _host_pipelines[key] = pl
        try:
            det = pl.decode(hm_cen, cen_offset, direction, z_coor, dim)
        except _lib.SfaError as e:
            raise RuntimeError(str(e))
        return torch.from_numpy(det)
    from ..fast import decode_device
    off = _f32c(cen_offset) if cen_offset is not None else None
    return decode_device(_f32c(hm_cen), off, _f32c(direction), _f32c(z_coor), _f32c(dim), K=K)


def get_yaw(direction):
    """evaluation_utils.py:108-109 (host helper kept for signature parity)."""
    return np.arctan2(direction[:, 0:1], direction[:, 1:2])


def post_processing(detections, num_classes=3, down_ratio=4, peak_thresh=0.2):
    """numpy [B,K,10] -> list (one dict per sample) of {class: ndarray[n,8] float32} —
    evaluation_utils.py:112-163 with the per-sample `ret.append` of the pristine
    "evaluation_utils copy.py":112-143 (the live copy appends outside the loop and returns only the
    last sample; it also prints every array — neither is reproduced).  Rows keep their score order."""
    from ..fast import post_process_dense
    _need_cuda()
    det = torch.as_tensor(np.ascontiguousarray(detections, dtype=np.float32))
    if det.shape[0] == 0:
        return []
    rows, cls, keep = post_process_dense(det.cuda(), num_classes, down_ratio, peak_thresh, cnf)
    rows, cls, keep = rows.cpu().numpy(), cls.cpu().numpy(), keep.cpu().numpy()
    ret = []
    for i in range(rows.shape[0]):
        ret.append({j: rows[i][(cls[i] == j) & keep[i]] for j in range(num_classes)})
    return ret


def convert_det_to_real_values(detections, num_classes=3):
    """BEV-pixel boxes -> metric lidar-frame boxes [cls, x, y, z, h, w, l, yaw] (one row per kept
    detection, classes in order) — evaluation_utils.py:177-193.  `detections` is one sample's dict as
    returned by post_processing.  Computed by `sfa_real_values` in float32, which is what numpy's float32
    scalars give step by step; the result is the float64 array the reference builds.  The batched device
    form is fast.post_process_dense(..., real=True)."""
    rows, cls = [], []
    for cls_id in range(num_classes):
        d = np.asarray(detections[cls_id], dtype=np.float32).reshape(-1, 8)
        if d.shape[0]:
            rows.append(d)
            cls.append(np.full(d.shape[0], cls_id, dtype=np.int32))
    if not rows:
        return np.array([])
    _need_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows_d = torch.from_numpy(np.ascontiguousarray(np.concatenate(rows, 0))).to(dev)
    cls_d = torch.from_numpy(np.concatenate(cls)).to(dev)
    real = torch.empty_like(rows_d)
    b = cnf.boundary
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sfa_real_values(_p(rows_d), _p(cls_d), rows_d.shape[0], float(cnf.bound_size_y),
                                               float(cnf.BEV_WIDTH), float(cnf.bound_size_x), float(cnf.BEV_HEIGHT),
                                               float(b["minX"]), float(b["minY"]), float(b["minZ"]), _p(real), _stream(rows_d)))
    return real.cpu().numpy().astype(np.float64)


def convert_sfa3d_to_2d_boxes(sfa_detections, calib, img_shape, min_confidence=0.3):
    """One sample's post_processing dict -> (image boxes [[x, y, w, h], ...], confidences) through the
    calibration object's V2C / R0 / P2 — test6.py:129-187 (test4.py:128-186; msac.py / slam.py:130-201
    pass the matrices in a dict and use min_confidence=0.2, both accepted here).  Literal behaviour
    kept: the value tested against min_confidence and returned as "confidence" is column 0 of the
    real-value row, which is the class id (evaluation_utils.py:191).  The batched device form is
    fast.project_boxes_dense."""
    from .. import fast
    boxes_2d, confidences = [], []
    if len(sfa_detections) == 0:
        return boxes_2d, confidences
    kitti_dets = np.asarray(convert_det_to_real_values(sfa_detections), dtype=np.float64).reshape(-1, 8)
    if kitti_dets.shape[0] == 0:
        return boxes_2d, confidences
    _need_cuda()
    get = (lambda k: calib[k]) if isinstance(calib, dict) else (lambda k: getattr(calib, k))
    dev = torch.device("cuda", torch.cuda.current_device())
    packed = fast.pack_calibration(get("V2C"), get("R0"), get("P2"), device=dev)
    box, valid = fast.project_boxes_dense(torch.from_numpy(kitti_dets[None]).to(dev), packed, img_shape,
                                          min_confidence=min_confidence)
    box, valid = box[0].cpu().numpy(), valid[0].cpu().numpy()
    for i in np.nonzero(valid)[0]:
        boxes_2d.append([int(v) for v in box[i]])
        confidences.append(kitti_dets[i, 0])
    return boxes_2d, confidences
