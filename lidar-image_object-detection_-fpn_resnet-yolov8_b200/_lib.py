"""ctypes binding of libsfa_b200.so (include/sfa_b200.h).  There is NO CPU fallback: if the CUDA
library cannot be loaded every entry point of this package raises."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfa_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_p = ctypes.c_void_p
i32, i64, f32, sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t


class SfaBevParams(ctypes.Structure):
    """struct SfaBevParams of include/sfa_b200.h."""
    _fields_ = [("min_x", f32), ("max_x", f32), ("min_y", f32), ("max_y", f32), ("min_z", f32), ("max_z", f32),
                ("discretization", f32), ("y_offset", f32), ("max_height", f32),
                ("height", i32), ("width", i32), ("apply_filter", i32), ("algorithm", i32)]


BEV_AUTO, BEV_TILED, BEV_GLOBAL_ATOMIC, BEV_TILED_TWO_KERNEL = 0, 1, 2, 3   # enum SfaBevAlgorithm


class SfaBevExtras(ctypes.Structure):
    """struct SfaBevExtras of include/sfa_b200.h (sweep-side extras of sfa_bev_rasterize_ex)."""
    _fields_ = [("mats", c_void_p), ("n_mats", i32), ("scales", c_void_p), ("hflip", c_void_p),
                ("second", ctypes.POINTER(SfaBevParams)), ("out_second", c_void_p)]


class SfaBvParams(ctypes.Structure):
    """struct SfaBvParams of include/sfa_b200.h (makeBVFeature geometry)."""
    _fields_ = [("min_x", f32), ("max_x", f32), ("min_y", f32), ("max_y", f32), ("min_z", f32), ("max_z", f32),
                ("discretization", f32), ("height_range", f32), ("height", i32), ("width", i32), ("point_floats", i32)]


class SfaKernelStat(ctypes.Structure):
    """struct SfaKernelStat of include/sfa_b200.h."""
    _fields_ = [("name", ctypes.c_char * 40), ("launches", ctypes.c_uint64), ("total_ms", ctypes.c_double)]


class SfaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsfa_b200: %s (status %d)" % (msg, code))
        self.code = code


# every symbol include/sfa_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "sfa_version": (ctypes.c_int, []),
    "sfa_last_error": (ctypes.c_char_p, []),
    "sfa_kernel_launches": (ctypes.c_uint64, []),
    "sfa_profile_begin": (ctypes.c_int, []),
    "sfa_profile_end": (ctypes.c_int, [ctypes.POINTER(SfaKernelStat), i32]),
    "sfa_bev_workspace_bytes": (sz, [i32, i64, ctypes.POINTER(SfaBevParams)]),
    "sfa_bev_workspace_init": (ctypes.c_int, [c_void_p, sz, c_void_p]),
    "sfa_bev_workspace_release": (ctypes.c_int, [c_void_p]),
    "sfa_bev_set_internal_lanes": (ctypes.c_int, [i32]),
    "sfa_bev_band_plan": (ctypes.c_int, [ctypes.POINTER(SfaBevParams), ctypes.POINTER(i32), ctypes.POINTER(i32),
                                         ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(i32)]),
    "sfa_bev_rasterize": (ctypes.c_int, [c_void_p, c_void_p, i32, i64, ctypes.POINTER(SfaBevParams), c_void_p,
                                         c_void_p, c_void_p, c_void_p, sz, c_void_p]),
    "sfa_bev_rasterize_ex": (ctypes.c_int, [c_void_p, c_void_p, i32, i64, ctypes.POINTER(SfaBevParams),
                                            ctypes.POINTER(SfaBevExtras), c_void_p, c_void_p, c_void_p, c_void_p, sz, c_void_p]),
    "sfa_selftest_division": (ctypes.c_int, [f32, ctypes.c_uint32, ctypes.c_uint64, c_void_p, c_void_p]),
    "sfa_filter_workspace_bytes": (sz, [i64]),
    "sfa_filter_lidar": (ctypes.c_int, [c_void_p, i64, ctypes.POINTER(SfaBevParams), c_void_p, c_void_p, c_void_p,
                                        sz, c_void_p]),
    "sfa_nms": (ctypes.c_int, [c_void_p, i32, i32, i32, c_void_p, c_void_p]),
    "sfa_decode_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
    "sfa_decode_workspace_init": (ctypes.c_int, [c_void_p, sz, c_void_p]),
    "sfa_topk": (ctypes.c_int, [c_void_p, i32, i32, i32, i32, i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, sz, c_void_p]),
    "sfa_decode": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, i32, i32, i32, i32, i32, c_void_p,
                                  c_void_p, i32, c_void_p, sz, c_void_p]),
    "sfa_decode_post": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, i32, i32, i32, i32, i32, c_void_p,
                                       c_void_p, i32, i32, f32, f32, f32, f32, f32, f32, f32, f32, f32, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, sz, c_void_p]),
    "sfa_post_process": (ctypes.c_int, [c_void_p, i32, i32, i32, f32, f32, f32, f32, f32, f32, f32, f32, f32, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_transform_points": (ctypes.c_int, [c_void_p, i32, i32, c_void_p, i32, i64, c_void_p, i32, c_void_p, c_void_p, i32,
                                            i32, c_void_p]),
    "sfa_bvfeature_workspace_bytes": (sz, [i32, i64, ctypes.POINTER(SfaBvParams)]),
    "sfa_bvfeature_rasterize": (ctypes.c_int, [c_void_p, c_void_p, i32, i64, ctypes.POINTER(SfaBvParams), c_void_p,
                                               c_void_p, sz, c_void_p]),
    "sfa_real_values": (ctypes.c_int, [c_void_p, c_void_p, i32, f32, f32, f32, f32, f32, f32, f32, c_void_p, c_void_p]),
    "sfa_project_boxes": (ctypes.c_int, [c_void_p, i32, c_void_p, i32, i32, c_void_p, i32, i32, i32, ctypes.c_double,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_pipeline_create": (c_void_p, [i32, i32, i64, ctypes.POINTER(SfaBevParams), c_void_p, i32, i32, i32, i32]),
    "sfa_pipeline_destroy": (None, [c_void_p]),
    "sfa_pipeline_bev_host": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, i32, c_void_p, c_void_p]),
    "sfa_pipeline_decode_host": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, i32,
                                                c_void_p]),
}

_lock = threading.Lock()
_lib = None


def load():
    """Returns the loaded library; raises (loudly) when it is missing — run build.py first."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libsfa_b200.so is not built (%s). Run `python __graft_entry__.py build`; this package has no "
                    "CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)  # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise SfaError(rc, (load().sfa_last_error() or b"").decode("utf-8", "replace"))


def last_error():
    return (load().sfa_last_error() or b"").decode("utf-8", "replace")


def kernel_launches():
    """Kernels libsfa_b200.so has launched in this process so far."""
    return int(load().sfa_kernel_launches())


class profile:
    """with _lib.profile() as p: ...   ->   p.stats = {kernel name: (launches, total_ms)} (device time
    from CUDA events the library records around each of its launches)."""

    def __enter__(self):
        check(load().sfa_profile_begin())
        self.stats = {}
        return self

    def __exit__(self, *exc):
        buf = (SfaKernelStat * 32)()
        n = load().sfa_profile_end(buf, 32)
        if n < 0:
            check(n)
        self.stats = {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n)}
        return False
