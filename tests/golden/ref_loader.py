"""Import the reference's hot-path modules IN PLACE from /root/reference (container only).

TEST INFRASTRUCTURE. Used only by tests/golden/make_golden.py to generate the committed
fixtures and by the (skipped-when-absent) live pinning test. Nothing here is shipped, and
nothing on the GPU box reads /root/reference.

Every reference module runs `while not src_dir.endswith("sfa")` at import
(data_process/kitti_bev_utils.py:13-15, utils/evaluation_utils.py:11-13,
data_process/kitti_data_utils.py:8-10), which spins forever unless some ancestor directory
name ends in "sfa".  Instead of copying sources into a directory called `sfa`, we answer
`os.path.realpath` with a virtual path under ".../sfa/" while the modules are imported.
"""
import contextlib
import importlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("SFA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "data_process", "kitti_bev_utils.py"))


@contextlib.contextmanager
def _virtual_sfa_root():
    real = os.path.realpath

    def fake(p, *a, **k):
        r = real(p, *a, **k)
        if r.startswith(REFERENCE_ROOT + os.sep):
            return os.path.join("/virtual/sfa", os.path.relpath(r, REFERENCE_ROOT))
        return r

    os.path.realpath = fake
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield
    finally:
        os.path.realpath = real


_cache = {}


def _script_function(script, name, env):
    """The reference keeps some steps inside its top-level scripts (test6.py imports ultralytics,
    open3d ... at module level, which are absent here): compile just the named function's own,
    unmodified source out of the script and bind the reference functions it calls."""
    import ast
    import warnings
    path = os.path.join(REFERENCE_ROOT, script)
    with open(path) as f, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        tree = ast.parse(f.read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    env = dict(env)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), env)
    return env[name]


def load():
    """Returns a namespace with the reference's own functions (unmodified code objects)."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    # our package mirrors the reference's module names (config / data_process / utils) only
    # INSIDE the package, never as top-level modules, so these imports resolve to the reference.
    for name in ("config", "data_process", "utils"):
        if name in sys.modules and REFERENCE_ROOT not in (getattr(sys.modules[name], "__file__", "") or ""):
            raise RuntimeError("top-level module %r already imported from elsewhere" % name)
    with _virtual_sfa_root():
        cnf = importlib.import_module("config.kitti_config")
        bev = importlib.import_module("data_process.kitti_bev_utils")
        dat = importlib.import_module("data_process.kitti_data_utils")
        evl = importlib.import_module("utils.evaluation_utils")
        tch = importlib.import_module("utils.torch_utils")
        trf = importlib.import_module("data_process.transformation")
        spec = importlib.util.spec_from_file_location(
            "ref_evaluation_utils_pristine", os.path.join(REFERENCE_ROOT, "utils", "evaluation_utils copy.py"))
        pristine = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pristine)
        spec = importlib.util.spec_from_file_location(
            "ref_argoverse_config", os.path.join(REFERENCE_ROOT, "config", "argoverse_config.py"))
        argo = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(argo)

    class NS:
        pass

    ns = NS()
    ns.cnf, ns.bev, ns.dat, ns.evl, ns.tch, ns.pristine, ns.argo = cnf, bev, dat, evl, tch, pristine, argo
    ns.makeBEVMap = bev.makeBEVMap
    ns.get_filtered_lidar = dat.get_filtered_lidar
    ns._nms, ns._topk, ns.decode = evl._nms, evl._topk, evl.decode
    ns._sigmoid = tch._sigmoid

    def post_processing_live(*a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            return evl.post_processing(*a, **k)

    def post_processing_pristine(*a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            return pristine.post_processing(*a, **k)

    ns.post_processing_live = post_processing_live
    ns.post_processing_pristine = post_processing_pristine
    ns.convert_det_to_real_values = evl.convert_det_to_real_values
    ns.lidar_to_camera_box = trf.lidar_to_camera_box
    ns.point_transform = trf.point_transform
    ns.Random_Rotation, ns.Random_Scaling = trf.Random_Rotation, trf.Random_Scaling
    ns.convert_sfa3d_to_2d_boxes = _script_function(
        "test6.py", "convert_sfa3d_to_2d_boxes",
        {"np": __import__("numpy"), "convert_det_to_real_values": evl.convert_det_to_real_values,
         "lidar_to_camera_box": trf.lidar_to_camera_box})
    ns.makeBVFeature_raw = _script_function("argoverse_test.py", "makeBVFeature", {"np": __import__("numpy")})

    def makeBVFeature(*a, **k):
        import numpy as np
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            return ns.makeBVFeature_raw(*a, **k)

    ns.makeBVFeature = makeBVFeature
    _cache["ns"] = ns
    return ns


@contextlib.contextmanager
def patched_geometry(ns, boundary, bev_h, bev_w, discretization):
    """Monkey-patch the reference's module-global geometry (config/kitti_config.py:23-47) — the
    only way to run its makeBEVMap on another range (Argoverse-range config)."""
    cnf = ns.cnf
    saved = (cnf.boundary, cnf.BEV_HEIGHT, cnf.BEV_WIDTH, cnf.DISCRETIZATION,
             cnf.bound_size_x, cnf.bound_size_y, cnf.bound_size_z)
    try:
        cnf.boundary = dict(boundary)
        cnf.BEV_HEIGHT, cnf.BEV_WIDTH, cnf.DISCRETIZATION = bev_h, bev_w, discretization
        cnf.bound_size_x = boundary["maxX"] - boundary["minX"]
        cnf.bound_size_y = boundary["maxY"] - boundary["minY"]
        cnf.bound_size_z = boundary["maxZ"] - boundary["minZ"]
        yield
    finally:
        (cnf.boundary, cnf.BEV_HEIGHT, cnf.BEV_WIDTH, cnf.DISCRETIZATION,
         cnf.bound_size_x, cnf.bound_size_y, cnf.bound_size_z) = saved
