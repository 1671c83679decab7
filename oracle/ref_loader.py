"""Import the reference's own hot-path modules (and its fpn_resnet_18), unmodified.

TEST INFRASTRUCTURE, like everything under oracle/.  Used by tests/golden/make_golden.py (fixtures), the live
pinning test, the BASELINE config[2] GPU test (the reference's model between our two stages), and
bench.py's `--impl reference` / `cpu_baseline` legs.  The product never imports it.

Where the reference comes from, first hit wins:
  1. $SFA_REFERENCE_ROOT
  2. /root/reference                      (build container only; read in place)
  3. <repo>/baseline/_ref/sfa               (git-ignored copy made by install(), i.e. by
                                           `python oracle/ref_loader.py install` or __graft_entry__.build();
                                           it travels to the GPU box with the snapshot, /root/reference does not)
Reference sources are never committed: baseline/_ref/ (where the bench contract keeps the unmodified reference) is in
.gitignore and not in .gpurunignore.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "sfa")
SOURCE_ROOT = "/root/reference"
_COPIED = ("config", "data_process", "utils", "models", "losses")            # SURVEY.md §9's recipe
_SCRIPTS = ("test6.py", "argoverse_test.py")                                 # functions are compiled out of these


def _has_reference(root):
    return bool(root) and os.path.isfile(os.path.join(root, "data_process", "kitti_bev_utils.py"))


def _find_root():
    for cand in (os.environ.get("SFA_REFERENCE_ROOT"), SOURCE_ROOT, INSTALLED_ROOT):
        if _has_reference(cand):
            return cand
    return os.environ.get("SFA_REFERENCE_ROOT") or SOURCE_ROOT


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def install(force=False) -> bool:
    """Copy the reference's Python packages for this path into baseline/_ref/sfa (a directory whose name ends in
    "sfa", which is what the reference's import preamble looks for).  No-op when /root/reference is absent."""
    import shutil
    if not _has_reference(SOURCE_ROOT):
        return _has_reference(INSTALLED_ROOT)
    if _has_reference(INSTALLED_ROOT) and not force:
        return True
    os.makedirs(INSTALLED_ROOT, exist_ok=True)
    for d in _COPIED:
        dst = os.path.join(INSTALLED_ROOT, d)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(SOURCE_ROOT, d), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in _SCRIPTS:
        shutil.copy(os.path.join(SOURCE_ROOT, f), os.path.join(INSTALLED_ROOT, f))
    return True


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


def create_model(arch="fpn_resnet_18", seed=0):
    """The reference's own network (models/model_utils.py:25-43 -> models/fpn_resnet.py:112-263), random init
    (imagenet_pretrained=False) under torch.manual_seed(seed), eval mode, configured like test.py:58-70.
    models/fpn_resnet.py:20 imports matplotlib.pyplot and never uses it: an empty stub stands in when absent."""
    import types
    import torch
    load()   # puts the reference root on sys.path with the realpath answer its preamble needs
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    with _virtual_sfa_root(), contextlib.redirect_stdout(io.StringIO()):
        model_utils = importlib.import_module("models.model_utils")
        cfg = types.SimpleNamespace(arch=arch, head_conv=64, imagenet_pretrained=False,
                                    heads={"hm_cen": 3, "cen_offset": 2, "direction": 2, "z_coor": 1, "dim": 3})
        torch.manual_seed(seed)
        model = model_utils.create_model(cfg)
    return model.eval()


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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "install":
        ok = install(force="--force" in sys.argv)
        print("reference %s at %s" % ("installed" if ok else "NOT available", INSTALLED_ROOT))
