"""CPU: the C-ABI library builds, loads and exports every symbol include/sfa_b200.h declares; the
host-side mirror (geometry rounding, module layout, error behaviour without a GPU) and the frame
sharding logic incl. a world_size-2 gloo run.  No compute call is made here (no GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import PKG, ROOT, pkg


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sfa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SFA_API\s+[\w\s\*]+?\b(sfa_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    names = header_symbols()
    assert len(names) >= 14 and "sfa_bev_rasterize" in names and "sfa_decode" in names
    lib = ctypes.CDLL(pkg("_lib").LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libsfa_b200.so does not export %s" % n
    # and the ctypes prototype table covers exactly the header
    assert sorted(pkg("_lib").PROTOTYPES) == names


def test_library_exports_nothing_else(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", pkg("_lib").LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert [e for e in exported if e.startswith("sfa_")] == header_symbols()
    assert not [e for e in exported if not e.startswith("sfa_") and not e.startswith("_")], exported


def test_version_and_host_only_entry_points(built_lib):
    lib = built_lib
    assert lib.sfa_version() == 100
    g = pkg("geometry").from_config(pkg("config.kitti_config"))
    assert ctypes.sizeof(g.params) == 52          # struct SfaBevParams: 9 floats + 4 int32
    n1 = lib.sfa_bev_workspace_bytes(1, 120000, ctypes.byref(g.params))
    n64 = lib.sfa_bev_workspace_bytes(64, 120000, ctypes.byref(g.params))
    # tiled: per ring frame 128 buckets of 8x the even share (7504 records) + one overflow list of max_points records
    assert n1 >= (128 * 7504 + 120000) * 16 and n64 >= n1 and n64 < (1 << 30)
    ga = pkg("geometry").from_config(pkg("config.kitti_config"), algorithm=pkg("_lib").BEV_GLOBAL_ATOMIC)
    na = lib.sfa_bev_workspace_bytes(64, 120000, ctypes.byref(ga.params))
    assert 608 * 608 * 12 <= na < (1 << 30)
    bad = pkg("_lib").SfaBevParams()
    assert lib.sfa_bev_workspace_bytes(1, 1000, ctypes.byref(bad)) == 0
    assert b"BEV size" in lib.sfa_last_error()
    assert lib.sfa_filter_workspace_bytes(120000) >= 4
    # argument validation happens before any CUDA call
    rc = lib.sfa_decode(None, None, None, None, None, 1, 3, 152, 152, 50, None, None, 0, None, 0, None)
    assert rc == -1 and b"NULL" in lib.sfa_last_error()
    rc = lib.sfa_bev_rasterize(None, None, -1, 0, ctypes.byref(g.params), None, None, None, None, 0, None)
    assert rc == -1


@pytest.mark.parametrize("H,W", [(608, 608), (304, 304), (1000, 1000), (1700, 1700), (100, 36), (800, 800), (64, 64)])
def test_band_plan_divides_exactly(built_lib, H, W):
    """The tiled BEV kernels map a cell to its band with a multiply-shift; the constants the host
    derives must reproduce integer division for every cell of the map."""
    class Cnf:
        BEV_HEIGHT, BEV_WIDTH, DISCRETIZATION = H, W, 50 / H
    g = pkg("geometry").BevGeometry(O.KITTI.boundary, Cnf)
    nb, cpb, shift = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    magic = ctypes.c_uint32()
    rc = built_lib.sfa_bev_band_plan(ctypes.byref(g.params), ctypes.byref(nb), ctypes.byref(cpb), ctypes.byref(magic),
                                     ctypes.byref(shift))
    assert rc == 1
    cells = np.arange(H * W, dtype=np.uint64)
    assert cpb.value % 4 == 0 and cpb.value <= 2944 and nb.value * cpb.value >= H * W > (nb.value - 1) * cpb.value
    q = ((cells * np.uint64(magic.value)) >> np.uint64(32)) >> np.uint64(shift.value)
    assert np.array_equal(q, cells // np.uint64(cpb.value))
    if (H, W) == (608, 608):
        assert (nb.value, cpb.value) == (128, 2888)


def test_band_plan_falls_back_to_global_atomic(built_lib):
    class Cnf:
        BEV_HEIGHT, BEV_WIDTH, DISCRETIZATION = 301, 301, 50 / 301
    g = pkg("geometry").BevGeometry(O.KITTI.boundary, Cnf)
    assert built_lib.sfa_bev_band_plan(ctypes.byref(g.params), None, None, None, None) == 0   # H*W % 4 != 0
    Cnf.BEV_HEIGHT = Cnf.BEV_WIDTH = 2400
    g = pkg("geometry").BevGeometry(O.KITTI.boundary, Cnf)
    assert built_lib.sfa_bev_band_plan(ctypes.byref(g.params), None, None, None, None) == 0   # more cells than 1024 bands hold
    g = pkg("geometry").BevGeometry(O.KITTI.boundary, Cnf, algorithm=pkg("_lib").BEV_TILED)
    assert built_lib.sfa_bev_band_plan(ctypes.byref(g.params), None, None, None, None) == -1  # cannot be forced


def test_geometry_float32_rounding_matches_numpy_promotion():
    cnf = pkg("config.kitti_config")
    g = pkg("geometry").from_config(cnf)
    p = g.params
    assert p.discretization == np.float32(50 / 608) and p.y_offset == np.float32(304.5)
    assert p.min_z == np.float32(-2.73) and p.max_z == np.float32(1.27) and p.max_height == np.float32(4.0)
    assert (p.height, p.width, p.apply_filter) == (608, 608, 1)
    assert cnf.boundary == O.KITTI.boundary and cnf.boundary_back == O.KITTI_BACK.boundary
    assert cnf.DISCRETIZATION == O.KITTI.DISCRETIZATION and cnf.bound_size_x == 50 and cnf.bound_size_y == 50
    argo = pkg("config.argoverse_config")
    assert argo.boundary == O.ARGOVERSE.boundary and argo.DISCRETIZATION == O.ARGOVERSE.DISCRETIZATION
    assert np.array_equal(g.lut64, O.density_lut64())


def test_mirror_keeps_reference_names():
    ev = pkg("utils.evaluation_utils")
    for name in ("_nms", "_topk", "decode", "get_yaw", "post_processing", "convert_det_to_real_values"):
        assert callable(getattr(ev, name))
    import inspect
    assert list(inspect.signature(ev.decode).parameters) == ["hm_cen", "cen_offset", "direction", "z_coor", "dim", "K"]
    assert inspect.signature(ev.decode).parameters["K"].default == 40
    assert list(inspect.signature(ev.post_processing).parameters) == ["detections", "num_classes", "down_ratio", "peak_thresh"]
    assert list(inspect.signature(pkg("data_process.kitti_bev_utils").makeBEVMap).parameters) == ["PointCloud_", "boundary"]
    assert list(inspect.signature(pkg("data_process.kitti_data_utils").get_filtered_lidar).parameters) == ["lidar", "boundary", "labels"]
    assert callable(pkg("utils.torch_utils")._sigmoid)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_product_raises_without_gpu(built_lib):
    heads = O.synth_heads(0, B=1, h=8, w=8)
    with pytest.raises(RuntimeError):
        pkg("utils.evaluation_utils").decode(*heads, K=5)
    with pytest.raises(RuntimeError):
        pkg("data_process.kitti_bev_utils").makeBEVMap(O.synth_sweep(0, 100), O.KITTI.boundary)
    with pytest.raises(RuntimeError):
        pkg("data_process.kitti_data_utils").get_filtered_lidar(O.synth_sweep(0, 100), O.KITTI.boundary)
    with pytest.raises(RuntimeError):
        pkg("fast").BevRasterizer(pkg("geometry").from_config(pkg("config.kitti_config")))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "sfa_oracle" not in text and "oracle/" not in text and "import oracle" not in text, f


def test_shard_range_partitions():
    sh = pkg("sharding")
    for n in (0, 1, 7, 64, 8192):
        for world in (1, 2, 3, 8):
            for mode in ("block", "cyclic"):
                seen = sorted(i for r in range(world) for i in sh.shard_range(n, r, world, mode))
                assert seen == list(range(n))
                sizes = [len(sh.shard_range(n, r, world, mode)) for r in range(world)]
                assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


_GLOO_WORKER = r'''
import os, sys, importlib
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
sh = importlib.import_module({pkg!r} + ".sharding")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
for mode in ("block", "cyclic"):
    n = 11
    idx = list(sh.shard_range(n, rank, 2, mode))
    local = torch.stack([torch.full((5, 10), float(i)) for i in idx]) if idx else torch.zeros(0, 5, 10)
    out = sh.gather_detections(local, n, mode)
    assert out.shape == (n, 5, 10)
    assert torch.equal(out[:, 0, 0], torch.arange(n, dtype=torch.float32)), (mode, out[:, 0, 0])
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_detections_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, pkg=PKG, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
