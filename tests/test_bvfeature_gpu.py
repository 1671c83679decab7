"""GPU: makeBVFeature (sfa_bvfeature_rasterize, through the C ABI) bit-exact against the reference
fixtures (tests/golden/bvfeature_small.npz, bvfeature_hashes.json) and the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def mods(cuda_device):
    return pkg("fast"), pkg("data_process.argoverse_bev_utils")


def dense(z, tag):
    shape = tuple(int(v) for v in z[tag + "_shape"])
    flat = np.zeros(int(np.prod(shape)), dtype=np.uint32)
    flat[z[tag + "_nz"]] = z[tag + "_val"]
    return flat.view(np.float32).reshape(shape)


def test_fixture_cases_through_the_mirror(mods):
    fast, mirror = mods
    z = np.load(os.path.join(GOLD, "bvfeature_small.npz"))
    for c in range(int(z["n_cases"])):
        tag = "c%02d" % c
        gname, kind = [str(v) for v in z[tag + "_meta"]]
        disc, bnd = O.BV_TEST_GEOMS[gname]
        got = mirror.makeBVFeature(z[tag + "_pts"], disc, bnd)
        want = dense(z, tag)
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (gname, kind)


def test_full_size_hashes(mods):
    fast, mirror = mods
    with open(os.path.join(GOLD, "bvfeature_hashes.json")) as f:
        rows = json.load(f)
    for row in rows:
        pts = O.synth_argoverse_sweep(row["seed"], row["n"], row["kind"])
        assert sha(pts) == row["input_sha256"]
        out = mirror.makeBVFeature(pts, O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
        assert sha(out) == row["output_sha256"]


@pytest.mark.parametrize("gname", ["argo", "coarse", "odd"])
def test_ragged_batch_against_oracle(mods, gname):
    """A ragged batch with an empty sweep, a 1-point sweep and an all-outside sweep; 20 frames of the
    800 x 800 map span several L2-sized chunks."""
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    disc, bnd = O.BV_TEST_GEOMS[gname]
    kinds = ["uniform", "adversarial", "dark", "outside", "adversarial"]
    sweeps = [O.synth_argoverse_sweep(900 + i, n, kinds[i % 5], bnd)
              for i, n in enumerate([50000, 70001, 0, 1, 333, 20000, 90000] + [4000] * 13)]
    lens = [s.shape[0] for s in sweeps]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    pts = torch.from_numpy(np.concatenate(sweeps)).to(dev)
    rast = fast.BvFeatureRasterizer(disc, bnd, max_batch=len(sweeps), device=dev)
    out = torch.full((len(sweeps), 3, rast.H, rast.W), 7.0, device=dev)     # stale contents must not leak
    got = rast(pts, offsets, max(lens), out=out)
    assert got.data_ptr() == out.data_ptr()
    got = got.cpu().numpy()
    for i, s in enumerate(sweeps):
        want = O.makeBVFeature(s, disc, bnd)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (gname, i, kinds[i % 5])


@pytest.mark.parametrize("kind,floats", [("xyz_only", 3), ("wide", 6)])
def test_point_layouts(mods, kind, floats):
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    sweeps = np.stack([O.synth_argoverse_sweep(950 + i, 30000, kind) for i in range(3)])
    assert sweeps.shape[2] == floats
    rast = fast.BvFeatureRasterizer(O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY, point_floats=floats, max_batch=3)
    got = rast(torch.from_numpy(sweeps).to(dev)).cpu().numpy()              # uniform batch, no offsets
    for i in range(3):
        want = O.makeBVFeature(sweeps[i], O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32))


def test_full_batch_properties_and_graph(mods):
    """64 sweeps of 250k points (BASELINE config 3's size): per-frame invariants that need no oracle,
    equality with per-frame calls, and CUDA-graph replay."""
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    B, N = 64, 250000
    g = torch.Generator(device=dev).manual_seed(5)
    pts = torch.empty((B, N, 4), device=dev)
    pts[..., 0] = torch.rand((B, N), generator=g, device=dev) * 90 - 45
    pts[..., 1] = torch.rand((B, N), generator=g, device=dev) * 90 - 45
    pts[..., 2] = torch.rand((B, N), generator=g, device=dev) * 5 - 3.5
    pts[..., 3] = torch.randint(0, 256, (B, N), generator=g, device=dev).float()
    rast = fast.BvFeatureRasterizer(O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY, max_batch=B)
    out = rast(pts)
    torch.cuda.synchronize()
    b = O.ARGO_BV_BOUNDARY
    inside = ((pts[..., 0] >= b["minX"]) & (pts[..., 0] <= b["maxX"]) & (pts[..., 1] >= b["minY"]) &
              (pts[..., 1] <= b["maxY"]) & (pts[..., 2] >= b["minZ"]) & (pts[..., 2] <= b["maxZ"]))
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert torch.all(out[:, 2].amax(dim=(1, 2)) == 1.0)                    # intensity is normalised by the frame max
    # density below saturation counts points: sum(round(10 * d)) == kept points when no cell holds > 10
    assert int((out[:, 0] >= 1.0).sum()) == 0
    assert torch.equal(torch.round(out[:, 0] * 10).sum(dim=(1, 2)).long(), inside.sum(dim=1))
    occupied = out[:, 0] > 0
    assert torch.all((out[:, 1] > 0) <= occupied) and torch.all((out[:, 2] > 0) <= occupied)
    one = rast(pts[17:18].contiguous())
    assert torch.equal(one[0], out[17])
    want = O.makeBVFeature(pts[3].cpu().numpy(), O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
    assert np.array_equal(out[3].cpu().numpy().view(np.uint32), want.view(np.uint32))
    # graph capture: no allocation, no sync inside the call
    out2 = torch.empty_like(out)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        rast(pts, out=out2)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            rast(pts, out=out2)
        out2.fill_(3.0)
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, out)


def test_crowded_bands_overflow_their_buckets(mods):
    """Every point of a sweep inside a 0.5 m x 80 m strip: a handful of bands receive far more than the
    8x even share a bucket holds, so most records travel through the frame's overflow list."""
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    sweeps = []
    for i, n in enumerate([100000, 60000, 100000]):
        pts = O.synth_argoverse_sweep(970 + i, n, "adversarial")
        if i != 1:
            pts[:, 0] = np.float32(10.0) + (pts[:, 0] % np.float32(0.5))
        sweeps.append(pts)
    lens = [s.shape[0] for s in sweeps]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    rast = fast.BvFeatureRasterizer(O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY, max_batch=3, max_points=100000)
    got = rast(torch.from_numpy(np.concatenate(sweeps)).to(dev), offsets, 100000).cpu().numpy()
    for i, s in enumerate(sweeps):
        want = O.makeBVFeature(s, O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), i
    assert float(got[0, 0].max()) == 1.0        # saturated density in the strip


def test_argument_errors(mods):
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    with pytest.raises(ValueError):
        mirror.makeBVFeature(np.zeros((10, 2), np.float32), 0.1, O.ARGO_BV_BOUNDARY)
    with pytest.raises(TypeError):
        mirror.makeBVFeature(np.zeros((10, 4), np.float64), 0.1, O.ARGO_BV_BOUNDARY)
    rast = fast.BvFeatureRasterizer(0.1, O.ARGO_BV_BOUNDARY, max_batch=2)
    with pytest.raises(ValueError):
        rast(torch.zeros((3, 10, 4), device=dev))
    with pytest.raises(ValueError):
        rast(torch.zeros((2, 10, 3), device=dev))
    with pytest.raises(TypeError):
        rast(torch.zeros((2, 10, 4), device=dev, dtype=torch.float64))
    lib = pkg("_lib")
    p = fast.bv_params(0.1, O.ARGO_BV_BOUNDARY)
    import ctypes
    rc = lib.load().sfa_bvfeature_rasterize(None, None, 1, 10, ctypes.byref(p), None, None, 0, None)
    assert rc == -1 and "out" in lib.last_error()
    out = torch.zeros((1, 3, 800, 800), device=dev)
    rc = lib.load().sfa_bvfeature_rasterize(ctypes.c_void_p(out.data_ptr()), None, 1, 0, ctypes.byref(p),
                                            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(rast.ws.data_ptr()), 0, None)
    assert rc == -2


def test_random_geometries_and_layouts(mods):
    """30 random makeBVFeature geometries (cell size, asymmetric ranges, map sizes from a few cells to
    ~0.5 M cells, 3 / 4 / 5 floats per point) on random sweeps that partly leave the boundary: exercises
    band plans with a partial last band, maps too small / too odd for the tiled path (global-atomic
    fallback), and both point-load paths."""
    fast, mirror = mods
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(4242)
    tiled = fallback = 0
    for trial in range(30):
        disc = float(rng.choice([0.05, 0.1, 0.125, 0.2, 0.25, 0.3, 0.5, 1.0]))
        sx, sy = float(rng.uniform(6, 90)), float(rng.uniform(6, 90))
        x0, y0 = float(rng.uniform(-60, 20)), float(rng.uniform(-60, 20))
        z0, sz = float(rng.uniform(-4, 0)), float(rng.choice([2.0, 4.0, 3.5, 6.0]))
        bnd = {"minX": x0, "maxX": x0 + sx, "minY": y0, "maxY": y0 + sy, "minZ": z0, "maxZ": z0 + sz}
        H, W = O.bv_feature_shape(disc, bnd)
        if H * W > 700000 or H < 1 or W < 1:
            continue
        floats = int(rng.choice([3, 4, 4, 4, 5]))
        n = int(rng.integers(1, 60000))
        kind = ["uniform", "adversarial", "dark"][trial % 3]
        pts = O.synth_argoverse_sweep(5000 + trial, n, kind, bnd)
        if floats == 3:
            pts = np.ascontiguousarray(pts[:, :3])
        elif floats == 5:
            pts = np.concatenate([pts, rng.uniform(0, 1, (n, 1)).astype(np.float32)], axis=1)
        rast = fast.BvFeatureRasterizer(disc, bnd, point_floats=floats, max_batch=2, max_points=65536, device=dev)
        both = torch.from_numpy(np.stack([pts, pts[::-1].copy()])).to(dev)       # the reduction is order-independent
        got = rast(both).cpu().numpy()
        want = O.makeBVFeature(pts, disc, bnd)
        assert got.shape[1:] == want.shape == (3, H, W), (trial, H, W)
        assert np.array_equal(got[0].view(np.uint32), want.view(np.uint32)), (trial, disc, bnd, floats, kind)
        assert np.array_equal(got[1].view(np.uint32), want.view(np.uint32)), (trial, "reversed")
        if floats == 4 and (H * W) % 4 == 0 and H * W <= 128 * 5120:
            tiled += 1
        else:
            fallback += 1
    assert tiled >= 5 and fallback >= 5, (tiled, fallback)
