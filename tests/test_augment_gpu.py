"""GPU: sweep augmentation (sfa_transform_points, through the C ABI) against the reference fixture
(tests/golden/augment_small.npz) and the oracle: float64 results and float32 sweeps bit-identical."""
import os

import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mods(cuda_device):
    return pkg("fast"), pkg("data_process.transformation")


@pytest.fixture(scope="module")
def fixture():
    return np.load(os.path.join(GOLD, "augment_small.npz"))


def test_point_transform_fixture(mods, fixture):
    fast, tr = mods
    z = fixture
    sweep = z["sweep"]
    for k, prm in enumerate(z["params"]):
        got = tr.point_transform(sweep[:, 0:3], *prm)
        assert got.dtype == np.float64 and got.shape == (len(sweep), 3)
        assert np.array_equal(got, z["pt%d" % k]), prm
        got64 = tr.point_transform(sweep[:, 0:3].astype(np.float64) * 1.000001, *prm)
        assert np.array_equal(got64, z["pt%d_f64in" % k]), prm
    assert tr.point_transform(np.zeros((0, 3), np.float32), 0, 0, 0, rz=0.1).shape == (0, 3)


def test_random_rotation_and_scaling_mirrors_follow_the_rng_stream(mods, fixture):
    fast, tr = mods
    z = fixture
    sweep, labels = z["sweep"], z["labels"]
    for k, seed in enumerate((1, 2, 3, 4)):
        np.random.seed(seed)
        # the label side is the caller's (reference's) box_transform; an identity stands in for it here
        lidar, lab = tr.Random_Rotation(limit_angle=np.pi / 4, p=1.0, box_transform=lambda b, *a, **kw: b)(sweep.copy(), labels.copy())
        assert np.array_equal(lidar.view(np.uint32), z["rot%d" % k].view(np.uint32))
        assert np.random.random() == float(z["rot%d_next" % k]) and np.array_equal(lab, labels)
        np.random.seed(seed)
        lidar, lab = tr.Random_Scaling(scaling_range=(0.95 + 0.01 * k, 1.05), p=1.0)(sweep.copy(), labels.copy())
        assert np.array_equal(lidar.view(np.uint32), z["scl%d" % k].view(np.uint32))
        assert np.array_equal(lab, z["scl%d_labels" % k]) and np.random.random() == float(z["scl%d_next" % k])
    np.random.seed(7)
    lidar, _ = tr.Random_Rotation(p=0.0)(sweep.copy(), labels.copy())     # not drawn: untouched
    assert np.array_equal(lidar.view(np.uint32), sweep.view(np.uint32))
    # labels without a box_transform: refuse (the reference always rotates the labels with the sweep, transformation.py:351)
    with pytest.raises(ValueError):
        tr.Random_Rotation(p=1.0)(sweep.copy(), labels.copy())
    np.random.seed(7)
    lidar, lab = tr.Random_Rotation(p=1.0)(sweep.copy(), np.zeros((0, 8), np.float32))    # no labels: fine
    assert lab.shape == (0, 8) and not np.array_equal(lidar, sweep)
    called = []
    tr.Random_Rotation(p=1.0, box_transform=lambda b, *a, **k: called.append((a, k)) or b)(sweep.copy(), labels.copy())
    assert called and called[0][1]["coordinate"] == "lidar"


def test_batched_device_form_then_raster(mods):
    """Ragged batch, one rotation + scale per sweep, then the BEV raster of the augmented sweeps:
    the whole training-side chain on the device equals oracle augmentation + oracle raster."""
    fast, tr = mods
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    sweeps = [O.synth_sweep(610 + i, n, O.KITTI, "outside") for i, n in enumerate([30000, 1, 0, 45001, 20000])]
    lens = [s.shape[0] for s in sweeps]
    angles = rng.uniform(-np.pi / 4, np.pi / 4, len(sweeps))
    factors = rng.uniform(0.95, 1.05, len(sweeps))
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    pts = torch.from_numpy(np.concatenate(sweeps)).to(dev)
    mats = torch.from_numpy(np.stack([np.stack(O.transform_matrices(0, 0, 0, rz=a)) for a in angles])).to(dev)
    scales = torch.tensor(factors, dtype=torch.float32, device=dev)
    aug = fast.transform_points_device(pts, mats=mats, scales=scales, offsets=offsets, max_points=max(lens))
    assert aug.data_ptr() != pts.data_ptr() and aug.dtype == torch.float32 and aug.shape == pts.shape
    want = [O.random_scaling_points(O.random_rotation_points(s, a), float(np.float32(f)))
            for s, a, f in zip(sweeps, angles, factors)]
    got = aug.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), np.concatenate(want).view(np.uint32))
    inplace = pts.clone()
    assert fast.transform_points_device(inplace, mats=mats, scales=scales, offsets=offsets, max_points=max(lens),
                                        out=inplace).data_ptr() == inplace.data_ptr()
    assert torch.equal(inplace, aug)
    geom = pkg("geometry").from_config(pkg("config.kitti_config"))
    bev = fast.BevRasterizer(geom, max_batch=len(sweeps), max_points=max(lens), device=dev)(aug, offsets, max(lens))
    for i, w in enumerate(want):
        assert np.array_equal(bev[i].cpu().numpy().view(np.uint32), O.make_bev_scatter(w, O.KITTI, True, np.float32).view(np.uint32))


def test_uniform_batch_xyz_only_and_errors(mods):
    fast, tr = mods
    dev = torch.device("cuda", 0)
    pts = torch.from_numpy(np.stack([O.synth_sweep(620 + i, 5000, O.KITTI, "uniform")[:, :3] for i in range(3)])).to(dev)
    mats = torch.from_numpy(np.stack([np.stack(O.transform_matrices(0.5 * i, -1.0, 0.25, rx=0.1 * i, rz=0.2)) if i else
                                      np.stack(O.transform_matrices(0, 0, 0, rx=1e-9, rz=0.2)) for i in range(3)])).to(dev)
    out = fast.transform_points_device(pts, mats=mats, out_dtype=torch.float64).cpu().numpy()
    for i in range(3):
        m = mats[i].cpu().numpy()
        chain = np.hstack([pts[i].cpu().numpy(), np.ones((5000, 1))])
        for k in range(m.shape[0]):
            chain = np.matmul(chain, m[k])
        assert np.array_equal(out[i], chain[:, 0:3])
    with pytest.raises(ValueError):
        fast.transform_points_device(pts, mats=mats[:2])
    with pytest.raises(ValueError):
        fast.transform_points_device(pts, mats=mats.float())
    with pytest.raises(TypeError):
        fast.transform_points_device(pts.cpu(), mats=mats)
    with pytest.raises(pkg("_lib").SfaError):
        fast.transform_points_device(pts, scales=torch.ones(3, device=dev), out_dtype=torch.float64)
