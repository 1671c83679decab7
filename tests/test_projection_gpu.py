"""GPU: lidar-frame boxes -> camera frame -> image boxes (sfa_project_boxes, through the C ABI) against
the reference fixtures (tests/golden/projection_small.npz) and the oracle.  float64 arithmetic on
both sides; numpy's matrix products may fuse or reorder the 3-4 term sums, hence 1e-9 relative on
the floats, and the integer boxes exact unless a bound sits within 1e-6 of an integer."""
import os

import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9


@pytest.fixture(scope="module")
def mods():
    return pkg("fast"), pkg("utils.evaluation_utils"), pkg("data_process.transformation")


@pytest.fixture(scope="module")
def fixture():
    return np.load(os.path.join(GOLD, "projection_small.npz"))


def oracle_dense(real, V2C, R0, P2, shape, thr):
    """Per row: (cam[7], bounds[4], ok) from the oracle."""
    cams, bounds, oks = [], [], []
    for r in real:
        cam = O.lidar_to_camera_box(r[1:].reshape(1, -1), V2C, R0, P2)[0]
        b, ok = O.project_box_to_image(cam, P2, shape)
        cams.append(cam), bounds.append(b), oks.append(ok and not (r[0] < thr))
    return np.array(cams).reshape(-1, 7), np.array(bounds).reshape(-1, 4), np.array(oks, bool)


def ints_match(box, valid, bounds, oks):
    assert np.array_equal(valid, oks)
    for b, f in zip(box[valid], bounds[valid]):
        want = [int(f[0]), int(f[1]), int(f[2] - f[0]), int(f[3] - f[1])]
        vals = [f[0], f[1], f[2] - f[0], f[3] - f[1]]
        for g, w, v in zip(b, want, vals):
            assert g == w or abs(v - round(v)) < 1e-6, (b, want)


def test_fixture_cases_dense_and_mirror(mods, fixture):
    fast, ev, tr = mods
    z = fixture
    for n in [str(s) for s in z["names"]]:
        dets = {j: z["%s_det%d" % (n, j)] for j in range(3)}
        V2C, R0, P2, shape = z[n + "_V2C"], z[n + "_R0"], z[n + "_P2"], tuple(int(v) for v in z[n + "_shape"])
        boxes, conf = ev.convert_sfa3d_to_2d_boxes(dets, {"V2C": V2C, "R0": R0, "P2": P2}, shape)
        assert np.array_equal(np.asarray(boxes, np.int64).reshape(-1, 4), z[n + "_boxes"]), n
        assert np.array_equal(np.asarray(conf, np.float64), z[n + "_conf"]), n
        real = z[n + "_real"]
        got = tr.lidar_to_camera_box(real[:, 1:], V2C, R0, P2)
        assert got.shape == z[n + "_cam"].shape and got.dtype == np.float64
        np.testing.assert_allclose(got, z[n + "_cam"], rtol=RTOL, atol=1e-12, equal_nan=True)


def test_mirror_accepts_calibration_object_and_defaults(mods, fixture):
    fast, ev, tr = mods
    z = fixture

    class Calib:
        V2C, R0, P2 = z["decode1_V2C"], z["decode1_R0"], z["decode1_P2"]

    dets = {j: z["decode1_det%d" % j] for j in range(3)}
    boxes, conf = ev.convert_sfa3d_to_2d_boxes(dets, Calib, (375, 1242))
    assert np.array_equal(np.asarray(boxes, np.int64).reshape(-1, 4), z["decode1_boxes"])
    b02, c02 = ev.convert_sfa3d_to_2d_boxes(dets, Calib, (375, 1242), min_confidence=0.2)
    assert (b02, c02) == (boxes, conf)          # class ids are 0, 1, 2: the 0.2 of msac/slam selects the same rows
    assert ev.convert_sfa3d_to_2d_boxes({}, Calib, (375, 1242)) == ([], [])
    # no calibration given: the dataset-average matrices of config/kitti_config.py:64-83
    cnf = pkg("config.kitti_config")
    boxes7 = z["decode1_real"][:5, 1:]
    want = O.lidar_to_camera_box(boxes7, cnf.Tr_velo_to_cam[:3], cnf.R0[:3, :3])
    np.testing.assert_allclose(tr.lidar_to_camera_box(boxes7), want, rtol=RTOL, atol=1e-12)
    x, y, zc = tr.lidar_to_camera(10.0, 2.0, -1.0)
    np.testing.assert_allclose([x, y, zc], O.lidar_to_camera(10.0, 2.0, -1.0, cnf.Tr_velo_to_cam[:3], cnf.R0[:3, :3]),
                               rtol=RTOL)
    assert tr.lidar_to_camera_box(np.zeros((0, 7))).shape == (0, 7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_dense_batch_per_frame_calibration(mods, dtype):
    """The device-resident chain decode -> post_process(real) -> project at B=16, K=50, one
    calibration per frame, against the oracle row by row."""
    fast, ev, tr = mods
    dev = torch.device("cuda", 0)
    B, K = 16, 50
    heads = O.synth_heads(77, B=B, tie_free=True)
    det = fast.decode_device(*[t.to(dev) for t in heads], K=K)
    rows, cls, keep, real = fast.post_process_dense(det, real=True)
    calibs = [O.synth_calibration(100 + b) for b in range(B)]
    packed = fast.pack_calibration(np.stack([c[0] for c in calibs]), np.stack([c[1] for c in calibs]),
                                   np.stack([c[2] for c in calibs]), device=dev)
    assert packed.shape == (B, 33)
    box, valid, cam, box_f = fast.project_boxes_dense(real.to(dtype), packed, (375, 1242), keep=keep,
                                                      want_cam=True, want_float=True)
    torch.cuda.synchronize()
    real_h, keep_h = real.cpu().numpy().astype(np.float64), keep.cpu().numpy()
    n_valid = 0
    for b in range(B):
        cams, bounds, oks = oracle_dense(real_h[b], *calibs[b], (375, 1242), 0.3)
        np.testing.assert_allclose(cam[b].cpu().numpy(), cams, rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(box_f[b].cpu().numpy(), bounds, rtol=RTOL, atol=1e-9)
        ints_match(box[b].cpu().numpy(), valid[b].cpu().numpy(), bounds, oks & keep_h[b])
        n_valid += int(valid[b].sum())
    assert n_valid > 100
    assert int(box[~valid].abs().sum()) == 0


def test_nan_inf_and_degenerate_rows(mods):
    fast, ev, tr = mods
    dev = torch.device("cuda", 0)
    V2C, R0, P2 = O.synth_calibration(3)
    real = np.array([
        [1, 20.0, 0.0, -1.0, 1.5, 1.6, 4.0, 0.3],
        [1, 20.0, 0.0, np.nan, 1.5, 1.6, 4.0, 0.3],      # NaN location: bounds fall back to the image border
        [1, 20.0, 0.0, -1.0, np.nan, 1.6, 4.0, 0.3],     # NaN height poisons every row of R @ corners (0 * NaN)
        [1, 20.0, 0.0, -1.0, 1.5, np.inf, 4.0, 0.3],     # inf width: inf - inf appears in the projection
        [1, 0.28, 0.0, -1.0, 1.5, 1.6, 4.0, 0.0],        # straddles the image plane
        [1, -20.0, 0.0, -1.0, 1.5, 1.6, 4.0, 0.3],       # behind the camera
        [1, 20.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0],        # zero-size box: rejected
        [0, 20.0, 0.0, -1.0, 1.5, 1.6, 4.0, 0.3],        # class 0 < 0.3: skipped
        [2, 20.0, 300.0, -1.0, 1.5, 1.6, 4.0, 0.3],      # far outside the image
        [2, 20.0, 0.0, -1.0, 1.5, 1.6, 4.0, 1e6],        # large angle (slow path of sin/cos)
    ], dtype=np.float64)
    packed = fast.pack_calibration(V2C, R0, P2, device=dev)
    box, valid, cam, box_f = fast.project_boxes_dense(torch.from_numpy(real[None]).to(dev), packed, (375, 1242),
                                                      want_cam=True, want_float=True)
    with np.errstate(all="ignore"):
        cams, bounds, oks = oracle_dense(real, V2C, R0, P2, (375, 1242), 0.3)
    np.testing.assert_allclose(cam[0].cpu().numpy(), cams, rtol=RTOL, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(box_f[0].cpu().numpy(), bounds, rtol=1e-7, atol=1e-6, equal_nan=True)
    ints_match(box[0].cpu().numpy(), valid[0].cpu().numpy(), bounds, oks)
    v = valid[0].cpu().numpy()
    assert v[0] and v[1] and not v[6] and not v[7]
    assert box[0, 1].tolist() == [0, 0, 1242, 375]


def test_preallocated_outputs_in_a_cuda_graph(mods):
    fast, ev, tr = mods
    dev = torch.device("cuda", 0)
    heads = O.synth_heads(78, B=4, tie_free=True)
    det = fast.decode_device(*[t.to(dev) for t in heads], K=50)
    rows, cls, keep, real = fast.post_process_dense(det, real=True)
    packed = fast.pack_calibration(*O.synth_calibration(0), device=dev)
    want_box, want_valid = fast.project_boxes_dense(real, packed, (375, 1242), keep=keep)
    box = torch.zeros((4, 50, 4), dtype=torch.int32, device=dev)
    valid = torch.zeros((4, 50), dtype=torch.uint8, device=dev)
    keep_u8 = keep.to(torch.uint8)
    lib = pkg("_lib")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fast.project_boxes_dense(real, packed, (375, 1242), keep=keep_u8, out=(box, valid))
        g = torch.cuda.CUDAGraph()
        box.zero_(), valid.zero_()
        with torch.cuda.graph(g, stream=s):
            fast.project_boxes_dense(real, packed, (375, 1242), keep=keep_u8, out=(box, valid))
        before = lib.kernel_launches()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(box, want_box) and torch.equal(valid.bool(), want_valid)
    assert lib.kernel_launches() == before


def test_argument_errors(mods):
    fast, ev, tr = mods
    dev = torch.device("cuda", 0)
    real = torch.zeros((2, 5, 8), device=dev)
    packed = fast.pack_calibration(*O.synth_calibration(0), device=dev)
    with pytest.raises(ValueError):
        fast.project_boxes_dense(real[..., :7], packed, (375, 1242))
    with pytest.raises(ValueError):
        fast.project_boxes_dense(real, packed.float(), (375, 1242))
    with pytest.raises(ValueError):
        fast.project_boxes_dense(real, packed.repeat(3, 1), (375, 1242))
    with pytest.raises(TypeError):
        fast.project_boxes_dense(real.cpu(), packed, (375, 1242))
    with pytest.raises(pkg("_lib").SfaError):
        fast.project_boxes_dense(real, packed, (0, 1242))
    box, valid = fast.project_boxes_dense(real[:0], packed, (375, 1242))
    assert box.shape == (0, 5, 4)
