"""CPU, build container only: re-run the oracle against the REFERENCE's own functions live
(imported in place from /root/reference by oracle/ref_loader.py).  Skipped wherever the
reference is absent (e.g. on the GPU box) — the committed fixtures cover that case."""
import numpy as np
import pytest
import torch

import ref_loader
import sfa_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

GEOMS = {"kitti": O.KITTI, "kitti_back": O.KITTI_BACK, "argoverse": O.ARGOVERSE}


@pytest.fixture(scope="module")
def ns():
    return ref_loader.load()


@pytest.mark.parametrize("gname,kind,n,seed", [
    ("kitti", "uniform", 120000, 5), ("kitti", "zties", 60000, 6), ("kitti", "gridaligned", 30000, 7),
    ("kitti", "bounds", 30000, 8), ("kitti", "nonfinite", 30000, 9), ("kitti_back", "zties", 60000, 10),
    ("argoverse", "zties", 250000, 11), ("kitti", "onecell", 20000, 12)])
def test_bev_oracle_equals_reference(ns, gname, kind, n, seed):
    geom = GEOMS[gname]
    sweep = O.synth_sweep(seed, n, geom, kind)
    with ref_loader.patched_geometry(ns, geom.boundary, geom.BEV_HEIGHT, geom.BEV_WIDTH, geom.DISCRETIZATION):
        filt = ns.get_filtered_lidar(sweep.copy(), geom.boundary)
        ref = ns.makeBEVMap(filt, geom.boundary)
    assert np.array_equal(O.get_filtered_lidar(sweep.copy(), geom.boundary).view(np.uint32), filt.view(np.uint32))
    assert np.array_equal(O.makeBEVMap(filt, geom.boundary, geom).view(np.uint64), ref.view(np.uint64))
    assert np.array_equal(O.make_bev_scatter(sweep, geom, True, np.float32).view(np.uint32),
                          ref.astype(np.float32).view(np.uint32))


def test_reference_raises_indexerror_outside_map(ns):
    sweep = O.synth_sweep(2, 1000, O.KITTI, "uniform")
    sweep[:3, 0] = 500.0
    with pytest.raises(IndexError):
        ns.makeBEVMap(sweep, O.KITTI.boundary)
    with pytest.raises(IndexError):
        O.bev_cell_selection(sweep, O.KITTI, apply_filter=False)


@pytest.mark.parametrize("seed,tie_free", [(1, True), (2, False), (3, True)])
def test_decode_oracle_equals_reference(ns, seed, tie_free):
    hm, off, direction, z, dim = O.synth_heads(seed, B=3, tie_free=tie_free)
    ref = ns.decode(hm.clone(), off.clone(), direction, z, dim, K=50).numpy()
    got = O.decode(hm.clone(), off.clone(), direction, z, dim, K=50).numpy()
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    detn = ref.astype(np.float32)
    want = ns.post_processing_pristine(detn.copy(), 3, 4, 0.2)
    have = O.post_processing(detn.copy(), 3, 4, 0.2)
    for w, h in zip(want, have):
        for j in range(3):
            assert np.array_equal(np.asarray(w[j], np.float32).reshape(-1, 8), np.asarray(h[j], np.float32).reshape(-1, 8))
    live = ns.post_processing_live(detn.copy(), 3, 4, 0.2)
    assert len(live) == 1 and len(O.post_processing_live_semantics(detn)) == 1
    for j in range(3):
        assert np.array_equal(np.asarray(live[0][j]).reshape(-1, 8), np.asarray(have[-1][j]).reshape(-1, 8))


def test_sigmoid_matches(ns):
    x = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(0))
    assert torch.equal(ns._sigmoid(x.clone()), O._sigmoid(x.clone()))


class _Calib:
    def __init__(self, V2C, R0, P2):
        self.V2C, self.R0, self.P2 = V2C, R0, P2


@pytest.mark.parametrize("seed", [21, 22])
def test_projection_oracle_equals_reference(ns, seed):
    """lidar_to_camera_box (data_process/transformation.py:99-107) and convert_sfa3d_to_2d_boxes
    (test6.py:129-187, compiled out of the script by ref_loader) on decoded detections."""
    heads = O.synth_heads(seed, B=2, tie_free=True)
    det = O.decode(*heads, K=50).numpy().astype(np.float32)
    V2C, R0, P2 = O.synth_calibration(seed)
    for sample in O.post_processing(det, peak_thresh=0.2):
        real = np.asarray(ns.convert_det_to_real_values(sample), np.float64).reshape(-1, 8)
        assert np.array_equal(ns.lidar_to_camera_box(real[:, 1:], V2C, R0, P2),
                              O.lidar_to_camera_box(real[:, 1:], V2C, R0, P2))
        want = ns.convert_sfa3d_to_2d_boxes(sample, _Calib(V2C, R0, P2), (375, 1242))
        assert want == O.convert_sfa3d_to_2d_boxes(sample, V2C, R0, P2, (375, 1242))
        assert len(want[0]) > 5


@pytest.mark.parametrize("gname,kind,n,seed", [
    ("argo", "uniform", 100000, 31), ("argo", "adversarial", 100000, 32), ("coarse", "adversarial", 50000, 33),
    ("odd", "wide", 20000, 34), ("argo", "xyz_only", 20000, 35), ("argo", "outside", 1000, 36), ("odd", "dark", 5000, 37)])
def test_bvfeature_oracle_equals_reference(ns, gname, kind, n, seed):
    """makeBVFeature (argoverse_test.py:199-254, compiled out of the script) vs the loop-free restatement."""
    disc, bnd = O.BV_TEST_GEOMS[gname]
    pts = O.synth_argoverse_sweep(seed, n, kind, bnd)
    want = ns.makeBVFeature(pts, disc, bnd)
    got = O.makeBVFeature(pts, disc, bnd)
    assert want.dtype == got.dtype == np.float32 and want.shape == got.shape
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
