"""CPU: property tests of the oracle itself — the literal lexsort+unique port and the scatter-max restatement (the
formulation the CUDA kernels implement) must agree on arbitrary small sweeps, including the awkward values."""
import numpy as np
from hypothesis import given, settings, strategies as st

import sfa_oracle as O

SPECIAL = [0.0, -0.0, 50.0, 25.0, -25.0, -2.73, 1.27, 50 / 608, -50 / 608, 49.99999, 1e-30, -1e-30, 24.999998]
coord = st.one_of(st.floats(-60, 60, width=32), st.sampled_from(SPECIAL))
point = st.tuples(coord, coord, st.one_of(st.floats(-4, 3, width=32), st.sampled_from([-2.73, 1.27, -0.0, 0.5, 0.25])),
                  st.floats(0, 1, width=32))


@settings(max_examples=60, deadline=None)
@given(st.lists(point, min_size=0, max_size=300), st.sampled_from(["kitti", "back", "argo"]))
def test_scatter_formulation_equals_lexsort_port(pts, gname):
    geom = {"kitti": O.KITTI, "back": O.KITTI_BACK, "argo": O.ARGOVERSE}[gname]
    sweep = np.asarray(pts, dtype=np.float32).reshape(-1, 4)
    # quantise z so that ties are frequent
    sweep[:, 2] = np.round(sweep[:, 2] * 4) / 4
    filt = O.get_filtered_lidar(sweep.copy(), geom.boundary)
    ref = O.makeBEVMap(filt, geom.boundary, geom).astype(np.float32)
    got = O.make_bev_scatter(sweep, geom, True, np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    rows, cols, win, counts, _ = O.bev_cell_selection(sweep, geom, True)
    assert counts.sum() <= sweep.shape[0] and len(set(zip(rows.tolist(), cols.tolist()))) == len(rows)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 20))
def test_decode_oracle_global_topk_property(seed, K):
    """decode's two-stage top-K equals a global top-K of the NMS'd map whenever the K scores are distinct."""
    import torch
    hm, off, d, z, dim = O.synth_heads(seed, B=1, C=3, h=12, w=9, tie_free=True)
    det = O.decode(hm.clone(), off, d, z, dim, K=K).numpy()[0]
    nms = O._nms(hm.clone())[0].reshape(-1).numpy()
    order = np.argsort(-nms, kind="stable")[:K]
    if len(np.unique(nms[order])) == K and nms[order][-1] > 0:
        assert np.array_equal(det[:, 0], nms[order])
        assert np.array_equal(det[:, 9].astype(int), order // (12 * 9))


bv_point = st.tuples(st.one_of(st.floats(-45, 45, width=32), st.sampled_from([-40.0, 40.0, 0.0, -0.0, 39.95, -39.95])),
                     st.one_of(st.floats(-45, 45, width=32), st.sampled_from([-40.0, 40.0, 0.1, -0.1])),
                     st.one_of(st.floats(-3.5, 1.5, width=32), st.sampled_from([-3.0, 1.0, -2.75])),
                     st.one_of(st.floats(-5, 255, width=32), st.sampled_from([0.0, -0.0, 255.0])))


def _bv_loop(points, discretization, boundary):
    """makeBVFeature's per-point loop (argoverse_test.py:238-242) spelled out, as a check of the
    loop-free oracle on small sweeps."""
    H, W = O.bv_feature_shape(discretization, boundary)
    x, y, z, i = (points[:, k] for k in range(4))
    mask = ((x >= boundary["minX"]) & (x <= boundary["maxX"]) & (y >= boundary["minY"]) & (y <= boundary["maxY"]) &
            (z >= boundary["minZ"]) & (z <= boundary["maxZ"]))
    if not np.any(mask):
        return np.zeros((3, H, W), dtype=np.float32)
    x, y, z, i = x[mask], y[mask], z[mask], i[mask]
    xi = np.clip(((boundary["maxX"] - x) / discretization).astype(np.int32), 0, H - 1)
    yi = np.clip(((y - boundary["minY"]) / discretization).astype(np.int32), 0, W - 1)
    h_map, d_map, i_map = (np.zeros((H, W), dtype=np.float32) for _ in range(3))
    for a, b, zz, ii in zip(xi, yi, z - boundary["minZ"], i):
        h_map[a, b] = max(h_map[a, b], zz)
        i_map[a, b] = max(i_map[a, b], ii)
        d_map[a, b] += 1
    d_map = np.clip(d_map / 10.0, 0, 1)
    if h_map.max() > 0:
        h_map = h_map / (boundary["maxZ"] - boundary["minZ"])
    if i_map.max() > 0:
        i_map = i_map / i_map.max()
    return np.stack([d_map, h_map, i_map], axis=0).astype(np.float32)


@settings(max_examples=60, deadline=None)
@given(st.lists(bv_point, min_size=0, max_size=200), st.integers(0, 2 ** 31 - 1))
def test_bvfeature_oracle_equals_the_loop_and_ignores_point_order(pts, seed):
    sweep = np.asarray(pts, dtype=np.float32).reshape(-1, 4)
    sweep[:, 2] = np.round(sweep[:, 2] * 4) / 4
    got = O.makeBVFeature(sweep, O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
    assert np.array_equal(got.view(np.uint32), _bv_loop(sweep, O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY).view(np.uint32))
    perm = np.random.default_rng(seed).permutation(len(sweep))
    assert np.array_equal(O.makeBVFeature(sweep[perm], O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY).view(np.uint32),
                          got.view(np.uint32))
    assert got.min() >= 0 and got[0].max() <= 1 and got[2].max() <= 1


@settings(max_examples=40, deadline=None)
@given(st.floats(-3.1, 3.1), st.floats(-5, 5), st.floats(-5, 5), st.integers(0, 2 ** 31 - 1))
def test_point_transform_oracle_properties(rz, tx, ty, seed):
    pts = np.random.default_rng(seed).uniform(-50, 50, (50, 3)).astype(np.float32)
    assert np.array_equal(O.point_transform(pts, 0, 0, 0), pts.astype(np.float64))           # identity chain is exact
    out = O.point_transform(pts, tx, ty, 0.5, rz=rz)
    assert np.array_equal(out[:, 2], pts[:, 2].astype(np.float64) + 0.5)                      # z only translated
    back = O.point_transform(out, 0, 0, 0, rz=-rz) - [tx, ty, 0.5]                            # (p + t) R R^-1 - t
    np.testing.assert_allclose(back, pts, atol=1e-4)                                          # rotation is orthogonal
    chain = np.hstack([pts, np.ones((50, 1))])
    for m in O.transform_matrices(tx, ty, 0.5, rz=rz):
        chain = chain @ m
    assert np.array_equal(chain[:, :3], out)
