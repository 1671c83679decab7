"""GPU parity of stage A (sweep filter + BEV rasterisation) against the CPU oracle, through the
C ABI (ctypes).  Integer quantities (occupied cells, winning point per cell, counts) and therefore
all three float planes must be BIT-EXACT (the float maps are produced by the same fp32 operations;
contract tolerance 1e-6 relative is asserted as well, and is met with zero error)."""
import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu

KINDS = ["uniform", "outside", "zties", "gridaligned", "bounds", "nonfinite", "onecell", "clustered"]


AUTO, TILED, ATOMIC, TWO_KERNEL = 0, 1, 2, 3   # enum SfaBevAlgorithm (include/sfa_b200.h)
# tiled = the single persistent bin+band kernel (bev_fused); two_kernel = the same algorithm as bev_bin + bev_band launches
ALGOS = [pytest.param(TILED, id="tiled"), pytest.param(TWO_KERNEL, id="two_kernel"), pytest.param(ATOMIC, id="atomic")]
TILED_ALGOS = [pytest.param(TILED, id="tiled"), pytest.param(TWO_KERNEL, id="two_kernel")]


def _geom(ogeom, apply_filter=True, algorithm=AUTO):
    class Cnf:
        BEV_HEIGHT, BEV_WIDTH, DISCRETIZATION = ogeom.BEV_HEIGHT, ogeom.BEV_WIDTH, ogeom.DISCRETIZATION
    return pkg("geometry").BevGeometry(ogeom.boundary, Cnf, apply_filter=apply_filter, algorithm=algorithm)


def _run_batch(cuda_device, sweeps, ogeom, apply_filter=True, max_batch=None, algorithm=AUTO):
    fast = pkg("fast")
    lens = [s.shape[0] for s in sweeps]
    rast = fast.BevRasterizer(_geom(ogeom, apply_filter, algorithm), max_batch=max_batch or max(1, len(sweeps)),
                              max_points=max(lens + [1]), device=cuda_device)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda_device)
    allpts = np.concatenate(sweeps, 0) if sum(lens) else np.zeros((0, 4), np.float32)
    pts = torch.from_numpy(allpts).to(cuda_device)
    if pts.numel() == 0:
        pts = torch.zeros((1, 4), dtype=torch.float32, device=cuda_device)[:0]
    out = rast(pts, offsets, max(lens) if lens else 0)
    torch.cuda.synchronize()
    return out.cpu().numpy(), rast


def _assert_bit_exact(got, want, what):
    assert got.dtype == np.float32 and want.dtype == np.float32
    if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        raise AssertionError("%s: %d cells differ, first at %s: got %r want %r" %
                             (what, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])]))
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)  # the contract's tolerance, trivially met


@pytest.mark.parametrize("algorithm", ALGOS)
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [0, 1, 5000, 120000])
def test_bev_single_sweep_bit_exact(cuda_device, kind, n, algorithm):
    n = min(n, 100000) if kind == "onecell" else n
    sweep = O.synth_sweep(17, n, O.KITTI, kind) if n else np.zeros((0, 4), np.float32)
    got, rast = _run_batch(cuda_device, [sweep], O.KITTI, algorithm=algorithm)
    want = O.make_bev_scatter(sweep, O.KITTI, True, np.float32)
    _assert_bit_exact(got[0], want, "%s n=%d" % (kind, n))
    assert rast.out_of_map_points() == 0


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_reference_fixture_hashes(cuda_device, algorithm):
    """One hop from the CUDA map to the REFERENCE: sha256 of the GPU output against tests/golden/bev_hashes.json, the
    hashes of the unmodified reference's makeBEVMap output (.astype(float32)) on the same seeded full-size sweeps
    (120k KITTI / 250k Argoverse range; written by tests/golden/make_golden.py, which imports the reference)."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "bev_hashes.json")) as f:
        cases = json.load(f)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    geoms = {"kitti": O.KITTI, "kitti_back": O.KITTI_BACK, "argoverse": O.ARGOVERSE}
    checked = 0
    for c in cases:
        geom = geoms[c["geom"]]
        sweep = O.synth_sweep(c["seed"], c["n"], geom, c["kind"])
        if sha(sweep) != c["input_sha256"]:
            continue   # numpy's generator changed since the fixture was made: nothing to compare with
        got, _ = _run_batch(cuda_device, [sweep], geom, algorithm=algorithm)
        assert sha(got[0]) == c["bev_f32_sha256"], (c["geom"], c["kind"], c["n"])
        assert int(np.count_nonzero(got[0][2])) == c["occupied"]
        checked += 1
    assert checked >= 12


def test_bev_matches_lexsort_formulation(cuda_device):
    """Against the oracle's literal lexsort+unique port (float64 map cast to float32)."""
    sweep = O.synth_sweep(3, 120000, O.KITTI, "zties")
    got, _ = _run_batch(cuda_device, [sweep], O.KITTI)
    ref = O.makeBEVMap(O.get_filtered_lidar(sweep.copy(), O.KITTI.boundary), O.KITTI.boundary, O.KITTI)
    _assert_bit_exact(got[0], ref.astype(np.float32), "lexsort port")


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_batch64_ragged_and_ring_reuse(cuda_device, algorithm):
    """BASELINE config[1] shape: 64 sweeps; ragged sizes; the workspace ring is reused several times,
    and the same geometry is run twice (the workspace must come back clean)."""
    rng = np.random.default_rng(5)
    sweeps = []
    for i in range(64):
        n = int(rng.integers(0, 130000)) if i % 7 else 120000
        sweeps.append(O.synth_sweep(100 + i, n, O.KITTI, KINDS[i % len(KINDS)] if n <= 100000 else "uniform")
                      if n else np.zeros((0, 4), np.float32))
    for attempt in range(2):
        got, rast = _run_batch(cuda_device, sweeps, O.KITTI, algorithm=algorithm)
        for i, s in enumerate(sweeps):
            _assert_bit_exact(got[i], O.make_bev_scatter(s, O.KITTI, True, np.float32), "frame %d" % i)


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_second_call_same_workspace(cuda_device, algorithm):
    fast = pkg("fast")
    rast = fast.BevRasterizer(_geom(O.KITTI, algorithm=algorithm), max_batch=4, max_points=61000, device=cuda_device)
    for seed in (1, 2, 3):
        sweeps = [O.synth_sweep(seed * 10 + j, 50000 + 1000 * j, O.KITTI, "zties") for j in range(11)]
        lens = [s.shape[0] for s in sweeps]
        offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda_device)
        pts = torch.from_numpy(np.concatenate(sweeps, 0)).to(cuda_device)
        got = rast(pts, offsets, max(lens)).cpu().numpy()
        for i, s in enumerate(sweeps):
            _assert_bit_exact(got[i], O.make_bev_scatter(s, O.KITTI, True, np.float32), "seed %d frame %d" % (seed, i))


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_back_boundary_negative_row_wrap(cuda_device, algorithm):
    """boundary_back (config/kitti_config.py:36-43): x in [-50, 0] -> negative rows wrap (numpy
    negative indexing) — rows 1..607 occupied, row 0 only from x == 0."""
    sweep = O.synth_sweep(8, 80000, O.KITTI_BACK, "uniform")
    got, _ = _run_batch(cuda_device, [sweep], O.KITTI_BACK, algorithm=algorithm)
    _assert_bit_exact(got[0], O.make_bev_scatter(sweep, O.KITTI_BACK, True, np.float32), "back")


@pytest.mark.parametrize("algorithm", ALGOS)
@pytest.mark.parametrize("kind", ["uniform", "zties", "outside"])
def test_bev_argoverse_range_250k(cuda_device, kind, algorithm):
    """BASELINE config[3]: Argoverse-range sweeps (~250k points, +-50 m, z in [-3, 5])."""
    sweep = O.synth_sweep(21, 250000, O.ARGOVERSE, kind)
    got, _ = _run_batch(cuda_device, [sweep], O.ARGOVERSE, algorithm=algorithm)
    _assert_bit_exact(got[0], O.make_bev_scatter(sweep, O.ARGOVERSE, True, np.float32), "argoverse " + kind)


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_without_filter_negative_z_and_nan_z(cuda_device, algorithm):
    """makeBEVMap alone (apply_filter=0): z not shifted, may be negative; NaN z sorts last."""
    sweep = O.synth_sweep(9, 60000, O.KITTI, "zties")
    sweep[::97, 2] = np.nan
    got, _ = _run_batch(cuda_device, [sweep], O.KITTI, apply_filter=False, algorithm=algorithm)
    want = O.make_bev_scatter(sweep, O.KITTI, apply_filter=False, dtype=np.float32)
    assert np.array_equal(got[0], want, equal_nan=True)
    assert np.array_equal(np.isnan(got[0]), np.isnan(want))


def test_bev_odd_grid_scalar_finalize(cuda_device):
    """H*W not a multiple of 4: AUTO falls back to the global-atomic path with its scalar finalize."""
    g = O.Geometry(boundary={"minX": 0, "maxX": 30, "minY": -15, "maxY": 15, "minZ": -2, "maxZ": 2},
                   BEV_HEIGHT=301, BEV_WIDTH=301)
    sweep = O.synth_sweep(4, 40000, g, "zties")
    got, _ = _run_batch(cuda_device, [sweep], g)
    _assert_bit_exact(got[0], O.make_bev_scatter(sweep, g, True, np.float32), "301x301")


@pytest.mark.parametrize("geomspec", [(304, 304), (1000, 1000), (1700, 1700), (2400, 2400), (100, 36)])
def test_bev_other_grids_tiled_band_plans(cuda_device, geomspec):
    """Other map sizes exercise other band plans of the tiled path (fewer / more than 128 bands, a last
    band that is partly outside the map; 1700^2 needs 982 bands) and, at 2400^2 (more cells than 1024 bands
    hold), the automatic switch to the global-atomic path; 1000^2 is the grid the literal Argoverse DISCRETIZATION=0.1 implies (argoverse_config.py:10)."""
    H, W = geomspec
    g = O.Geometry(boundary={"minX": 0, "maxX": 40, "minY": -20, "maxY": 20, "minZ": -2, "maxZ": 2},
                   BEV_HEIGHT=H, BEV_WIDTH=W, DISCRETIZATION=40 / H)
    if W != H:   # keep y inside the map: the y range follows W cells of size 40/H
        half = (40 / H) * W / 2
        g = O.Geometry(boundary={"minX": 0, "maxX": 40, "minY": -half, "maxY": half, "minZ": -2, "maxZ": 2},
                       BEV_HEIGHT=H, BEV_WIDTH=W, DISCRETIZATION=40 / H)
    sweeps = [O.synth_sweep(70 + i, 50000, g, k) for i, k in enumerate(["zties", "uniform", "outside"])]
    got, _ = _run_batch(cuda_device, sweeps, g)
    for i, s in enumerate(sweeps):
        _assert_bit_exact(got[i], O.make_bev_scatter(s, g, True, np.float32), "%dx%d frame %d" % (H, W, i))


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_max_height_not_a_power_of_two(cuda_device, algorithm):
    """maxZ - minZ = 3.0: the height plane needs the true fp32 division (for 4.0 / 8.0 the kernel
    multiplies by the exact reciprocal instead)."""
    g = O.Geometry(boundary={"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -1.0, "maxZ": 2.0})
    sweeps = [O.synth_sweep(90 + i, 60000, g, k) for i, k in enumerate(["uniform", "zties"])]
    got, _ = _run_batch(cuda_device, sweeps, g, algorithm=algorithm)
    for i, s in enumerate(sweeps):
        _assert_bit_exact(got[i], O.make_bev_scatter(s, g, True, np.float32), "max_h=3 frame %d" % i)
    ref = O.makeBEVMap(O.get_filtered_lidar(sweeps[0].copy(), g.boundary), g.boundary, g)
    _assert_bit_exact(got[0], ref.astype(np.float32), "max_h=3 vs lexsort port")


@pytest.mark.parametrize("algorithm", TILED_ALGOS)
def test_bev_bucket_overflow_paths(cuda_device, algorithm):
    """A band's bucket holds 8x the even share of a sweep; what does not fit goes to the frame's overflow list.
    Sweeps concentrated in one band / two bands / one cell overflow massively; mixed with ordinary frames in one batch
    (the overflow counters are per ring frame and reset per chunk), run twice on the same workspace."""
    rng = np.random.default_rng(11)
    b = O.KITTI.boundary

    def strip(n, x_lo, x_hi, zq=8):
        return np.stack([rng.uniform(x_lo, x_hi, n), rng.uniform(b["minY"], b["maxY"], n),
                         np.round(rng.uniform(b["minZ"], b["maxZ"], n) * zq) / zq, rng.uniform(0, 1, n)], 1).astype(np.float32)

    sweeps = [strip(120000, 20.0, 20.7),                                   # one band: ~105k records overflow
              O.synth_sweep(901, 120000, O.KITTI, "uniform"),
              np.concatenate([strip(60000, 3.0, 3.5), strip(60000, 44.0, 44.6)]),   # two overflowing bands
              O.synth_sweep(902, 100000, O.KITTI, "onecell"),
              strip(120000, 0.0, 6.0, zq=2),                               # 8 bands, each about at its bucket's capacity
              O.synth_sweep(903, 50000, O.KITTI, "zties")]
    for attempt in range(2):
        got, _ = _run_batch(cuda_device, sweeps, O.KITTI, algorithm=algorithm)
        for i, s in enumerate(sweeps):
            _assert_bit_exact(got[i], O.make_bev_scatter(s, O.KITTI, True, np.float32), "overflow frame %d" % i)


@pytest.mark.parametrize("algorithm", TILED_ALGOS)
def test_bev_crowded_band_streams_records(cuda_device, algorithm):
    """More records in one band than its threads hold in registers (8 x 512): the band kernel
    re-reads them from L2 once per phase.  Also a single cell holding > 63 points (LUT saturation)."""
    rng = np.random.default_rng(3)
    n = 90000
    b = O.KITTI.boundary
    sweep = np.stack([rng.uniform(10.0, 10.7, n), rng.uniform(b["minY"], b["maxY"], n),
                      np.round(rng.uniform(b["minZ"], b["maxZ"], n) * 16) / 16, rng.uniform(0, 1, n)], 1).astype(np.float32)
    sweep[:500, 0] = 10.3
    sweep[:500, 1] = 3.21
    got, _ = _run_batch(cuda_device, [sweep], O.KITTI, algorithm=algorithm)
    _assert_bit_exact(got[0], O.make_bev_scatter(sweep, O.KITTI, True, np.float32), "crowded band")


def test_bev_negative_zero_and_nan_payload_bits(cuda_device):
    """Without the filter z is stored as given: a cell whose winner is -0.0 must output -0.0 / max_h
    (= -0.0) like the reference, and +0.0 / -0.0 tie (lowest index wins)."""
    sweep = np.zeros((6, 4), np.float32)
    sweep[:, 0] = [1.0, 1.0, 5.0, 5.0, 9.0, 9.0]
    sweep[:, 1] = 0.5
    sweep[:, 2] = [-0.0, 0.0, 0.0, -0.0, -1.0, -0.0]
    sweep[:, 3] = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]
    for algorithm in (TILED, TWO_KERNEL, ATOMIC):
        got, _ = _run_batch(cuda_device, [sweep], O.KITTI, apply_filter=False, algorithm=algorithm)
        want = O.makeBEVMap(sweep, O.KITTI.boundary, O.KITTI).astype(np.float32)
        assert np.array_equal(got[0], want)
        if algorithm != ATOMIC:   # the tiled path also keeps the sign bit of a zero winner
            _assert_bit_exact(got[0], want, "signed zero")


def test_bev_out_of_map_is_counted_not_written(cuda_device):
    """Unfiltered points beyond the map make the reference raise IndexError; the kernel skips and
    counts them, and the drop-in wrapper turns the count into IndexError."""
    sweep = O.synth_sweep(2, 1000, O.KITTI, "uniform")
    sweep[:10, 0] = 500.0  # row 6080
    got, rast = _run_batch(cuda_device, [sweep], O.KITTI, apply_filter=False)
    assert rast.out_of_map_points() == 10
    want = O.make_bev_scatter(sweep[10:], O.KITTI, apply_filter=False, dtype=np.float32)
    # indices shift by 10 but winners are the same points
    _assert_bit_exact(got[0], want, "skip out-of-map")
    with pytest.raises(IndexError):
        pkg("data_process.kitti_bev_utils").makeBEVMap(sweep, O.KITTI.boundary)


def test_bev_full_size_properties(cuda_device):
    """Size-independent properties at BASELINE's full batch (64 x 120k): (1) permutation invariance
    of everything but tie order is covered elsewhere; here: density plane decodes to counts whose sum
    equals the number of kept, un-cropped points; occupied cells agree across the three planes;
    re-running is idempotent (bitwise identical output)."""
    sweeps = [O.synth_sweep(1000 + i, 120000, O.KITTI, "uniform") for i in range(64)]
    got1, rast = _run_batch(cuda_device, sweeps, O.KITTI)
    got2, _ = _run_batch(cuda_device, sweeps, O.KITTI)
    assert np.array_equal(got1.view(np.uint32), got2.view(np.uint32))
    lut = O.density_lut64().astype(np.float32)
    for i in (0, 31, 63):
        counts = np.searchsorted(lut, got1[i, 2])
        assert counts.max() < 63
        rows, cols, win, cnt, _ = O.bev_cell_selection(sweeps[i], O.KITTI, True)
        assert counts.sum() == cnt.sum()
        assert np.array_equal(got1[i, 2] > 0, counts > 0)
        # intensity is exactly the winner's 4th column
        assert np.array_equal(got1[i, 0][rows, cols], sweeps[i][win, 3])


def test_host_pipeline_and_dropin_makeBEVMap(cuda_device):
    """Host-buffer C ABI (sfa_pipeline_bev_host) and the reference-signature wrappers:
    makeBEVMap returns the reference's float64 map bit for bit."""
    bevmod = pkg("data_process.kitti_bev_utils")
    datmod = pkg("data_process.kitti_data_utils")
    sweep = O.synth_sweep(12, 120000, O.KITTI, "outside")
    filt = datmod.get_filtered_lidar(sweep, O.KITTI.boundary)
    want_f = O.get_filtered_lidar(sweep.copy(), O.KITTI.boundary)
    assert filt.dtype == np.float32 and np.array_equal(filt.view(np.uint32), want_f.view(np.uint32))
    m = bevmod.makeBEVMap(filt, O.KITTI.boundary)
    want = O.makeBEVMap(want_f, O.KITTI.boundary, O.KITTI)
    assert m.dtype == np.float64 and m.shape == (3, 608, 608)
    assert np.array_equal(m.view(np.uint64), want.view(np.uint64))
    fused = bevmod.makeBEVMap_from_raw(sweep, O.KITTI.boundary)
    _assert_bit_exact(fused, want.astype(np.float32), "fused raw")
    # batched host pipeline, ragged, more frames than one chunk
    fast = pkg("fast")
    pl = fast.HostPipeline(_geom(O.KITTI), max_frames=20, max_points=130000, C=0, h=1, w=1, K=1)
    sweeps = [O.synth_sweep(40 + i, 100000 + 1000 * i, O.KITTI, "zties") for i in range(20)]
    lens = [s.shape[0] for s in sweeps]
    out, bad = pl.bev(np.concatenate(sweeps), np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
    assert bad == 0
    for i, s in enumerate(sweeps):
        _assert_bit_exact(out[i], O.make_bev_scatter(s, O.KITTI, True, np.float32), "host frame %d" % i)
    pl.close()


@pytest.mark.parametrize("algorithm", ALGOS)
def test_bev_uniform_batch_without_offsets(cuda_device, algorithm):
    """offsets = NULL: every sweep holds exactly max_points points (the batched API's uniform form)."""
    fast = pkg("fast")
    B, N = 5, 30000
    sweeps = [O.synth_sweep(600 + i, N, O.KITTI, k) for i, k in enumerate(["uniform", "zties", "outside", "bounds", "clustered"])]
    pts = torch.from_numpy(np.stack(sweeps)).to(cuda_device)
    rast = fast.BevRasterizer(_geom(O.KITTI, algorithm=algorithm), max_batch=B, max_points=N, device=cuda_device)
    got = rast.rasterize_uniform(pts).cpu().numpy()
    for i, s in enumerate(sweeps):
        _assert_bit_exact(got[i], O.make_bev_scatter(s, O.KITTI, True, np.float32), "uniform frame %d" % i)
    with pytest.raises(ValueError):
        rast(pts.reshape(-1, 4)[:-1], None, N)


def test_bev_concurrent_streams_share_no_state(cuda_device):
    """Two rasterisers (own workspaces) driven concurrently from two CUDA streams, repeatedly: the library keeps
    no hidden global device state, so both results stay bit-exact."""
    fast = pkg("fast")
    sweeps_a = [O.synth_sweep(700 + i, 90000, O.KITTI, "zties") for i in range(6)]
    sweeps_b = [O.synth_sweep(800 + i, 70000, O.KITTI, "clustered") for i in range(6)]
    ra = fast.BevRasterizer(_geom(O.KITTI), max_batch=6, max_points=90000, device=cuda_device)
    rb = fast.BevRasterizer(_geom(O.KITTI), max_batch=6, max_points=70000, device=cuda_device)
    pa = torch.from_numpy(np.stack(sweeps_a)).to(cuda_device)
    pb = torch.from_numpy(np.stack(sweeps_b)).to(cuda_device)
    sa, sb = torch.cuda.Stream(device=cuda_device), torch.cuda.Stream(device=cuda_device)
    torch.cuda.synchronize()
    for _ in range(5):
        with torch.cuda.stream(sa):
            oa = ra.rasterize_uniform(pa)
        with torch.cuda.stream(sb):
            ob = rb.rasterize_uniform(pb)
    torch.cuda.synchronize()
    for i in range(6):
        _assert_bit_exact(oa[i].cpu().numpy(), O.make_bev_scatter(sweeps_a[i], O.KITTI, True, np.float32), "stream a %d" % i)
        _assert_bit_exact(ob[i].cpu().numpy(), O.make_bev_scatter(sweeps_b[i], O.KITTI, True, np.float32), "stream b %d" % i)


def test_bev_and_decode_in_a_cuda_graph(cuda_device):
    """The device API allocates nothing and never synchronises: a whole BEV + decode + post-process step captures
    into one CUDA graph and replays with new inputs in the same buffers."""
    fast = pkg("fast")
    B, N = 4, 50000
    rast = fast.BevRasterizer(_geom(O.KITTI), max_batch=B, max_points=N, device=cuda_device)
    pts = torch.empty((B, N, 4), dtype=torch.float32, device=cuda_device)
    heads = [torch.empty_like(t, device=cuda_device) for t in O.synth_heads(0, B=B)]
    bev = torch.empty((B, 3, 608, 608), dtype=torch.float32, device=cuda_device)
    det = torch.empty((B, 50, 10), dtype=torch.float32, device=cuda_device)
    pp = (torch.empty((B, 50, 8), device=cuda_device), torch.empty((B, 50), dtype=torch.int32, device=cuda_device),
          torch.empty((B, 50), dtype=torch.uint8, device=cuda_device))
    ws = fast.DecodeWorkspace(cuda_device, B, 3, 152, 152, 50)

    def step():
        rast.rasterize_uniform(pts, out=bev)
        fast.decode_device(*heads, K=50, out=det, workspace=ws)
        fast.post_process_dense(det, out=pp)

    def load(seed):
        sweeps = [O.synth_sweep(seed * 10 + i, N, O.KITTI, "zties") for i in range(B)]
        hh = O.synth_heads(seed, B=B, tie_free=True)
        pts.copy_(torch.from_numpy(np.stack(sweeps)))
        for dst, src in zip(heads, hh):
            dst.copy_(src)
        return sweeps, hh

    load(1)
    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for seed in (2, 3):
        sweeps, hh = load(seed)
        g.replay()
        torch.cuda.synchronize()
        for i in range(B):
            _assert_bit_exact(bev[i].cpu().numpy(), O.make_bev_scatter(sweeps[i], O.KITTI, True, np.float32), "graph %d" % i)
        want = O.decode(*[t.clone() for t in hh], K=50).numpy()
        assert np.array_equal(det.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("d", [50 / 608, 100 / 608, 0.1, 0.2, 0.3, 40 / 1000, 1.0, 3.0, 1e-3, 255.0, 4.0, 3.9, 6.0, 0.5, 97.0])
def test_exact_division_matches_div_rn_over_every_float_in_range(cuda_device, d):
    """The hoisted-reciprocal division of bev_bin against the compiler's IEEE div.rn for EVERY float in
    [2^-30, 2^12) and its negative (2.8e8 bit patterns per sign, covers every coordinate a map can index), plus a
    stretch of denormals / tiny values and of huge values / inf / NaN where the routine must fall back."""
    import ctypes
    lib = pkg("_lib").load()
    bad = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    d32 = float(np.float32(d))
    lo, hi = np.float32(2.0 ** -30).view(np.uint32), np.float32(2.0 ** 12).view(np.uint32)
    for first, count in ((int(lo), int(hi) - int(lo)), (0, 1 << 24), (int(np.float32(2.0 ** 90).view(np.uint32)), (0x7FC00010 - int(np.float32(2.0 ** 90).view(np.uint32))))):
        pkg("_lib").check(lib.sfa_selftest_division(d32, first, count, ctypes.c_void_p(bad.data_ptr()), stream))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_bev_random_geometries(cuda_device):
    """40 random map geometries (size, cell size, boundary placement incl. negative x / asymmetric y, max height) x
    random sweeps that partly leave the boundary: exercises band plans with tiny / partial last bands, the
    proven-in-range fast path vs the per-point range test, the exact-division and the true-division height paths."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        H = int(rng.choice([8, 20, 64, 100, 152, 304, 400, 608, 700]))
        W = int(rng.choice([8, 36, 64, 152, 304, 608, 640]))
        if (H * W) % 4:
            W += 4 - (W % 4) if (H % 4) else 0
        size_x = float(rng.choice([10.0, 25.0, 50.0, 80.0]))
        d = size_x / H
        min_x = float(rng.choice([0.0, 0.0, -size_x, -size_x / 2, 3.0]))
        half_y = d * W / 2 * float(rng.choice([1.0, 1.0, 0.7]))      # 0.7: the filter keeps y well inside the map
        min_z, max_z = float(rng.choice([-2.73, -3.0, -1.0])), float(rng.choice([1.27, 5.0, 2.0]))
        boundary = {"minX": min_x, "maxX": min_x + size_x, "minY": -half_y, "maxY": half_y, "minZ": min_z, "maxZ": max_z}
        if min_x < 0 and -min_x / d > H:       # negative rows would wrap beyond the map: the reference raises there
            continue
        g = O.Geometry(boundary=boundary, BEV_HEIGHT=H, BEV_WIDTH=W, DISCRETIZATION=d)
        n = int(rng.integers(1, 20000))
        sweep = O.synth_sweep(int(rng.integers(0, 1 << 30)), n, g, str(rng.choice(["uniform", "outside", "zties", "bounds", "gridaligned"])))
        try:
            want = O.make_bev_scatter(sweep, g, True, np.float32)
        except IndexError:
            continue                            # the reference would raise for this geometry
        got, rast = _run_batch(cuda_device, [sweep], g)
        _assert_bit_exact(got[0], want, "trial %d H=%d W=%d d=%g minX=%g" % (trial, H, W, d, min_x))
        assert rast.out_of_map_points() == 0


def test_front_and_back_pair_of_the_two_sides_demo(cuda_device):
    """demo_dataset.py:70-88: one sweep around the sensor, rasterised with cnf.boundary and with
    cnf.boundary_back (rows of negative x wrap; points with x == 0 land in both maps)."""
    fast = pkg("fast")
    rng = np.random.default_rng(77)
    sweeps = []
    for n in (60000, 1, 33333):
        pts = np.empty((n, 4), dtype=np.float32)
        pts[:, 0] = rng.uniform(-55, 55, n)
        pts[:, 1] = rng.uniform(-27, 27, n)
        pts[:, 2] = np.round(rng.uniform(-3, 1.5, n) * 8) / 8
        pts[:, 3] = rng.uniform(0, 1, n)
        pts[: n // 50, 0] = 0.0
        pts[n // 50: n // 25, 0] = -0.0
        sweeps.append(pts)
    lens = [s.shape[0] for s in sweeps]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda_device)
    pts = torch.from_numpy(np.concatenate(sweeps)).to(cuda_device)
    pair = fast.FrontBackRasterizer(max_batch=3, max_points=max(lens), device=cuda_device)
    front, back = pair(pts, offsets, max(lens))
    for i, s in enumerate(sweeps):
        _assert_bit_exact(front[i].cpu().numpy(), O.make_bev_scatter(s, O.KITTI, True, np.float32), "front %d" % i)
        _assert_bit_exact(back[i].cpu().numpy(), O.make_bev_scatter(s, O.KITTI_BACK, True, np.float32), "back %d" % i)
    assert float(back[0, 2, 0].sum()) > 0 and float(front[0, 2, 0].sum()) > 0      # the x == 0 points, row 0 of both
    assert pair.front.out_of_map_points() == 0


def test_bev_extras_augmentation_prologue_and_hflip(cuda_device):
    """sfa_bev_rasterize_ex: Random_Rotation + Random_Scaling applied to the points inside the raster kernel (no augmented
    sweep in HBM) and the horizontal flip as a mirrored column index.  Oracle: the reference's order of operations —
    augment the sweep (data_process/transformation.py:349-352, :366-368), get_filtered_lidar + makeBEVMap, then
    torch.flip(bev_map, [-1]) (data_process/kitti_dataset.py:93-97)."""
    fast = pkg("fast")
    rng = np.random.default_rng(21)
    kinds = ["outside", "zties", "uniform", "clustered", "bounds", "outside", "gridaligned"]
    lens = [30000, 1, 0, 45001, 20000, 120000, 7777]
    sweeps = [O.synth_sweep(700 + i, n, O.KITTI, k) if n else np.zeros((0, 4), np.float32) for i, (n, k) in enumerate(zip(lens, kinds))]
    B = len(sweeps)
    angles = rng.uniform(-np.pi / 4, np.pi / 4, B)
    factors = rng.uniform(0.95, 1.05, B).astype(np.float32)
    flips = np.array([0, 1, 1, 0, 1, 1, 0], dtype=np.uint8)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda_device)
    pts = torch.from_numpy(np.concatenate(sweeps)).to(cuda_device)
    mats = torch.from_numpy(np.stack([np.stack(O.transform_matrices(0, 0, 0, rz=a)) for a in angles])).to(cuda_device)
    scales = torch.from_numpy(factors).to(cuda_device)
    hflip = torch.from_numpy(flips).to(cuda_device)
    rast = fast.BevRasterizer(_geom(O.KITTI), max_batch=B, max_points=max(lens), device=cuda_device)
    before = pts.clone()
    got = rast(pts, offsets, max(lens), mats=mats, scales=scales, hflip=hflip).cpu().numpy()
    assert torch.equal(pts, before)   # the sweep itself is not modified
    for i, s in enumerate(sweeps):
        aug = O.random_scaling_points(O.random_rotation_points(s.copy(), angles[i]), factors[i]) if s.shape[0] else s
        want = O.make_bev_scatter(aug, O.KITTI, True, np.float32)
        if flips[i]:
            want = np.ascontiguousarray(np.flip(want, -1))
        _assert_bit_exact(got[i], want, "augmented frame %d" % i)
    # each extra alone; hflip only == np.flip of the plain map
    plain = rast(pts, offsets, max(lens)).cpu().numpy()
    only_flip = rast(pts, offsets, max(lens), hflip=torch.ones(B, dtype=torch.uint8, device=cuda_device)).cpu().numpy()
    assert np.array_equal(only_flip.view(np.uint32), np.flip(plain, -1).view(np.uint32))
    only_rot = rast(pts, offsets, max(lens), mats=mats).cpu().numpy()
    _assert_bit_exact(only_rot[3], O.make_bev_scatter(O.random_rotation_points(sweeps[3].copy(), angles[3]), O.KITTI, True, np.float32), "rotation only")
    assert rast.out_of_map_points() == 0


def test_bev_front_back_single_pass_many_frames(cuda_device):
    """Front + back maps from one read of each sweep, more sweeps than half the workspace ring (several chunks), uniform
    batch form, and a second call on the same workspace."""
    fast = pkg("fast")
    rng = np.random.default_rng(5)
    B, N = 41, 9000
    pts_np = np.empty((B, N, 4), np.float32)
    pts_np[..., 0] = rng.uniform(-52, 52, (B, N)); pts_np[..., 1] = rng.uniform(-26, 26, (B, N))
    pts_np[..., 2] = np.round(rng.uniform(-3, 1.5, (B, N)) * 4) / 4; pts_np[..., 3] = rng.uniform(0, 1, (B, N))
    pts = torch.from_numpy(pts_np).to(cuda_device)
    pair = fast.FrontBackRasterizer(max_batch=B, max_points=N, device=cuda_device)
    for attempt in range(2):
        front, back = pair(pts.reshape(-1, 4), None, N)
        f, b = front.cpu().numpy(), back.cpu().numpy()
        for i in (0, 7, 15, 16, 17, 31, 32, 40):
            _assert_bit_exact(f[i], O.make_bev_scatter(pts_np[i], O.KITTI, True, np.float32), "front %d" % i)
            _assert_bit_exact(b[i], O.make_bev_scatter(pts_np[i], O.KITTI_BACK, True, np.float32), "back %d" % i)


def test_stream8192_shard_frames(cuda_device):
    """BASELINE config[4] (bench.py --config stream8192): the sweeps of the 8192-frame stream are generated on the device
    from their frame number and sharded frame i -> rank i mod G.  Frames of one rank's shard: reproducible from the
    frame id alone (whatever the rank / batch position), and their maps bit-exact against the oracle."""
    import bench
    fast, sharding = pkg("fast"), pkg("sharding")
    shard = list(sharding.shard_range(8192, 1, 4, "cyclic"))
    assert shard[:3] == [1, 5, 9] and len(shard) == 2048
    ids = shard[:3] + [shard[-1]]
    N = 120000
    pts = bench.stream_frames_device(ids, N, cuda_device, torch)
    again = bench.stream_frames_device([ids[-1], ids[0]], N, cuda_device, torch)
    assert torch.equal(again[0], pts[-1]) and torch.equal(again[1], pts[0])
    rast = fast.BevRasterizer(_geom(O.KITTI), max_batch=len(ids), max_points=N, device=cuda_device)
    got = rast.rasterize_uniform(pts).cpu().numpy()
    host = pts.cpu().numpy()
    b = O.KITTI.boundary
    assert host[..., 0].min() >= b["minX"] and host[..., 0].max() <= b["maxX"] and host[..., 2].min() >= b["minZ"] - 1e-6
    for i in range(len(ids)):
        _assert_bit_exact(got[i], O.make_bev_scatter(host[i], O.KITTI, True, np.float32), "stream frame %d" % ids[i])


def test_bev_internal_lanes_give_identical_maps(cuda_device):
    """sfa_bev_set_internal_lanes: the 8-frame chunks of one call on 1, 2 or 3 library-owned streams (fork / join around the
    caller's stream) — identical maps, ragged batch with empty sweeps, several calls on one workspace, and inside a CUDA graph."""
    fast, lib = pkg("fast"), pkg("_lib").load()
    rng = np.random.default_rng(9)
    lens = [int(rng.integers(0, 40000)) if i % 5 else 0 for i in range(37)]
    sweeps = [O.synth_sweep(800 + i, n, O.KITTI, KINDS[i % len(KINDS)]) if n else np.zeros((0, 4), np.float32) for i, n in enumerate(lens)]
    pts = torch.from_numpy(np.concatenate(sweeps)).to(cuda_device)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=cuda_device)
    want = [O.make_bev_scatter(s, O.KITTI, True, np.float32) for s in sweeps]
    try:
        for n_lanes in (1, 2, 3, 0):
            assert lib.sfa_bev_set_internal_lanes(n_lanes) == 0
            rast = fast.BevRasterizer(_geom(O.KITTI, algorithm=TWO_KERNEL), max_batch=len(lens), max_points=max(lens), device=cuda_device)
            for attempt in range(2):
                got = rast(pts, offsets, max(lens)).cpu().numpy()
                for i in range(len(lens)):
                    _assert_bit_exact(got[i], want[i], "lanes=%d frame %d" % (n_lanes, i))
            out = torch.zeros((len(lens), 3, 608, 608), device=cuda_device)
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(device=cuda_device)
            s.wait_stream(torch.cuda.current_stream(cuda_device))
            with torch.cuda.graph(g, stream=s):
                rast(pts, offsets, max(lens), out=out)
            out.zero_()
            g.replay(); g.replay()
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy().view(np.uint32), np.stack(want).view(np.uint32)), n_lanes
            del g, rast
        assert lib.sfa_bev_set_internal_lanes(9) != 0      # out of range: refused
    finally:
        lib.sfa_bev_set_internal_lanes(0)


@pytest.mark.parametrize("algorithm", ALGOS)
@pytest.mark.parametrize("min_z", [-2.73, 0.0, -1e-30])
def test_bev_heights_at_and_just_above_min_z(cuda_device, algorithm, min_z):
    """bev_band may max-reduce the bits of the FINAL height z / max_height instead of an ordered key of z (filter on, power-of-two
    max_height, |minZ| not tiny: `hkey` in tiled_launch_chunk).  Heights of exactly 0 (z == minZ: the key of an occupied cell then
    equals the empty cell's), of one and a few ulps above minZ, alone in a cell and sharing one (ties on 0 included); with
    minZ = 0 or tiny the sweep holds denormal z values and the kernel has to take the ordered-key form."""
    g = O.Geometry(boundary={"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": min_z, "maxZ": min_z + 4.0})
    rng = np.random.default_rng(5)
    n = 40000
    s = O.synth_sweep(123, n, g, "uniform")
    mz = np.float32(min_z)
    steps = [mz]
    for _ in range(4):
        steps.append(np.nextafter(steps[-1], np.float32(np.inf), dtype=np.float32))
    if min_z == 0.0 or abs(min_z) < 1e-20:
        steps += [np.float32(1e-45), np.float32(3e-45), np.float32(1e-39), np.float32(2e-38), mz + np.float32(1e-38)]
    zs = np.array(steps, np.float32)
    s[: n // 2, 2] = zs[rng.integers(0, len(zs), n // 2)]
    # a quarter of those share cells: many points on few positions
    k = n // 8
    s[:k, 0] = np.float32(10.0) + (rng.integers(0, 40, k) * g.DISCRETIZATION).astype(np.float32)
    s[:k, 1] = np.float32(1.0)
    got, _ = _run_batch(cuda_device, [s, s[::-1].copy()], g, algorithm=algorithm)
    _assert_bit_exact(got[0], O.make_bev_scatter(s, g, True, np.float32), "min_z=%g" % min_z)
    _assert_bit_exact(got[1], O.make_bev_scatter(s[::-1].copy(), g, True, np.float32), "min_z=%g reversed" % min_z)
