"""GPU parity of stage B (_nms, _topk, decode, post_processing) against the CPU oracle and the
committed reference fixtures, through the C ABI (ctypes).

Bars: _nms output, top-K scores / spatial indices / classes bit-exact on tie-free inputs; with
ties (where torch.topk's order is implementation-defined) detections are compared after the
canonical within-tie ordering and the GPU order must follow its documented rule (lower class, then
lower y*w+x).  Gathered regression values are copies and the x/y sums are single fp32 adds, so
decoded boxes are bit-exact too (contract tolerance 1e-5 asserted as well).  post_processing's yaw
uses atan2f vs numpy's arctan2: tolerance 1e-5."""
import json
import os

import numpy as np
import pytest
import torch

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ev():
    return pkg("utils.evaluation_utils")


def _cuda(heads, dev):
    return tuple(t.to(dev) if t is not None else None for t in heads)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("seed,B", [(1, 1), (2, 4), (3, 64)])
def test_decode_tie_free_bit_exact(cuda_device, seed, B):
    heads = O.synth_heads(seed, B=B, tie_free=True)
    want, winds, wcls = O.decode(*[t.clone() for t in heads], K=50, return_inds=True)
    got = _ev().decode(*_cuda(heads, cuda_device), K=50)
    assert got.is_cuda and got.shape == (B, 50, 10) and got.dtype == torch.float32
    g = got.cpu().numpy()
    assert np.array_equal(_bits(g), _bits(want.numpy()))
    np.testing.assert_allclose(g, want.numpy(), rtol=1e-5, atol=0)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_decode_with_ties_canonical_order(cuda_device, seed):
    heads = O.synth_heads(seed, B=16, tie_free=False)
    want = O.decode(*[t.clone() for t in heads], K=50).numpy()
    g = _ev().decode(*_cuda(heads, cuda_device), K=50).cpu().numpy()
    assert np.array_equal(_bits(O.canonical_detections(g)), _bits(O.canonical_detections(want)))
    # the GPU's own order among equal scores: class ascending, then spatial index ascending
    for b in range(g.shape[0]):
        s, c = g[b, :, 0], g[b, :, 9]
        assert np.all(s[:-1] >= s[1:])
        eq = s[:-1] == s[1:]
        assert np.all(c[:-1][eq] <= c[1:][eq])


def test_nms_bit_exact_and_plateaus(cuda_device):
    hm = O.synth_heads(5, B=3)[0]
    hm[:, :, 4:7, 4:7] = 0.97           # plateau: every cell kept
    hm[:, :, 0, :] = 0.5                # constant border row
    hm[0, 0, 20, 20] = float("nan")     # max_pool2d propagates NaN; NaN == NaN is false -> 0 * ... = NaN
    want = O._nms(hm.clone()).numpy()
    got = _ev()._nms(hm.to(cuda_device)).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    got_cpu_in = _ev()._nms(hm)         # CPU tensor in -> CPU tensor out (device round trip inside)
    assert not got_cpu_in.is_cuda and np.array_equal(got_cpu_in.numpy(), want, equal_nan=True)


def test_topk_bit_exact_tie_free(cuda_device):
    hm = O.synth_heads(6, B=5, tie_free=True)[0]
    ws, wi, wc, wy, wx = O._topk(hm.clone(), K=50)
    gs, gi, gc, gy, gx = _ev()._topk(hm.to(cuda_device), K=50)
    assert gi.dtype == torch.int64 and gc.dtype == torch.int32
    assert torch.equal(gs.cpu(), ws) and torch.equal(gi.cpu(), wi) and torch.equal(gc.cpu(), wc)
    assert torch.equal(gy.cpu(), wy) and torch.equal(gx.cpu(), wx)


@pytest.mark.parametrize("K", [1, 7, 40, 100, 128])
def test_decode_other_K(cuda_device, K):
    heads = O.synth_heads(20 + K, B=2, tie_free=True)
    want = O.decode(*[t.clone() for t in heads], K=K).numpy()
    g = _ev().decode(*_cuda(heads, cuda_device), K=K).cpu().numpy()
    assert np.array_equal(_bits(g), _bits(want))


# (3, 40, 164): the peak-keep tile alone stays under the 48-KB default shared-memory limit, tile + static arrays do not
@pytest.mark.parametrize("C,h,w", [(3, 200, 200), (1, 8, 8), (3, 16, 40), (5, 33, 7), (3, 152, 152), (3, 40, 164), (4, 24, 148)])
def test_decode_other_shapes(cuda_device, C, h, w):
    """200x200 heads are what the Argoverse scripts feed decode (argoverse_test.py:669, 800x800 BEV)."""
    K = min(50, h * w // 16)   # >= K true peaks, so suppressed (zero, tied) cells never reach the top K
    heads = O.synth_heads(31, B=3, C=C, h=h, w=w, tie_free=True)
    want = O.decode(*[t.clone() for t in heads], K=K).numpy()
    g = _ev().decode(*_cuda(heads, cuda_device), K=K).cpu().numpy()
    assert np.array_equal(_bits(g), _bits(want))


def test_decode_without_offset_head(cuda_device):
    hm, off, d, z, dim = O.synth_heads(8, B=2, tie_free=True)
    want = O.decode(hm.clone(), None, d, z, dim, K=50).numpy()
    g = _ev().decode(hm.to(cuda_device), None, d.to(cuda_device), z.to(cuda_device), dim.to(cuda_device), K=50)
    assert np.array_equal(_bits(g.cpu().numpy()), _bits(want))


def test_decode_massive_ties_threshold_group(cuda_device):
    """A heat map with only 7 distinct values: the K-th value is shared by thousands of cells; the
    kernel must take exactly K and break ties toward lower class / lower index."""
    g = torch.Generator().manual_seed(4)
    hm = (torch.randint(1, 8, (2, 3, 152, 152), generator=g).float() / 8.0)
    _, off, d, z, dim = O.synth_heads(9, B=2)
    got = _ev().decode(hm.to(cuda_device), off.to(cuda_device), d.to(cuda_device), z.to(cuda_device),
                       dim.to(cuda_device), K=50).cpu().numpy()
    nms = O._nms(hm.clone())
    for b in range(2):
        flat = nms[b].reshape(-1)
        order = np.lexsort((np.arange(flat.numel()), -flat.numpy().astype(np.float64)))[:50]
        assert np.array_equal(got[b, :, 0], flat.numpy()[order])
        assert np.array_equal(got[b, :, 9].astype(np.int64), order // (152 * 152))
        sp = order % (152 * 152)
        assert np.array_equal(got[b, :, 1], (sp % 152).astype(np.float32) + off[b, 0].reshape(-1).numpy()[sp])
        assert np.array_equal(got[b, :, 2], (sp // 152).astype(np.float32) + off[b, 1].reshape(-1).numpy()[sp])


def test_decode_all_zero_and_negative_heat(cuda_device):
    """Raw (pre-sigmoid) heat maps are legal inputs too: negative values, zeros after NMS outrank
    negative peaks exactly as in torch (heat * keep gives +-0.0 for suppressed cells)."""
    g = torch.Generator().manual_seed(10)
    hm = torch.randn(2, 3, 40, 40, generator=g) - 3.0
    _, off, d, z, dim = O.synth_heads(10, B=2, h=40, w=40)
    want = O.decode(hm.clone(), off, d, z, dim, K=20).numpy()
    got = _ev().decode(*_cuda((hm, off, d, z, dim), cuda_device), K=20).cpu().numpy()
    # scores are all (+-)0.0 here (suppressed cells outrank every negative peak): compare values
    assert np.array_equal(got[:, :, 0], want[:, :, 0])
    assert np.array_equal(np.sort(got[:, :, 0], axis=1), np.sort(want[:, :, 0], axis=1))


def test_decode_K_larger_than_map_raises_like_topk(cuda_device):
    heads = O.synth_heads(2, B=1, C=3, h=4, w=4)
    with pytest.raises(RuntimeError):
        O.decode(*[t.clone() for t in heads], K=50)
    with pytest.raises(RuntimeError):
        _ev().decode(*_cuda(heads, cuda_device), K=50)


def test_decode_cpu_tensors_go_through_host_pipeline(cuda_device):
    """Most reference scripts pin tensors to the CPU (test.py:50): CPU in -> CPU out via
    sfa_pipeline_decode_host (H2D, kernel, D2H inside the library)."""
    heads = O.synth_heads(14, B=19, tie_free=True)   # 19 frames: more than two chunks
    want = O.decode(*[t.clone() for t in heads], K=50).numpy()
    got = _ev().decode(*heads, K=50)
    assert not got.is_cuda
    assert np.array_equal(_bits(got.numpy()), _bits(want))


def test_decode_golden_fixtures(cuda_device):
    z = np.load(os.path.join(GOLD, "decode_small.npz"))
    for cid in range(int(z["n_cases"])):
        tag = "d%d" % cid
        K = int(z[tag + "_K"])
        heads = tuple(torch.from_numpy(z["%s_%s" % (tag, n)].copy()) for n in ("hm", "off", "dir", "z", "dim"))
        got = _ev().decode(*_cuda(heads, cuda_device), K=K).cpu().numpy()
        want = z[tag + "_det"]
        assert np.array_equal(_bits(O.canonical_detections(got)), _bits(O.canonical_detections(want))), tag
        nms = _ev()._nms(heads[0].to(cuda_device)).cpu().numpy()
        assert np.array_equal(_bits(nms), _bits(z[tag + "_nms"])), tag
    with open(os.path.join(GOLD, "decode_hashes.json")) as f:
        import hashlib
        for c in json.load(f):
            heads = O.synth_heads(c["seed"], B=c["B"], tie_free=c["tie_free"])
            sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
            if sha(np.concatenate([t.numpy().ravel() for t in heads])) != c["input_sha256"]:
                continue
            got = _ev().decode(*_cuda(heads, cuda_device), K=c["K"]).cpu().numpy()
            assert sha(O.canonical_detections(got)) == c["det_canonical_sha256"]


def test_post_processing_matches_reference_semantics(cuda_device):
    heads = O.synth_heads(15, B=4, tie_free=True)
    det = O.decode(*[t.clone() for t in heads], K=50).numpy().astype(np.float32)
    want = O.post_processing(det.copy(), 3, 4, 0.2)
    got = _ev().post_processing(det.copy(), 3, 4, 0.2)
    assert len(got) == 4
    for w, g in zip(want, got):
        for j in range(3):
            wj = np.asarray(w[j], np.float32).reshape(-1, 8)
            assert g[j].shape == wj.shape and g[j].dtype == np.float32
            assert np.array_equal(_bits(g[j][:, :7]), _bits(wj[:, :7]))            # exact fp32 arithmetic
            np.testing.assert_allclose(g[j][:, 7], wj[:, 7], rtol=1e-5, atol=1e-6)  # atan2f vs np.arctan2
    assert _ev().post_processing(det[:0]) == []
    # golden: the reference's own outputs
    z = np.load(os.path.join(GOLD, "decode_small.npz"))
    det = z["d0_det"].astype(np.float32)
    got = _ev().post_processing(det, 3, 4, 0.2)
    for i in range(det.shape[0]):
        for j in range(3):
            wj = z["d0_pp_s%d_c%d" % (i, j)]
            assert np.array_equal(_bits(got[i][j][:, :7]), _bits(wj[:, :7]))
            np.testing.assert_allclose(got[i][j][:, 7], wj[:, 7], rtol=1e-5, atol=1e-6)
    real = _ev().convert_det_to_real_values(got[0])
    np.testing.assert_allclose(np.asarray(real, np.float64).reshape(-1, 8), z["d0_real_s0"], rtol=1e-5, atol=1e-6)


def test_decode_full_batch_properties(cuda_device):
    """BASELINE config[1] size (B=64): scores descending; every reported cell is a 3x3 local maximum
    of its class plane; idempotent; per-frame result independent of the batch it is in."""
    heads = O.synth_heads(16, B=64)
    dev = _cuda(heads, cuda_device)
    g1 = _ev().decode(*dev, K=50)
    g2 = _ev().decode(*dev, K=50)
    assert torch.equal(g1, g2)
    single = _ev().decode(*[t[37:38].contiguous() for t in dev], K=50)
    assert torch.equal(single[0], g1[37])
    g = g1.cpu().numpy()
    hm = heads[0].numpy()
    off = heads[1].numpy()
    assert np.all(g[:, :, 0][:, :-1] >= g[:, :, 0][:, 1:])
    for b in (0, 63):
        c = g[b, :, 9].astype(int)
        x = np.round(g[b, :, 1] - 0.5).astype(int)  # x + off, off in (0,1)
        y = np.round(g[b, :, 2] - 0.5).astype(int)
        for k in range(50):
            xs = [xx for xx in (x[k] - 1, x[k], x[k] + 1) if 0 <= xx < 152]
            found = False
            for xx in xs:
                for yy in (y[k] - 1, y[k], y[k] + 1):
                    if 0 <= yy < 152 and hm[b, c[k], yy, xx] == g[b, k, 0] and \
                            np.float32(xx) + off[b, 0, yy, xx] == g[b, k, 1]:
                        patch = hm[b, c[k], max(0, yy - 1):yy + 2, max(0, xx - 1):xx + 2]
                        found = found or patch.max() == g[b, k, 0]
            assert found


def test_decode_fused_sigmoid_tolerance(cuda_device):
    """`_sigmoid` (utils/torch_utils.py:44-45) applied to the raw hm_cen / cen_offset logits
    inside the decode kernels.  torch's sigmoid and expf-based sigmoid agree to an ulp or two, so this path is held to
    a tolerance (scores 2e-7 abs, boxes 1e-5) while the un-fused path stays bit-exact; detections whose score is not
    within 1e-6 of a neighbour must be the same cells."""
    fast = pkg("fast")
    g = torch.Generator().manual_seed(21)
    B = 8
    hm_raw = torch.randn(B, 3, 152, 152, generator=g) * 2.0 - 2.0
    hm_raw[:, :, 5:9, 5:9] = -20.0          # clamps to 1e-4: a plateau
    hm_raw[0, 1, 40, 40] = 30.0             # clamps to 1 - 1e-4
    off_raw = torch.randn(B, 2, 152, 152, generator=g)
    _, _, d, z, dim = O.synth_heads(22, B=B)
    want = O.decode(O._sigmoid(hm_raw.clone()), O._sigmoid(off_raw.clone()), d, z, dim, K=50).numpy()
    got = fast.decode_device(hm_raw.to(cuda_device), off_raw.to(cuda_device), d.to(cuda_device), z.to(cuda_device),
                             dim.to(cuda_device), K=50, apply_sigmoid=True).cpu().numpy()
    np.testing.assert_allclose(got[:, :, 0], want[:, :, 0], rtol=0, atol=2e-7)
    assert got[0, 0, 0] == np.float32(1 - 1e-4)
    checked = 0
    for b in range(B):
        s = want[b, :, 0].astype(np.float64)
        gap = np.minimum(np.abs(np.diff(s, prepend=np.inf)), np.abs(np.diff(s, append=-np.inf)))
        for k in np.flatnonzero(gap > 1e-6):
            np.testing.assert_allclose(got[b, k], want[b, k], rtol=1e-5, atol=1e-5)
            assert got[b, k, 9] == want[b, k, 9]
            checked += 1
    assert checked > B * 40
    # the flag off is the bit-exact path on already-activated heads
    exact = fast.decode_device(O._sigmoid(hm_raw.clone()).to(cuda_device), O._sigmoid(off_raw.clone()).to(cuda_device),
                               d.to(cuda_device), z.to(cuda_device), dim.to(cuda_device), K=50).cpu().numpy()
    assert np.array_equal(_bits(O.canonical_detections(exact)), _bits(O.canonical_detections(want)))


def test_convert_det_to_real_values_device_and_host(cuda_device):
    """convert_det_to_real_values (evaluation_utils.py:177-193), dense on the device and
    as the reference-signature host function, against the reference's own output (golden) and the oracle."""
    fast = pkg("fast")
    z = np.load(os.path.join(GOLD, "decode_small.npz"))
    det = z["d0_det"].astype(np.float32)
    pp = _ev().post_processing(det, 3, 4, 0.2)
    host = _ev().convert_det_to_real_values(pp[0])
    np.testing.assert_allclose(np.asarray(host, np.float64).reshape(-1, 8), z["d0_real_s0"], rtol=1e-6, atol=1e-6)
    rows, cls, keep, real = fast.post_process_dense(torch.from_numpy(det).to(cuda_device), real=True)
    real, cls, keep = real.cpu().numpy(), cls.cpu().numpy(), keep.cpu().numpy()
    dense = np.concatenate([real[0][(cls[0] == j) & keep[0]] for j in range(3)], 0)
    np.testing.assert_allclose(dense.astype(np.float64), z["d0_real_s0"], rtol=1e-6, atol=1e-6)
    heads = O.synth_heads(33, B=5, tie_free=True)
    det5 = O.decode(*[t.clone() for t in heads], K=50).numpy().astype(np.float32)
    rows, cls, keep, real = fast.post_process_dense(torch.from_numpy(det5).to(cuda_device), real=True)
    ref_pp = O.post_processing(det5)
    for i in range(5):
        want = np.asarray(O.convert_det_to_real_values(ref_pp[i]), np.float64).reshape(-1, 8)
        r, c, k = real[i].cpu().numpy(), cls[i].cpu().numpy(), keep[i].cpu().numpy()
        got = np.concatenate([r[(c == j) & k] for j in range(3)], 0)
        np.testing.assert_allclose(got.astype(np.float64), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,h,w,K", [(3, 152, 152, 50), (3, 200, 200, 128), (1, 40, 600, 30), (5, 33, 70, 17), (2, 64, 512, 50)])
def test_decode_plateau_maps_exact_tie_order(cuda_device, C, h, w, K):
    """Clamped-sigmoid style maps: a floor plateau everywhere (every plateau cell is a kept peak), a few cells above it —
    fewer than K, so the top K must be completed with the LOWEST-index plateau cells — plus a second, higher plateau.
    The candidate kernel lists only the K lowest-index floor cells per column group (the rest can never be selected);
    the result must equal the exact (score desc, class asc, index asc) order.  Shapes cover one / several column
    groups, rows wider than the cap bitmap (cap off), and K up to 128."""
    g = torch.Generator().manual_seed(C * 1000 + w)
    B = 3
    hm = torch.full((B, C, h, w), 1e-4)
    n_peaks = max(1, K // 3)
    for b in range(B):
        idx = torch.randperm(C * h * w, generator=g)[:n_peaks]
        hm[b].view(-1)[idx] = torch.rand(n_peaks, generator=g) * 0.5 + 0.3
    hm[:, C - 1, h // 2:h // 2 + 3, 2:7] = 0.25          # a second plateau above the floor
    _, off, d, z, dim = O.synth_heads(5, B=B, C=C, h=h, w=w)
    got = _ev().decode(*_cuda((hm, off, d, z, dim), cuda_device), K=K).cpu().numpy()
    nms = O._nms(hm.clone())
    for b in range(B):
        flat = nms[b].reshape(-1).numpy()
        order = np.lexsort((np.arange(flat.size), -flat.astype(np.float64)))[:K]
        assert np.array_equal(got[b, :, 0], flat[order])
        assert np.array_equal(got[b, :, 9].astype(np.int64), order // (h * w))
        sp = order % (h * w)
        assert np.array_equal(got[b, :, 1], (sp % w).astype(np.float32) + off[b, 0].reshape(-1).numpy()[sp])
        assert np.array_equal(got[b, :, 2], (sp // w).astype(np.float32) + off[b, 1].reshape(-1).numpy()[sp])


def test_error_paths_report_status_and_message(cuda_device):
    """The C ABI never throws and never corrupts: bad calls return a negative status with a message."""
    import ctypes
    L = pkg("_lib")
    lib = L.load()
    fast = pkg("fast")
    heads = _cuda(O.synth_heads(1, B=2, tie_free=True), cuda_device)
    with pytest.raises(RuntimeError, match="K=129"):
        fast.decode_device(*heads, K=129)                      # K > 128 is not supported
    det = torch.empty((2, 50, 10), device=cuda_device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    small = torch.empty(1024, dtype=torch.uint8, device=cuda_device)
    ws = small.data_ptr() + (-small.data_ptr()) % 256
    rc = lib.sfa_decode(p(heads[0]), p(heads[1]), p(heads[2]), p(heads[3]), p(heads[4]), 2, 3, 152, 152, 50, p(det), None, 0,
                        ctypes.c_void_p(ws), 512, stream)
    assert rc == -2 and "workspace too small" in L.last_error()   # SFA_ERR_WORKSPACE_TOO_SMALL
    rc = lib.sfa_decode(p(heads[0]), p(heads[1]), p(heads[2]), p(heads[3]), p(heads[4]), 2, 3, 152, 152, 50, p(det), None, 0,
                        ctypes.c_void_p(ws + 4), 1 << 30, stream)
    assert rc == -1 and "aligned" in L.last_error()
    # BEV: workspace sized for fewer points than the call declares
    geom = pkg("geometry").from_config(pkg("config.kitti_config"))
    rast = fast.BevRasterizer(geom, max_batch=2, max_points=1000, device=cuda_device)
    pts = torch.zeros((2, 5000, 4), device=cuda_device)
    with pytest.raises(ValueError):
        rast.rasterize_uniform(pts)
    lut = torch.from_numpy(geom.lut32.copy()).to(cuda_device)
    out = torch.empty((2, 3, 608, 608), device=cuda_device)
    rc = lib.sfa_bev_rasterize(p(pts), None, 2, 5000, ctypes.byref(geom.params), p(lut), p(out), None,
                               ctypes.c_void_p(rast._ws_ptr), rast._ws_bytes, stream)
    assert rc == -2 and "workspace too small" in L.last_error()
    torch.cuda.synchronize()
    # and the library is still healthy afterwards
    got = fast.decode_device(*heads, K=50).cpu().numpy()
    want = O.decode(*[t.cpu().clone() for t in heads], K=50).numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_decode_with_nan_and_inf_cells(cuda_device):
    """NaN / +inf / -inf in the heat map (evaluation_utils.py:21-26 semantics: NaN propagates through
    max_pool2d, `-inf * 0` is NaN) — tiles holding such values take the scalar walker, the other tiles of
    the same frame the 4x4-cells-per-thread one; both must agree with the reference's torch ops.  NaN
    scores rank first in torch.topk in an unspecified order, so those rows are compared as a set."""
    heads = list(O.synth_heads(44, B=3, tie_free=True))
    hm = heads[0]
    hm[0, 0, 20, 20] = float("nan")          # slab 1
    hm[0, 2, 100, 7] = float("nan")          # slab 6, another class
    hm[0, 1, 60, 151] = float("inf")         # right edge
    hm[0, 1, 130, 0] = float("-inf")         # under larger neighbours -> -inf * 0 = NaN
    hm[1, 0, 0, 0] = float("inf")            # corner; frame 2 stays ordinary
    want = O.decode(*[t.clone() for t in heads], K=50).numpy()
    got = _ev().decode(*_cuda(heads, cuda_device), K=50).cpu().numpy()
    for b in range(3):
        w_nan, g_nan = np.isnan(want[b, :, 0]), np.isnan(got[b, :, 0])
        assert w_nan.sum() == g_nan.sum() == (3 if b == 0 else 0)
        assert np.array_equal(_bits(got[b][~g_nan]), _bits(want[b][~w_nan])), b
        key = lambda rows: rows[np.lexsort((rows[:, 2], rows[:, 1], rows[:, 9]))][:, 1:]
        assert np.array_equal(key(got[b][g_nan]), key(want[b][w_nan]))
    assert np.isinf(got[0, 3, 0]) and np.isinf(got[1, 0, 0])


def test_decode_with_fused_post_processing(cuda_device):
    """sfa_decode_post: the dense post_processing rows written by the decode's own epilogue equal the stand-alone
    sfa_post_process of the same detections bit for bit (same device function), detections unchanged."""
    fast = pkg("fast")
    heads = [t.to(cuda_device) for t in O.synth_heads(91, B=5)]
    B, K = 5, 50
    det0 = fast.decode_device(*heads, K=K)
    rows0, cls0, keep0, real0 = fast.post_process_dense(det0, real=True)
    det1 = torch.empty_like(det0)
    post = (torch.empty((B, K, 8), device=cuda_device), torch.empty((B, K), dtype=torch.int32, device=cuda_device),
            torch.empty((B, K), dtype=torch.uint8, device=cuda_device), torch.empty((B, K, 8), device=cuda_device))
    fast.decode_device(*heads, K=K, out=det1, post=post)
    torch.cuda.synchronize()
    assert torch.equal(det0, det1)
    assert torch.equal(rows0.view(torch.int32), post[0].view(torch.int32)) and torch.equal(cls0, post[1])
    assert torch.equal(keep0, post[2].bool()) and torch.equal(real0.view(torch.int32), post[3].view(torch.int32))
    ref = O.post_processing(det0.cpu().numpy().astype(np.float32))
    r, c, k = post[0].cpu().numpy(), post[1].cpu().numpy(), post[2].cpu().numpy().astype(bool)
    for i in range(B):
        for j in range(3):
            w = np.asarray(ref[i][j], np.float32).reshape(-1, 8)
            g = r[i][(c[i] == j) & k[i]]
            assert g.shape == w.shape and np.array_equal(g[:, :7], w[:, :7])
            np.testing.assert_allclose(g[:, 7], w[:, 7], rtol=1e-5, atol=1e-6)
