"""Generates the committed golden fixtures by running the REFERENCE's own functions.

    python tests/golden/make_golden.py          (build container only: needs /root/reference)

TEST INFRASTRUCTURE.  The reference ships no tests or golden vectors for this path,
so the pin is: the reference's unmodified functions (imported in place by oracle/ref_loader.py) executed on
seeded synthetic inputs, outputs stored here.  Nothing reads /root/reference at test time.

  bev_small.npz      explicit inputs + sparse reference outputs (float64 map, filtered sweep) for
                     small sweeps of every adversarial kind / geometry
  bev_hashes.json    sha256 of input and of the reference output for full-size sweeps (120k / 250k)
                     that the tests regenerate from the seed
  decode_small.npz   explicit heads + reference _nms/_topk/decode/post_processing/convert outputs
  decode_hashes.json sha256 pins at the full 152x152 head size, B=4, K=50
  bvfeature_small.npz  small sweeps of every kind -> sparse reference makeBVFeature output
                     (argoverse_test.py:199-254); bvfeature_hashes.json pins 250k-point sweeps
  augment_small.npz  sweep -> the reference's point_transform (float64), Random_Rotation and
                     Random_Scaling results (float32 sweeps, seeded np.random; data_process/transformation.py)
  projection_small.npz  post-processed detections + calibration -> the reference's lidar_to_camera_box
                     rows and convert_sfa3d_to_2d_boxes image boxes (test6.py:129-187), incl. boxes
                     behind the camera, straddling the image plane, outside the image and NaN rows
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
import sfa_oracle as O  # noqa: E402  (only its synthetic-input generators are used here)

KINDS = ["uniform", "outside", "zties", "gridaligned", "bounds", "nonfinite", "onecell", "clustered"]
GEOMS = {"kitti": O.KITTI, "kitti_back": O.KITTI_BACK, "argoverse": O.ARGOVERSE}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_bev(ns, sweep, geom):
    """get_filtered_lidar + makeBEVMap exactly as data_process/kitti_dataset.py:64-66 chains them."""
    with ref_loader.patched_geometry(ns, geom.boundary, geom.BEV_HEIGHT, geom.BEV_WIDTH, geom.DISCRETIZATION):
        filt = ns.get_filtered_lidar(sweep.copy(), geom.boundary)
        bev = ns.makeBEVMap(filt, geom.boundary)
    return filt, bev


def sparse(bev):
    flat = bev.reshape(-1)
    nz = np.flatnonzero(flat != 0)
    return nz.astype(np.int32), flat[nz]


def make_bev(ns):
    small, hashes = {}, []
    case = 0
    for gname, geom in GEOMS.items():
        for kind in KINDS:
            n = 3000
            sweep = O.synth_sweep(1000 + case, n, geom, kind)
            filt, bev = ref_bev(ns, sweep, geom)
            nz, val = sparse(bev)
            tag = "c%02d" % case
            small[tag + "_pts"] = sweep
            small[tag + "_filt"] = filt
            small[tag + "_nz"] = nz
            small[tag + "_val"] = val
            small[tag + "_meta"] = np.array([gname, kind])
            case += 1
    small["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "bev_small.npz"), **small)

    full = [("kitti", k, 120000, 2000 + i) for i, k in enumerate(KINDS) if k != "onecell"]
    full += [("kitti", "onecell", 100000, 2100), ("kitti_back", "uniform", 120000, 2200),
             ("kitti_back", "zties", 120000, 2201), ("argoverse", "uniform", 250000, 2300),
             ("argoverse", "zties", 250000, 2301), ("argoverse", "outside", 250000, 2302)]
    for gname, kind, n, seed in full:
        geom = GEOMS[gname]
        sweep = O.synth_sweep(seed, n, geom, kind)
        filt, bev = ref_bev(ns, sweep, geom)
        hashes.append({"geom": gname, "kind": kind, "n": n, "seed": seed, "input_sha256": sha(sweep),
                       "filtered_sha256": sha(filt), "filtered_rows": int(filt.shape[0]),
                       "bev_f64_sha256": sha(bev), "bev_f32_sha256": sha(bev.astype(np.float32)),
                       "occupied": int(np.count_nonzero(bev[2]))})
    with open(os.path.join(HERE, "bev_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    return case, len(hashes)


def ref_decode_bundle(ns, heads, K):
    hm, off, direction, z, dim = heads
    out = {}
    out["nms"] = ns._nms(hm.clone()).numpy()
    ts, ti, tc, ty, tx = ns._topk(ns._nms(hm.clone()), K=K)
    out["topk_score"], out["topk_inds"], out["topk_clses"] = ts.numpy(), ti.numpy(), tc.numpy()
    out["topk_ys"], out["topk_xs"] = ty.numpy(), tx.numpy()
    det = ns.decode(hm.clone(), off.clone(), direction, z, dim, K=K)
    out["det"] = det.numpy()
    out["det_nooff"] = ns.decode(hm.clone(), None, direction, z, dim, K=K).numpy()
    detn = det.numpy().astype(np.float32)
    per_sample = [ns.post_processing_live(detn[i:i + 1].copy(), 3, 4, 0.2)[0] for i in range(detn.shape[0])]
    pristine = ns.post_processing_pristine(detn.copy(), 3, 4, 0.2)
    live = ns.post_processing_live(detn.copy(), 3, 4, 0.2)
    for i, d in enumerate(per_sample):
        for j in range(3):
            out["pp_s%d_c%d" % (i, j)] = np.asarray(d[j], dtype=np.float32).reshape(-1, 8)
            assert np.array_equal(out["pp_s%d_c%d" % (i, j)],
                                  np.asarray(pristine[i][j], dtype=np.float32).reshape(-1, 8)), "pristine != per-sample live"
    assert len(live) == 1  # the live copy returns the last sample only (evaluation_utils.py:158)
    real = ns.convert_det_to_real_values(per_sample[0])
    out["real_s0"] = np.asarray(real, dtype=np.float64).reshape(-1, 8)
    return out


def make_decode(ns):
    small = {}
    cases = [(0, 2, 3, 24, 24, 10, False), (1, 3, 3, 16, 40, 20, True), (2, 1, 3, 152, 152, 50, True),
             (3, 2, 1, 8, 8, 5, False), (4, 2, 3, 24, 24, 10, "plateau")]
    for cid, B, C, h, w, K, mode in cases:
        heads = O.synth_heads(300 + cid, B=B, C=C, h=h, w=w, tie_free=(mode is True))
        if mode == "plateau":   # equal-valued neighbourhoods: every cell of a plateau is kept by _nms
            hm = heads[0]
            hm[:, :, 4:7, 4:7] = 0.97
            hm[:, 0, 10:12, 10:14] = 0.93
            hm[:, :, 0, 0] = 0.99
            hm[:, :, -1, -1] = 0.98
        tag = "d%d" % cid
        for name, t in zip(("hm", "off", "dir", "z", "dim"), heads):
            small["%s_%s" % (tag, name)] = t.numpy()
        small[tag + "_K"] = np.array(K)
        if C == 3:
            for k, v in ref_decode_bundle(ns, heads, K).items():
                small["%s_%s" % (tag, k)] = v
        else:
            hm, off, direction, z, dim = heads
            small[tag + "_det"] = ns.decode(hm.clone(), off.clone(), direction, z, dim, K=K).numpy()
            small[tag + "_nms"] = ns._nms(hm.clone()).numpy()
    small["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "decode_small.npz"), **small)

    hashes = []
    for seed, tie_free in ((400, True), (401, False), (402, True)):
        heads = O.synth_heads(seed, B=4, C=3, h=152, w=152, tie_free=tie_free)
        b = ref_decode_bundle(ns, heads, 50)
        tied = 0
        for i in range(4):
            s = b["det"][i, :, 0]
            tied += int(np.sum(s[1:] == s[:-1]))
        hashes.append({"seed": seed, "tie_free": tie_free, "B": 4, "K": 50,
                       "input_sha256": sha(np.concatenate([t.numpy().ravel() for t in heads])),
                       "nms_sha256": sha(b["nms"]), "det_sha256": sha(b["det"]),
                       "det_canonical_sha256": sha(O.canonical_detections(b["det"])),
                       "topk_inds_sha256": sha(b["topk_inds"]), "adjacent_equal_scores": tied})
    with open(os.path.join(HERE, "decode_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    return len(cases), len(hashes)


BV_GEOMS = O.BV_TEST_GEOMS
BV_KINDS = ["uniform", "adversarial", "xyz_only", "dark", "outside", "wide"]


def make_bvfeature(ns):
    small, c = {}, 0
    for gname, (disc, bnd) in BV_GEOMS.items():
        for kind in BV_KINDS:
            pts = O.synth_argoverse_sweep(700 + c, 3000 if gname == "argo" else 2000, kind, bnd)
            ref = ns.makeBVFeature(pts, disc, bnd)
            assert ref.dtype == np.float32
            tag = "c%02d" % c
            small[tag + "_meta"] = np.array([gname, kind])
            small[tag + "_pts"] = pts
            flat = ref.ravel().view(np.uint32)
            nz = np.flatnonzero(flat)
            small[tag + "_nz"], small[tag + "_val"] = nz.astype(np.int64), flat[nz]
            small[tag + "_shape"] = np.array(ref.shape)
            c += 1
    small["n_cases"] = np.array(c)
    np.savez_compressed(os.path.join(HERE, "bvfeature_small.npz"), **small)
    hashes = []
    for seed, kind, n in ((800, "uniform", 250000), (801, "adversarial", 250000), (802, "xyz_only", 120000)):
        pts = O.synth_argoverse_sweep(seed, n, kind)
        ref = ns.makeBVFeature(pts, O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY)
        hashes.append({"seed": seed, "kind": kind, "n": n, "input_sha256": sha(pts), "output_sha256": sha(ref),
                       "occupied": int((ref[0] > 0).sum())})
    with open(os.path.join(HERE, "bvfeature_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    return c, len(hashes)


AUG_PARAMS = [(0, 0, 0, 0, 0, 0.3), (1.5, -2.0, 0.25, 0, 0, -0.7), (0, 0, 0, 0.1, 0.2, 0.3), (0.1, 0.2, 0.3, 0, 0, 0),
              (0, 0, 0, 0, 0, np.pi / 4), (0, 0, 0, -1.3, 0, 0)]


def make_augment(ns):
    small = {}
    sweep = O.synth_sweep(600, 4000, O.KITTI, "outside")
    small["sweep"] = sweep
    small["params"] = np.array(AUG_PARAMS, dtype=np.float64)
    for k, prm in enumerate(AUG_PARAMS):
        small["pt%d" % k] = ns.point_transform(sweep[:, 0:3], *prm)
        small["pt%d_f64in" % k] = ns.point_transform(sweep[:, 0:3].astype(np.float64) * 1.000001, *prm)
    labels = np.array([[10.0, 1.0, -1.0, 1.5, 1.6, 4.0, 0.3], [30.0, -5.0, -0.8, 1.6, 1.7, 4.2, -1.0]])
    for k, seed in enumerate((1, 2, 3, 4)):
        np.random.seed(seed)
        lidar, _ = ns.Random_Rotation(limit_angle=np.pi / 4, p=1.0)(sweep.copy(), labels.copy())
        small["rot%d" % k], small["rot%d_next" % k] = lidar, np.array(np.random.random())
        np.random.seed(seed)
        lidar, lab = ns.Random_Scaling(scaling_range=(0.95 + 0.01 * k, 1.05), p=1.0)(sweep.copy(), labels.copy())
        small["scl%d" % k], small["scl%d_labels" % k], small["scl%d_next" % k] = lidar, lab, np.array(np.random.random())
    small["labels"] = labels
    np.savez_compressed(os.path.join(HERE, "augment_small.npz"), **small)
    return len(AUG_PARAMS), 4


class _Calib:
    def __init__(self, V2C, R0, P2):
        self.V2C, self.R0, self.P2 = V2C, R0, P2


def projection_cases():
    """(name, per-class detection dict in post_processing's format, calibration seed, image shape)."""
    cases = []
    for s in range(3):
        heads = O.synth_heads(500 + s, B=1, tie_free=True)
        det = O.decode(*heads, K=50).numpy().astype(np.float32)
        cases.append(("decode%d" % s, O.post_processing(det, peak_thresh=0.2)[0], s, (375, 1242)))
    f = np.float32
    # score, x_px, y_px, z, h, w_px, l_px, yaw   (x_px across y, y_px along x: evaluation_utils.py:184-185)
    rows = np.array([
        [0.9, 304.0, 243.2, 1.0, 1.5, 20.0, 48.0, 0.3],      # 20 m ahead, centred
        [0.9, 304.0, 3.0, 1.0, 1.5, 20.0, 48.0, 1.2],        # straddles the image plane (depth changes sign)
        [0.9, 10.0, 12.0, 1.0, 1.5, 20.0, 48.0, -2.0],       # far left, mostly outside the image
        [0.9, 600.0, 30.0, 1.0, 1.5, 20.0, 48.0, 0.0],       # far right, outside the image
        [0.9, 304.0, 600.0, 2.0, 3.0, 30.0, 120.0, 3.1],     # 49 m ahead, a few pixels
        [0.9, 304.0, 243.2, np.nan, 1.5, 20.0, 48.0, 0.3],   # NaN height -> Python max/min fall back to the border
        [0.9, 304.0, 100.0, 1.0, 0.0, 0.0, 0.0, 0.0],        # degenerate box: zero area, rejected
        [0.9, 304.0, 120.0, 30.0, 1.5, 20.0, 48.0, 0.5],     # above the image
    ], dtype=f)
    cases.append(("crafted", {0: rows[:2].copy(), 1: rows.copy(), 2: rows[::-1].copy()}, 7, (375, 1242)))
    cases.append(("empty", {0: np.zeros((0, 8), f), 1: np.zeros((0, 8), f), 2: np.zeros((0, 8), f)}, 1, (375, 1242)))
    cases.append(("small_image", {0: rows[:1].copy(), 1: rows[:5].copy(), 2: rows[2:6].copy()}, 2, (128, 416)))
    return cases


def make_projection(ns):
    small = {}
    cases = projection_cases()
    for name, dets, cseed, shape in cases:
        V2C, R0, P2 = O.synth_calibration(cseed)
        for j in range(3):
            small["%s_det%d" % (name, j)] = np.asarray(dets[j], np.float32).reshape(-1, 8)
        small[name + "_V2C"], small[name + "_R0"], small[name + "_P2"] = V2C, R0, P2
        small[name + "_shape"] = np.array(shape)
        real = np.asarray(ns.convert_det_to_real_values(dets), np.float64).reshape(-1, 8)
        small[name + "_real"] = real
        small[name + "_cam"] = (ns.lidar_to_camera_box(real[:, 1:], V2C, R0, P2) if len(real)
                                else np.zeros((0, 7)))
        with np.errstate(all="ignore"):
            boxes, conf = ns.convert_sfa3d_to_2d_boxes(dets, _Calib(V2C, R0, P2), shape)
        small[name + "_boxes"] = np.asarray(boxes, np.int64).reshape(-1, 4)
        small[name + "_conf"] = np.asarray(conf, np.float64)
    small["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "projection_small.npz"), **small)
    return len(cases)


def main():
    if not ref_loader.available():
        raise SystemExit("reference not present; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)
    ns = ref_loader.load()
    only = sys.argv[1:]
    if not only or "bev" in only:
        print("bev: %d small cases, %d hashed" % make_bev(ns))
    if not only or "decode" in only:
        print("decode: %d small cases, %d hashed" % make_decode(ns))
    if not only or "bvfeature" in only:
        print("bvfeature: %d small cases, %d hashed" % make_bvfeature(ns))
    if not only or "augment" in only:
        print("augment: %d point_transform cases, %d seeded draws" % make_augment(ns))
    if not only or "projection" in only:
        print("projection: %d cases" % make_projection(ns))


if __name__ == "__main__":
    main()
