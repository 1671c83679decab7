"""CPU: the parts of bench.py's contract that need no GPU — the algorithmic-byte formula of SURVEY.md §8d, the sharding of
the 8192-frame stream (BASELINE configs[4]), and the reference arm's JSON line (the reference's own functions when the
baseline/_ref copy or /root/reference is present, the oracle port otherwise)."""
import json
import os
import subprocess
import sys

import numpy as np

import bench
import ref_loader
from conftest import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_baseline_md():
    bev, dec = bench.bytes_per_frame(120_000)
    assert bev == 6_355_968 and dec == 292_048 and bev + dec == 6_648_016          # BASELINE.md §2, KITTI config
    assert sum(bench.bytes_per_frame(250_000)) == 8_728_016                          # Argoverse-range config
    kb = bench.kernel_bytes(120_000)
    assert kb["bev_bin"] + kb["bev_band"] == kb["bev_fused"] == bev
    assert kb["peak_candidates"] + kb["peak_select"] == dec


def test_stream_sharding_is_a_partition():
    sharding = pkg("sharding")
    for world in (1, 2, 4, 8):
        shards = [list(sharding.shard_range(8192, r, world, "cyclic")) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(8192))
        assert all(len(s) % 64 == 0 for s in shards)          # whole 64-frame batches on every rank
        assert all(s[0] == r and s[1] == r + world for r, s in enumerate(shards))


def test_synthetic_sweeps_are_seeded_and_inside_the_boundary():
    a = bench.synth_sweeps(5, 2, 1000, "kitti", "uniform")
    b = bench.synth_sweeps(5, 2, 1000, "kitti", "uniform")
    assert a.dtype == np.float32 and a.shape == (2, 1000, 4) and np.array_equal(a, b)
    assert a[..., 0].min() >= 0 and a[..., 0].max() <= 50 and np.abs(a[..., 1]).max() <= 25
    g = bench.synth_sweeps(5, 1, 1000, "argoverse", "uniform")
    assert g[..., 0].min() >= -50 and g[..., 0].max() <= 50 and g[..., 2].max() <= 5
    r = bench.synth_sweeps(5, 1, 20000, "kitti", "lidar1r")
    rad = np.hypot(r[0, :, 0], r[0, :, 1])
    assert np.median(rad) < 15 and rad.max() > 50            # density falls with range


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BEV+decode frames/s" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if ref_loader.available() else "port") and cb["cores"] >= 1
    assert cb["single_process_1_thread"]["value"] > 0 and "sample" in cb
