#!/usr/bin/env python
"""bench.py — BEV rasterisation + peak decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path, headline config
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's own functions on host cores
    python bench.py --config {headline,density1r,argoverse,stream8192,loop32} ...

Configs (BASELINE.json `configs`):
  headline    configs[1]: batch of 64 synthetic KITTI sweeps (120,000 points uniform in the KITTI boundary) ->
              64 x [3,608,608] BEV maps, and 64 frames of heads (hm 3x152x152, cen_offset 2, direction 2, z 1, dim 3)
              -> _nms/_topk/decode K=50 -> [64,50,10] -> dense post_processing.  One step = one such batch per GPU.
  density1r   the same with sweeps whose range density falls as 1/r (a spinning lidar's), not uniform.
  argoverse   configs[3]: 250,000-point sweeps in the Argoverse range (config/argoverse_config.py:16-23, x,y in
              [-50,50], z in [-3,5], D = 100/608), same grid and heads; 8,728,016 algorithmic bytes per frame.
  stream8192  configs[4]: a stream of 8192 KITTI sweeps, seeds 0..8191, frame i -> rank i mod G
              (sharding.shard_range(..., "cyclic")), device-resident, processed in 64-frame batches; one step = one
              batch of the rank's shard, the default step count covers the whole stream once.
  loop32      configs[2]: the full inference loop at batch 32 — GPU BEV -> the reference's random-init fpn_resnet_18
              (stock PyTorch, from the git-ignored baseline/_ref copy) -> _sigmoid -> GPU decode -> post-processing;
              end-to-end frames/s and the hot path's share of the step by CUDA events.
N > 1: every rank runs the same per-GPU batch on its own frames (weak scaling; frames shard with no collective on the
data path).

ONE JSON line (rank 0):
  value        frames/s, whole job, inputs resident in HBM, K steps replayed as CUDA graphs, CUDA-event timed, max over
               ranks.  The inputs rotate over `sets` distinct batches (> L2) so that no step finds its inputs in L2.
  e2e          same metric through the host-buffer C ABI (sfa_pipeline_bev_host / _decode_host): pinned host
               sweeps + heads in, BEV maps + detections back in pinned host memory.
  roofline     dominant kernel: algorithmic bytes per launch / CUDA-event duration of that launch, measured live in a
               separate un-captured pass with events around every library launch; `traffic` = DRAM bytes per launch of
               that kernel in this configuration from the committed ncu capture (profiles/roofline_traffic.json).
  cpu_baseline the reference's own functions (baseline/_ref copy; the oracle port when absent) on the host cores of this
               box, bounded sample (N=1 only), plus a one-process one-thread figure.
"""
import argparse
import contextlib
import importlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

BEV_H, BEV_W = 608, 608
HEAD_C, HEAD_H, HEAD_W, TOPK = 3, 152, 152, 50
KITTI_BOUNDARY = {"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}      # config/kitti_config.py:23-30
ARGO_BOUNDARY = {"minX": -50, "maxX": 50, "minY": -50, "maxY": 50, "minZ": -3, "maxZ": 5}           # config/argoverse_config.py:16-23

WORKLOADS = {
    "headline": {"n_points": 120_000, "geom": "kitti", "dist": "uniform", "batch": 64,
                 "what": "batch of %(B)d synthetic KITTI sweeps (%(N)d pts, uniform in the KITTI boundary)"},
    "density1r": {"n_points": 120_000, "geom": "kitti", "dist": "lidar1r", "batch": 64,
                  "what": "batch of %(B)d synthetic KITTI sweeps (%(N)d pts, range density ~ 1/r in the front 90 degrees)"},
    "argoverse": {"n_points": 250_000, "geom": "argoverse", "dist": "uniform", "batch": 64,
                  "what": "batch of %(B)d synthetic Argoverse-range sweeps (%(N)d pts, uniform in config/argoverse_config.py's boundary, D = 100/608)"},
    "stream8192": {"n_points": 120_000, "geom": "kitti", "dist": "uniform", "batch": 64, "stream": 8192,
                   "what": "stream of 8192 synthetic KITTI sweeps (%(N)d pts, seeds 0..8191), frame i -> rank i mod G, in batches of %(B)d"},
    "loop32": {"n_points": 120_000, "geom": "kitti", "dist": "uniform", "batch": 32,
               "what": "full inference loop, batch %(B)d: %(N)d-pt KITTI sweeps -> GPU BEV -> reference fpn_resnet_18 (random init, stock PyTorch fp32) -> _sigmoid -> GPU decode"},
}


def bytes_per_frame(n_points):
    """16 N + 12 H W + 4 C h w + 32*8 K + 40 K."""
    bev = 16 * n_points + 12 * BEV_H * BEV_W
    dec = 4 * HEAD_C * HEAD_H * HEAD_W + 32 * 8 * TOPK + 40 * TOPK
    return bev, dec


def kernel_bytes(n_points):
    """Algorithmic bytes per FRAME each kernel is responsible for (DESIGN.md "kernels")."""
    return {
        "bev_fused": 16 * n_points + 12 * BEV_H * BEV_W,   # single persistent kernel: points in, planes out (records stay in L2)
        "bev_raster": 16 * n_points, "bev_finalize": 12 * BEV_H * BEV_W,
        "bev_bin": 16 * n_points, "bev_band": 12 * BEV_H * BEV_W,
        "peak_candidates": 4 * HEAD_C * HEAD_H * HEAD_W,
        "peak_select": 32 * 8 * TOPK + 40 * TOPK,
        "post_process": 40 * TOPK + 32 * TOPK + 5 * TOPK,
    }


def pkg(sub=None):
    return importlib.import_module(PKG if sub is None else PKG + "." + sub)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def synth_sweeps(seed, batch, n_points, geom, dist):
    """[batch, n_points, 4] float32, drawn in float64 and cast."""
    b = ARGO_BOUNDARY if geom == "argoverse" else KITTI_BOUNDARY
    rng = np.random.default_rng(seed)
    pts = np.empty((batch, n_points, 4), dtype=np.float32)
    if dist == "lidar1r":      # range density ~ 1/r (rings of a spinning lidar), azimuth uniform in the front 90 degrees
        r = np.exp(rng.uniform(np.log(2.0), np.log(70.0), (batch, n_points)))
        az = rng.uniform(-np.pi / 4, np.pi / 4, (batch, n_points))
        pts[:, :, 0], pts[:, :, 1] = r * np.cos(az), r * np.sin(az)
    else:
        pts[:, :, 0] = rng.uniform(b["minX"], b["maxX"], (batch, n_points))
        pts[:, :, 1] = rng.uniform(b["minY"], b["maxY"], (batch, n_points))
    pts[:, :, 2] = rng.uniform(b["minZ"], b["maxZ"], (batch, n_points))
    pts[:, :, 3] = rng.uniform(0, 1, (batch, n_points))
    return pts


# ------------------------------------------------------------------------------------------------
# CPU arm: one frame = filter -> makeBEVMap -> .float() -> decode (B=1, as every reference script calls
# it) -> post_processing.  kind "reference": the reference's own, unmodified functions
# from the baseline/_ref copy; kind "port": the oracle's restatement (only where that copy is absent).
# ------------------------------------------------------------------------------------------------
def reference_kind():
    import ref_loader
    return "reference" if ref_loader.available() else "port"


def _cpu_frame_fn(kind, geom_name):
    import torch
    import sfa_oracle as O
    ogeom = O.ARGOVERSE if geom_name == "argoverse" else O.KITTI
    if kind == "reference":
        import ref_loader
        ns = ref_loader.load()
        sink = io.StringIO()

        def frame(s, h):
            with contextlib.redirect_stdout(sink):     # the reference's post_processing prints every detection
                ctx = (ref_loader.patched_geometry(ns, ogeom.boundary, ogeom.BEV_HEIGHT, ogeom.BEV_WIDTH, ogeom.DISCRETIZATION)
                       if geom_name == "argoverse" else contextlib.nullcontext())
                with ctx:
                    filt = ns.get_filtered_lidar(s, ogeom.boundary)                        # kitti_dataset.py:88
                    bev = torch.from_numpy(ns.makeBEVMap(filt, ogeom.boundary)).float()    # :90, test.py:124
                    det = ns.decode(h[0], h[1], h[2], h[3], h[4], K=TOPK).numpy().astype(np.float32)   # test.py:167-172
                    pp = ns.evl.post_processing(det, 3, 4, 0.2)                            # test.py:173
            sink.seek(0); sink.truncate()
            return float(bev[1].sum()) + float(det[0, 0, 0]) + len(pp)
        return frame

    def frame(s, h):
        filt = O.get_filtered_lidar(s, ogeom.boundary)
        bev = torch.from_numpy(O.makeBEVMap(filt, ogeom.boundary, ogeom)).float()
        det = O.decode(h[0], h[1], h[2], h[3], h[4], K=TOPK).numpy().astype(np.float32)
        pp = O.post_processing(det, 3, 4, 0.2)
        return float(bev[1].sum()) + float(det[0, 0, 0]) + len(pp)
    return frame


def _cpu_worker(wid, kind, wl, fpw_max, fpw_now, n_steps, barrier, out_q):
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    import sfa_oracle as O
    torch.set_num_threads(1)
    frame = _cpu_frame_fn(kind, wl["geom"])
    sweeps = [synth_sweeps(10_000 + wid * fpw_max + j, 1, wl["n_points"], wl["geom"], wl["dist"])[0] for j in range(fpw_max)]
    heads = [O.synth_heads(20_000 + wid * fpw_max + j, B=1) for j in range(fpw_max)]
    checksum = 0.0
    for _ in range(n_steps):
        barrier.wait()
        for j in range(fpw_now.value):
            checksum += frame(sweeps[j], heads[j])
        barrier.wait()
    out_q.put((wid, checksum))


def run_cpu_arm(wl, steps, warmup, batch, kind, workers=None, budget_s=120.0):
    """`workers` single-threaded processes (default: one per host core); each step = every worker runs its share of
    the batch between two barriers.  Returns (frames_per_step, [seconds per timed step], n_workers).  A calibration
    pass (one frame per worker) sizes the per-step share so that steps+warmup passes fit `budget_s`."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    n_workers = min(workers or host_cores(), batch)
    fpw_max = max(1, math.ceil(batch / n_workers))
    fpw_now = ctx.Value("i", 1)
    barrier = ctx.Barrier(n_workers + 1)
    q = ctx.Queue()
    total_steps = 1 + warmup + steps
    procs = [ctx.Process(target=_cpu_worker, args=(w, kind, wl, fpw_max, fpw_now, total_steps, barrier, q), daemon=True)
             for w in range(n_workers)]
    for p in procs:
        p.start()
    times = []
    for i in range(total_steps):
        barrier.wait(timeout=1800)
        t0 = time.perf_counter()
        barrier.wait(timeout=1800)
        dt = time.perf_counter() - t0
        if i == 0:  # calibration: one frame per worker
            fit = int(budget_s / max(dt, 1e-3) / max(1, steps + warmup))
            fpw_now.value = max(1, min(fpw_max, fit))
        elif i > warmup:
            times.append(dt)
    for _ in procs:
        q.get(timeout=60)
    for p in procs:
        p.join(timeout=60)
    return fpw_now.value * n_workers, times, n_workers


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_block(wl, batch, steps, warmup, budget_s):
    """All cores + the one-process one-thread figure BASELINE.md §3 asks for."""
    kind = reference_kind()
    fps, times, cores = run_cpu_arm(wl, steps=steps, warmup=warmup, batch=batch, kind=kind, budget_s=budget_s)
    v = fps * len(times) / sum(times)
    f1, t1, _ = run_cpu_arm(wl, steps=1, warmup=0, batch=8, kind=kind, workers=1, budget_s=8.0)
    v1 = f1 * len(t1) / sum(t1)
    what = ("the reference's own get_filtered_lidar + makeBEVMap + .float() + decode(B=1, K=%d) + post_processing (unmodified, "
            "imported from the baseline/_ref copy)" % TOPK if kind == "reference" else
            "oracle port of get_filtered_lidar + makeBEVMap (numpy lexsort+unique) + .float() + decode(B=1, K=%d, torch max_pool2d+topk) "
            "+ post_processing (reference copy absent)" % TOPK)
    return {"value": round(v, 2), "unit": "frames/s", "cores": cores, "kind": kind, "cpu": cpu_model_name(),
            "single_process_1_thread": {"value": round(v1, 2), "unit": "frames/s", "frames": f1 * len(t1)},
            "sample": "%d frames (%d timed passes of %d after %d warm-up) of the same workload, one single-threaded process per core; %s"
                      % (fps * len(times), len(times), fps, warmup, what)}, fps, times


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return False

    def summary(self, t0=None, t1=None):
        """Samples inside [t0, t1] (host monotonic time of the timed region); when the region is shorter than the
        sampling period, the samples of the preceding settle window (same load) are used and the window says so."""
        inside = [l for (t, l) in self.lines if t0 is None or (t0 <= t <= t1 + 0.1)]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [l for (_, l) in self.lines], "settle + warm-up + timed region"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in inside:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
def product_geometry(geom_name):
    if geom_name == "argoverse":
        class Cnf:
            BEV_HEIGHT, BEV_WIDTH, DISCRETIZATION = BEV_H, BEV_W, (ARGO_BOUNDARY["maxX"] - ARGO_BOUNDARY["minX"]) / BEV_H
        return pkg("geometry").BevGeometry(ARGO_BOUNDARY, Cnf)
    return pkg("geometry").from_config(pkg("config.kitti_config"))


def synth_heads_host(seed, batch, torch):
    import sfa_oracle as O  # only its synthetic-input generator
    return O.synth_heads(seed, B=batch, C=HEAD_C, h=HEAD_H, w=HEAD_W)


def stream_frames_device(frame_ids, n_points, dev, torch):
    """Sweeps of the 8192-frame stream, generated on the device from their frame number (seed = frame id): x, y, z,
    intensity uniform in the KITTI boundary (float32 arithmetic on the device; parity of this config is checked through
    size-independent properties and by copying sample frames back, tests/test_bev_gpu.py)."""
    b = KITTI_BOUNDARY
    lo = torch.tensor([b["minX"], b["minY"], b["minZ"], 0.0], device=dev)
    span = torch.tensor([b["maxX"] - b["minX"], b["maxY"] - b["minY"], b["maxZ"] - b["minZ"], 1.0], device=dev)
    out = torch.empty((len(frame_ids), n_points, 4), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for j, fid in enumerate(frame_ids):
        g.manual_seed(int(fid))
        out[j] = torch.rand((n_points, 4), generator=g, device=dev) * span + lo
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU); see the module docstring" % args.gpus)
        args.gpus = world
    wl = WORKLOADS[args.config]
    B, N = (args.batch or wl["batch"]), wl["n_points"]
    bytes_bev, bytes_dec = bytes_per_frame(N)
    bytes_frame = bytes_bev + bytes_dec
    KB = kernel_bytes(N)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised (workers are forked)
        cpu_baseline, _, _ = cpu_baseline_block(wl, B, steps=2, warmup=1, budget_s=20.0)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pkg("build").build()
    lib = pkg("_lib")
    fast = pkg("fast")
    geom = product_geometry(wl["geom"])
    sets = args.sets

    # ---- inputs resident in HBM ---------------------------------------------------------------------
    stream_info = None
    if "stream" in wl:
        shard = list(pkg("sharding").shard_range(wl["stream"], rank, world, "cyclic"))
        n_batches = len(shard) // B          # whole batches of the rank's shard (8192 / G is a multiple of 64 for G <= 8)
        sets = n_batches
        dev_pts = [stream_frames_device(shard[i * B:(i + 1) * B], N, dev, torch).reshape(-1, 4) for i in range(n_batches)]
        heads_host = synth_heads_host(1_000_000 * rank + 7, B, torch)
        dev_heads = [tuple(t.to(dev) for t in heads_host)] * n_batches     # decode inputs: one resident set (1 MB / frame)
        host_sets = None
        stream_info = {"frames_total": wl["stream"], "frames_this_rank": len(shard), "batches_this_rank": n_batches,
                       "sharding": "sharding.shard_range(8192, rank, world, 'cyclic'): frame i -> rank i mod G",
                       "first_frames_this_rank": shard[:4]}
        if args.steps is None:
            args.steps = n_batches
    else:
        host_sets = []
        for s in range(sets):
            base = 1_000_000 * rank + 1000 * s
            host_sets.append((synth_sweeps(base, B, N, wl["geom"], wl["dist"]), synth_heads_host(base + 7, B, torch)))
        dev_pts = [torch.from_numpy(p).to(dev).reshape(-1, 4) for p, _ in host_sets]
        dev_heads = [tuple(t.to(dev) for t in h) for _, h in host_sets]
    if args.steps is None:
        args.steps = 2000

    # A step runs on an "engine": the batch is split over `lanes` independent rasterisers, each with its own workspace
    # and CUDA stream, and the decode runs on a stream of its own.  `pipelines` engines (own workspaces and outputs) take
    # the steps in turn on their own launch streams.  Every step does the full work; nothing is shared or reused.
    lanes = max(1, min(args.lanes, B))
    lane_frames = [(B * i // lanes, B * (i + 1) // lanes) for i in range(lanes)]
    lane_offsets = [torch.arange(b1 - b0 + 1, dtype=torch.int64, device=dev) * N if args.ragged_api else None
                    for b0, b1 in lane_frames]

    class Engine:
        def __init__(self):
            self.rasts = [fast.BevRasterizer(geom, max_batch=b1 - b0, max_points=N, device=dev) for b0, b1 in lane_frames]
            # lanes 1.. and (last) the decode, which gets the lower stream priority: its latency-bound kernels fill the
            # slots the BEV kernels leave instead of competing for them
            self.side = [torch.cuda.Stream(device=dev, priority=(-1 if (i < lanes - 1 or not args.decode_low_priority) else 0))
                         for i in range(lanes)]
            self.launch = torch.cuda.Stream(device=dev, priority=-1)
            self.bev_out = torch.empty((B, 3, BEV_H, BEV_W), dtype=torch.float32, device=dev)
            self.det_out = torch.empty((B, TOPK, 10), dtype=torch.float32, device=dev)
            self.pp_out = (torch.empty((B, TOPK, 8), dtype=torch.float32, device=dev),
                           torch.empty((B, TOPK), dtype=torch.int32, device=dev),
                           torch.empty((B, TOPK), dtype=torch.uint8, device=dev))
            self.dec_ws = fast.DecodeWorkspace(dev, B, HEAD_C, HEAD_H, HEAD_W, TOPK)
            self.graphs = []

        def step_serial(self, s):
            """Same launches as step(), all on the current stream (per-kernel event timing, ncu)."""
            pts, heads = dev_pts[s % sets], dev_heads[s % sets]
            if args.only != "decode":
                for i, (b0, b1) in enumerate(lane_frames):
                    self.rasts[i](pts[b0 * N:b1 * N], lane_offsets[i], N, out=self.bev_out[b0:b1])
            if args.only != "bev":
                self.decode(heads)

        def decode(self, heads):
            if args.separate_post:
                fast.decode_device(*heads, K=TOPK, out=self.det_out, workspace=self.dec_ws)
                fast.post_process_dense(self.det_out, out=self.pp_out)
            else:   # the dense post-processing rows come out of the decode's own epilogue (sfa_decode_post)
                fast.decode_device(*heads, K=TOPK, out=self.det_out, workspace=self.dec_ws, post=self.pp_out)

        def step(self, s):
            if args.eager or (lanes == 1 and not args.decode_stream):
                return self.step_serial(s)
            pts, heads = dev_pts[s % sets], dev_heads[s % sets]
            main = torch.cuda.current_stream(dev)
            for st in self.side:
                st.wait_stream(main)
            if args.only != "decode":
                for i, (b0, b1) in enumerate(lane_frames):
                    with torch.cuda.stream(main if i == 0 else self.side[i - 1]):
                        self.rasts[i](pts[b0 * N:b1 * N], lane_offsets[i], N, out=self.bev_out[b0:b1])
            if args.only != "bev":
                with torch.cuda.stream(self.side[-1]):
                    self.decode(heads)
            for st in self.side:
                main.wait_stream(st)

    class _Eager:
        def __init__(self, eng, s):
            self.eng, self.s = eng, s

        def replay(self):
            self.eng.step(self.s)

    if args.config == "loop32":
        return run_loop32(args, torch, dev, fast, geom, dev_pts, B, N, sets, cpu_baseline, bytes_frame, lib)

    n_pipe = 1 if args.eager else max(1, args.pipelines)
    # The library itself deals the 8-frame chunks of a call to two internal streams (sfa_bev_set_internal_lanes).  With several
    # engines / lanes overlapping at this level that is redundant (measured: -4 %), so the multi-engine schedule turns it off;
    # the `single_call` figure below is the library's default as one caller on one stream gets it.
    lib.load().sfa_bev_set_internal_lanes(1 if (n_pipe > 1 or lanes > 1) else 0)
    engines = [Engine() for _ in range(n_pipe)]
    launches_per_step = 0
    cap_stream = torch.cuda.Stream(device=dev)
    for eng in engines:
        for s in range(min(sets, 4)):   # eager warm-up (also loads every kernel)
            eng.step(s)
        torch.cuda.synchronize()
        for s in range(sets):   # then one graph per input set
            n0 = lib.kernel_launches()
            if args.eager:
                g = _Eager(eng, s)
                g.replay()
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream):
                    eng.step(s)
            launches_per_step = lib.kernel_launches() - n0
            eng.graphs.append(g)
    step_serial = engines[0].step_serial

    def replay(i):
        """Step i: engine i % n_pipe on its own launch stream (steps of one engine stay in order)."""
        eng = engines[i % n_pipe]
        if n_pipe == 1:
            eng.graphs[i % sets].replay()
        else:
            with torch.cuda.stream(eng.launch):
                eng.graphs[i % sets].replay()

    def fork():
        main = torch.cuda.current_stream(dev)
        if n_pipe > 1:
            for eng in engines:
                eng.launch.wait_stream(main)

    def join():
        main = torch.cuda.current_stream(dev)
        if n_pipe > 1:
            for eng in engines:
                main.wait_stream(eng.launch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        # settle: >= 0.5 s of the same steps first (not the warm-up the caller asked for: clocks ramp up and the
        # 100-ms clock sampler gets samples under this load even when the timed region lasts only a few ms) ...
        n_settle, t_w = 0, time.monotonic()
        fork()
        while not args.eager and time.monotonic() - t_w < args.settle_s:
            replay(n_settle)
            n_settle += 1
            if n_settle % 16 == 0:
                torch.cuda.synchronize()
        join()
        barrier()
        # ... then exactly W warm-up steps, then exactly K timed steps
        fork()
        for i in range(args.warmup):
            replay(i)
        join()
        barrier()
        t_begin = time.monotonic()
        e0.record()
        fork()
        for i in range(args.steps):
            replay(i)
        join()
        e1.record()
        barrier()
        t_end = time.monotonic()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    frames = B * args.steps * world
    value = frames / (ms_total * 1e-3)

    # ---- stage ablation in the SAME schedule: the timed graphs again with one stage left out ---------------------------
    stage = None
    if not args.eager and args.only == "both" and not args.no_stage_ablation and rank == 0 and world == 1:
        def timed_only(which, n):
            saved = args.only
            args.only = which
            try:
                graphs = []
                for eng in engines:
                    gs = []
                    for s_ in range(sets):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=cap_stream):
                            eng.step(s_)
                        gs.append(g)
                    graphs.append(gs)
            finally:
                args.only = saved

            def rep(i):
                eng = engines[i % n_pipe]
                if n_pipe == 1:
                    graphs[0][i % sets].replay()
                else:
                    with torch.cuda.stream(eng.launch):
                        graphs[i % n_pipe][i % sets].replay()
            fork()
            for i in range(3 * n_pipe):
                rep(i)
            join()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            fork()
            for i in range(n):
                rep(i)
            join()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n
        n_ab = max(60, min(args.steps, 400))
        ms_bev, ms_dec = timed_only("bev", n_ab), timed_only("decode", n_ab)
        hbm_gbs_, _ = peaks()
        bev_gbs = bytes_bev * B / (ms_bev * 1e-3) / 1e9
        stage = {"bev_only_ms_per_step": round(ms_bev, 5), "decode_only_ms_per_step": round(ms_dec, 5),
                 "both_ms_per_step": round(ms_total / args.steps, 5),
                 "decode_adds_ms_per_step": round(ms_total / args.steps - ms_bev, 5), "steps": n_ab,
                 "bev_stage_roofline": {"achieved": round(bev_gbs, 1), "peak": hbm_gbs_, "unit": "GB/s", "frac": round(bev_gbs / hbm_gbs_, 4),
                                        "bytes_per_frame": bytes_bev,
                                        "what": "bev_bin + bev_band of a step as the timed schedule runs them (same engines, streams and "
                                                "graphs, decode left out): algorithmic bytes / in-schedule time"}}

    # ---- the same step as ONE caller issues it: one engine, one BEV call per step on one stream (+ the decode stream),
    #      the library's internal lanes at their default ------------------------------------------------------------
    single_call = None
    if not args.eager and args.only == "both" and (n_pipe > 1 or lanes > 1) and not args.no_single_call:
        lib.load().sfa_bev_set_internal_lanes(0)
        eng1 = Engine()
        for s_ in range(min(sets, 4)):
            eng1.step(s_)
        torch.cuda.synchronize()
        for s_ in range(sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap_stream):
                eng1.step(s_)
            eng1.graphs.append(g)
        n1 = max(50, min(args.steps, 400))
        for i in range(20):
            eng1.graphs[i % sets].replay()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n1):
            eng1.graphs[i % sets].replay()
        s1.record()
        barrier()
        ms1 = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ms1], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms1 = float(t.item())
        single_call = {"value": round(B * n1 * world / (ms1 * 1e-3), 1), "unit": "frames/s", "steps": n1,
                       "ms_per_step": round(ms1 / n1, 5),
                       "what": "one engine, steps back to back on one stream: one sfa_bev_rasterize call per step (the library overlaps "
                               "its 8-frame chunks on 2 internal streams) + the decode on a second stream"}
        lib.load().sfa_bev_set_internal_lanes(1)
        del eng1

    # ---- per-kernel device time (un-captured, serialised pass, events around every library launch) ---------------
    torch.cuda.synchronize()
    with lib.profile() as prof:
        # a ~10 ms spin kernel first: every launch and event of the pass queues up behind it, so the
        # event brackets measure back-to-back device time, not the host's launch latency
        torch.cuda._sleep(20_000_000)
        for i in range(max(4, min(sets, 8))):
            step_serial(i)
        torch.cuda.synchronize()
    hbm_gbs, peak_src = peaks()
    n_prof_steps = max(4, min(sets, 8))
    kern = {}
    for name, (n_launch, tot_ms) in prof.stats.items():
        per_step_ms = tot_ms / n_prof_steps
        kern[name] = {"launches_per_step": n_launch / n_prof_steps, "ms_per_step": round(per_step_ms, 5),
                      "ms_per_launch": round(tot_ms / n_launch, 5)}
        if name in KB:
            gbs = KB[name] * B / (per_step_ms * 1e-3) / 1e9
            kern[name]["algorithmic_GBps"] = round(gbs, 1)
            kern[name]["frac_of_hbm_peak"] = round(gbs / hbm_gbs, 4)
    dominant = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    roofline = None
    if dominant and dominant in KB:
        frames_per_launch = B / kern[dominant]["launches_per_step"]
        # DRAM bytes of one whole step in THIS schedule (ncu range replay over several overlapped steps: a per-kernel
        # capture serialises the launches and flushes the L2-resident bucket hand-off between them), profiles/README.md
        traffic, traffic_src, traffic_alg = None, None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                ent = json.load(f).get(args.config)
            if ent and int(ent.get("engines", 0)) == n_pipe and int(ent.get("frames_per_bev_launch", 0)) == int(frames_per_launch):
                traffic, traffic_src = int(ent["dram_bytes_per_step"]), ent.get("source")
                traffic_alg = int(ent["algorithmic_bytes_per_step"])
        achieved = KB[dominant] * frames_per_launch / (kern[dominant]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": hbm_gbs, "unit": "GB/s",
                    "frac": round(achieved / hbm_gbs, 4), "traffic": traffic, "traffic_per": "step (%d frames, all kernels)" % B,
                    "traffic_algorithmic": traffic_alg, "traffic_source": traffic_src,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": int(KB[dominant] * frames_per_launch),
                    "frames_per_launch": int(frames_per_launch),
                    "timing": "CUDA events around each launch of a serialised, un-captured pass of the same step",
                    "note": "isolated launches of %d frames: alone, a launch this small pays its partial last wave and ramp in full "
                            "(the same kernel at 32 frames per launch: 0.52); in the timed schedule a second engine fills those gaps "
                            "— the in-schedule, whole-path figure is roofline_path" % int(frames_per_launch)}
        if stage:   # the BEV stage (both kernels) in the timed schedule, decode left out: see stage_ablation
            roofline["in_schedule_stage"] = {"kernels": "bev_bin + bev_band", "achieved": stage["bev_stage_roofline"]["achieved"],
                                             "frac": stage["bev_stage_roofline"]["frac"]}
    path_gbs = value / world * bytes_frame / 1e9
    roofline_path = {"bytes_per_frame": bytes_frame, "achieved": round(path_gbs, 1), "peak": hbm_gbs, "unit": "GB/s",
                     "frac": round(path_gbs / hbm_gbs, 4), "per": "GPU, whole path (all kernels of a step), from `value`"}

    # ---- e2e: host buffers through the host-pipeline C ABI ------------------------------------------
    e2e = None
    if not args.no_e2e and host_sets is not None and args.only == "both":
        e2e = run_e2e(args, torch, dist, dev, local_rank, world, fast, geom, host_sets[0], B, N, engines[0], barrier)

    if rank == 0:
        line = {
            "metric": "BEV+decode frames/s", "value": round(value, 1), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (wl["what"] % {"B": B, "N": N}) + " -> %d x [3,608,608] BEV + _nms/_topk/decode K=%d on [%d,%d,%d] heads "
                                   "+ dense post_processing, per GPU per step" % (B, TOPK, HEAD_C, HEAD_H, HEAD_W),
                       "name": args.config if args.only == "both" else "%s (ABLATION: %s stage only)" % (args.config, args.only),
                       "frames_per_step_per_gpu": B,
                       "l2_policy": "inputs rotate over %d distinct batches (%.0f MB) > L2" %
                       (sets, sets * B * (16 * N + 44 * HEAD_H * HEAD_W) / 1e6),
                       "cuda_graph": not args.eager, "settle_steps_before_warmup": n_settle,
                       "streams": "%d engine(s) alternating steps; per engine %d BEV lane(s)%s" %
                                  (n_pipe, lanes, " + 1 decode stream" if (lanes > 1 or args.decode_stream) else ", decode on the same stream"),
                       "sharding": "frames, no collective on the data path"},
            "gpu_launches": int(launches_per_step * args.steps),
            "e2e": e2e, "roofline": roofline, "roofline_path": roofline_path, "stage_ablation": stage, "single_call": single_call,
            "kernels_serialised": kern,
            "cpu_baseline": cpu_baseline, "clocks": clocks.summary(t_begin, t_end),
        }
        if stream_info:
            line["config"]["stream"] = stream_info
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, torch, dist, dev, local_rank, world, fast, geom, host_set, B, N, engine, barrier):
    """The same metric through the reference-facing host-buffer C ABI: pinned host sweeps + heads in, host BEV maps +
    detections out, copies inside the timed region."""
    pkg("_lib").load().sfa_bev_set_internal_lanes(0)   # the host pipeline as any caller gets it (library default)
    pts_h, heads_h = host_set
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pts_pin = pin(pts_h.reshape(-1, 4))
    heads_pin = tuple(t.contiguous().pin_memory() for t in heads_h)
    offs_h = (np.arange(B + 1, dtype=np.int64) * N)
    # `workers` independent host threads, each with a BEV pipeline and a decode pipeline of its own (own streams,
    # staging buffers, pinned outputs), take the steps in turn: within a step the decode's uploads overlap the BEV maps'
    # downloads (PCIe is full duplex; ctypes releases the GIL), and one step's uploads overlap the previous step's downloads.
    from concurrent.futures import ThreadPoolExecutor
    n_work = max(1, args.e2e_workers)
    pts_np = pts_pin.numpy()
    heads_np = [t.numpy() for t in heads_pin]

    class HostWorker:
        def __init__(self):
            self.bev = fast.HostPipeline(geom, max_frames=B, max_points=N, C=0, h=1, w=1, K=1, device=local_rank)
            self.dec = fast.HostPipeline(geom, max_frames=B, max_points=0, C=HEAD_C, h=HEAD_H, w=HEAD_W, K=TOPK, device=local_rank)
            self.bev_out = torch.empty((B, 3, BEV_H, BEV_W), dtype=torch.float32).pin_memory().numpy()
            self.det_out = torch.empty((B, TOPK, 10), dtype=torch.float32).pin_memory().numpy()
            self.side = ThreadPoolExecutor(max_workers=1)

        def step(self):
            fut = self.side.submit(self.dec.decode, *heads_np, out=self.det_out)
            self.bev.bev(pts_np, offs_h, out=self.bev_out)
            fut.result()

        def run(self, n):
            for _ in range(n):
                self.step()

        def close(self):
            self.side.shutdown()
            self.bev.close()
            self.dec.close()

    workers = [HostWorker() for _ in range(n_work)]
    pool = ThreadPoolExecutor(max_workers=n_work)
    e2e_steps = max(n_work, min(args.steps, 200))
    share = [e2e_steps // n_work + (1 if i < e2e_steps % n_work else 0) for i in range(n_work)]
    list(pool.map(lambda w: w.run(2), workers))   # warm-up
    barrier()
    t0 = time.perf_counter()
    list(pool.map(lambda wn: wn[0].run(wn[1]), zip(workers, share)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    bev_pin, det_pin = torch.from_numpy(workers[0].bev_out), torch.from_numpy(workers[0].det_out)
    # the host path must deliver what the device path computes for the same inputs (set 0)
    engine.step_serial(0)
    torch.cuda.synchronize()
    e2e_ok = bool(torch.equal(engine.bev_out.cpu(), bev_pin) and torch.equal(engine.det_out.cpu(), det_pin))
    if not e2e_ok:
        raise SystemExit("e2e outputs differ from the device-resident path")
    h2d = pts_pin.numel() * 4 + sum(t.numel() * 4 for t in heads_pin) + offs_h.nbytes
    d2h = bev_pin.numel() * 4 + det_pin.numel() * 4
    value = B * e2e_steps * world / dt
    e2e = {"value": round(value, 1), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": round(dt / e2e_steps * 1e3, 4), "steps": e2e_steps,
           "outputs_equal_device_path": e2e_ok,
           "pcie_GBps_per_gpu": {"h2d": round(h2d * e2e_steps / dt / 1e9, 2), "d2h": round(d2h * e2e_steps / dt / 1e9, 2)},
           "api": "sfa_pipeline_bev_host + sfa_pipeline_decode_host (pinned host sweeps/heads in, host BEV maps + "
                  "detections out), %d host workers taking steps in turn" % n_work}
    cpath = os.path.join(ROOT, "profiles", "pcie_ceiling.json")
    if os.path.exists(cpath):   # memcpy-only ceiling of the same buffers (tools/pcie_ceiling.py), per GPU count
        with open(cpath) as f:
            ent = json.load(f).get(str(world))
        if ent:
            e2e["memcpy_only_ceiling_frames_per_s"] = ent["frames_per_s"]
            e2e["frac_of_memcpy_ceiling"] = round(value / ent["frames_per_s"], 3)
    pool.shutdown()
    for wk in workers:
        wk.close()
    return e2e


def run_loop32(args, torch, dev, fast, geom, dev_pts, B, N, sets, cpu_baseline, bytes_frame, lib):
    """BASELINE configs[2]: sweeps -> GPU BEV -> the reference's own fpn_resnet_18 (random init; the backbone is out of
    scope and stays stock PyTorch) -> _sigmoid -> GPU decode -> dense post-processing, batch 32, all on one stream.
    The reference network is the checker-side copy (baseline/_ref, see oracle/ref_loader.py): it is not product code."""
    import ref_loader
    if not ref_loader.available():
        emit({"metric": "BEV+decode frames/s", "config": {"name": "loop32"}, "unavailable":
              "the reference's fpn_resnet_18 is not present (baseline/_ref/sfa is made by __graft_entry__.build() where /root/reference exists)"})
        return
    tu = pkg("utils.torch_utils")
    net = ref_loader.create_model("fpn_resnet_18", seed=0).to(dev)
    rast = fast.BevRasterizer(geom, max_batch=B, max_points=N, device=dev)
    bev = torch.empty((B, 3, BEV_H, BEV_W), dtype=torch.float32, device=dev)
    det = torch.empty((B, TOPK, 10), dtype=torch.float32, device=dev)
    pp = (torch.empty((B, TOPK, 8), dtype=torch.float32, device=dev), torch.empty((B, TOPK), dtype=torch.int32, device=dev),
          torch.empty((B, TOPK), dtype=torch.uint8, device=dev))
    ws = fast.DecodeWorkspace(dev, B, HEAD_C, HEAD_H, HEAD_W, TOPK)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(s, timed=False):
        with torch.no_grad():
            if timed: ev[0].record()
            rast(dev_pts[s % sets], None, N, out=bev)
            if timed: ev[1].record()
            out = net(bev)                                  # test.py:149
            hm, off = tu._sigmoid(out["hm_cen"]), tu._sigmoid(out["cen_offset"])   # test.py:150-151
            if timed: ev[2].record()
            fast.decode_device(hm, off, out["direction"], out["z_coor"], out["dim"], K=TOPK, out=det, workspace=ws)
            fast.post_process_dense(det, out=pp)
            if timed: ev[3].record()

    for s in range(max(3, args.warmup)):
        step(s)
    torch.cuda.synchronize()
    n0 = lib.kernel_launches()
    steps = min(args.steps or 50, 200)
    t_bev = t_net = t_dec = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clocks:
        t_begin = time.monotonic()
        e0.record()
        for s in range(steps):
            step(s, timed=True)
            ev[3].synchronize()
            t_bev += ev[0].elapsed_time(ev[1]); t_net += ev[1].elapsed_time(ev[2]); t_dec += ev[2].elapsed_time(ev[3])
        e1.record()
        torch.cuda.synchronize()
        t_end = time.monotonic()
    ms = e0.elapsed_time(e1) / steps
    hot = (t_bev + t_dec) / steps
    emit({"metric": "BEV+decode frames/s", "value": round(B / (ms * 1e-3), 1), "unit": "frames/s", "n_gpus": 1, "steps": steps,
          "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": WORKLOADS["loop32"]["what"] % {"B": B, "N": N}, "name": "loop32", "frames_per_step_per_gpu": B,
                     "l2_policy": "inputs rotate over %d distinct batches" % sets, "cuda_graph": False,
                     "backbone": "reference models/fpn_resnet.py fpn_resnet_18, 12,728,353 parameters, torch.manual_seed(0), eval, fp32, stock PyTorch/cuDNN"},
          "loop": {"bev_ms": round(t_bev / steps, 4), "backbone_plus_sigmoid_ms": round(t_net / steps, 4),
                   "decode_post_ms": round(t_dec / steps, 4), "hot_path_share_of_step": round(hot / (t_bev + t_net + t_dec) * steps, 4),
                   "hot_path_frames_per_s": round(B / (hot * 1e-3), 1),
                   "timing": "CUDA events around the three stages on one stream; value = whole loop including the host's per-step event sync"},
          "gpu_launches": int(lib.kernel_launches() - n0), "e2e": None, "roofline": None, "cpu_baseline": cpu_baseline,
          "clocks": clocks.summary(t_begin, t_end)})


def run_reference(args):
    """The reference's own implementation of the path on all host cores (baseline/_ref copy of its Python packages; the
    oracle port only where that copy is absent).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config if args.config != "loop32" else "headline"]
    B, N = (args.batch or wl["batch"]), wl["n_points"]
    block, fps, times = cpu_baseline_block(wl, B, steps=args.steps, warmup=args.warmup, budget_s=120.0)
    total = sum(times)
    v = block["value"]
    if fps < B:
        block["sample"] += "; bounded below the %d-frame batch so that %d steps fit the time budget" % (B, args.steps + args.warmup)
    line = {
        "impl": "reference", "metric": "BEV+decode frames/s", "value": v, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / len(times) * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (wl["what"] % {"B": B, "N": N}) + " -> %d x [3,608,608] BEV + _nms/_topk/decode K=%d on [%d,%d,%d] heads "
                               "+ dense post_processing, per GPU per step" % (B, TOPK, HEAD_C, HEAD_H, HEAD_W),
                   "name": args.config, "frames_per_step_per_gpu": B, "frames_per_step_sampled": fps,
                   "runs_on": "host cores (reference numpy/torch code path)"},
        "gpu_launches": 0, "cpu_baseline": block,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_RESULT_FD = None


def emit(line):
    """The ONE line of this run on the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--pipelines", type=int, default=3,
                    help="engines (own workspaces, outputs, streams) that take the steps in turn, so consecutive steps overlap")
    ap.add_argument("--lanes", type=int, default=1, help="independent BEV streams the batch is split over")
    ap.add_argument("--decode-stream", type=int, default=1, help="1: the decode runs on a stream of its own next to the BEV lane(s)")
    ap.add_argument("--decode-low-priority", type=int, default=0, help="1: the decode stream runs at lower priority than the BEV lanes")
    ap.add_argument("--e2e-workers", type=int, default=2, help="host threads of the e2e leg, each with its own pipelines")
    ap.add_argument("--settle-s", type=float, default=0.5, help="seconds of untimed steps before the warm-up (clock ramp, sampler)")
    ap.add_argument("--only", default="both", choices=["both", "bev", "decode"],
                    help="ablation: time one stage alone in the same schedule (the printed value is then NOT the metric)")
    ap.add_argument("--separate-post", action="store_true", help="post_processing as its own launch instead of the decode's epilogue")
    ap.add_argument("--ragged-api", action="store_true", help="pass an offsets array instead of the uniform-batch form")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stage-ablation", action="store_true", help="skip the BEV-only / decode-only runs of the timed schedule")
    ap.add_argument("--no-single-call", action="store_true", help="skip the one-engine / one-stream secondary measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true",
                    help="profiling aid (ncu): plain launches instead of CUDA-graph replays, no settle window; "
                         "the printed value is not a bench number")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries print meanwhile (NCCL's version banner under
    # torchrun, warnings of worker processes) is routed to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        args.steps = 4 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.warmup = 5 if args.warmup is None else max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
