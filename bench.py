#!/usr/bin/env python
"""bench.py — BEV rasterisation + peak decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): a batch of 64 synthetic KITTI-shaped sweeps (120,000 points each,
uniform in the KITTI boundary) -> 64 x [3,608,608] BEV maps, and 64 frames of heads (hm 3x152x152,
cen_offset 2, direction 2, z 1, dim 3) -> _nms/_topk/decode K=50 -> [64,50,10] -> dense
post_processing.  One "step" = one such batch per GPU.  N > 1: every rank runs the same per-GPU
batch on its own frames (weak scaling, frames shard with no collective on the data path).

Reported on ONE JSON line (rank 0):
  value     frames/s, whole job, inputs resident in HBM, K steps replayed as CUDA graphs, CUDA-event
            timed, max over ranks.  The inputs rotate over `sets` distinct batches (> L2) so that no
            step finds its inputs in L2.
  e2e       same metric through the host-buffer C ABI (sfa_pipeline_bev_host / _decode_host):
            pinned host sweeps+heads in, BEV maps + detections back in pinned host memory.
  roofline  dominant kernel: algorithmic bytes per launch / CUDA-event duration of that launch,
            measured live in a separate un-captured pass with events around every library launch.
  cpu_baseline  the oracle port of the reference (numpy lexsort/unique + torch max_pool/topk) on the
            host cores of this box, bounded sample (N=1 only).
"""
import argparse
import importlib
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

N_POINTS, BEV_H, BEV_W = 120_000, 608, 608
HEAD_C, HEAD_H, HEAD_W, TOPK = 3, 152, 152, 50
# 16 N + 12 H W + 4 C h w + 32*8 K + 40 K
BYTES_BEV = 16 * N_POINTS + 12 * BEV_H * BEV_W
BYTES_DECODE = 4 * HEAD_C * HEAD_H * HEAD_W + 32 * 8 * TOPK + 40 * TOPK
BYTES_FRAME = BYTES_BEV + BYTES_DECODE
# algorithmic bytes per FRAME each kernel is responsible for (DESIGN.md "kernels")
KERNEL_BYTES = {
    "bev_raster": 16 * N_POINTS,            # global-atomic path: points in
    "bev_finalize": 12 * BEV_H * BEV_W,     #                     planes out
    "bev_bin": 16 * N_POINTS,               # tiled path: points in (records stay in L2)
    "bev_band": 12 * BEV_H * BEV_W,         #             planes out
    "peak_candidates": 4 * HEAD_C * HEAD_H * HEAD_W,    # heat map in (candidate list stays in L2)
    "peak_select": 32 * 8 * TOPK + 40 * TOPK,           # gathered regression sectors + detections out
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


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, one frame = filter -> makeBEVMap -> .float() -> decode
# (B=1, as every reference script calls it) -> post_processing   
# ------------------------------------------------------------------------------------------------
def _cpu_worker(wid, fpw_max, fpw_now, n_steps, barrier, out_q):
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    import sfa_oracle as O
    torch.set_num_threads(1)
    sweeps = [O.synth_sweep(10_000 + wid * fpw_max + j, N_POINTS, O.KITTI, "uniform") for j in range(fpw_max)]
    heads = [O.synth_heads(20_000 + wid * fpw_max + j, B=1) for j in range(fpw_max)]
    checksum = 0.0
    for _ in range(n_steps):
        barrier.wait()
        for j in range(fpw_now.value):
            s, h = sweeps[j], heads[j]
            filt = O.get_filtered_lidar(s, O.KITTI.boundary)
            bev = torch.from_numpy(O.makeBEVMap(filt, O.KITTI.boundary, O.KITTI)).float()
            det = O.decode(h[0], h[1], h[2], h[3], h[4], K=TOPK).numpy().astype(np.float32)
            pp = O.post_processing(det, 3, 4, 0.2)
            checksum += float(bev[1].sum()) + float(det[0, 0, 0]) + len(pp)
        barrier.wait()
    out_q.put((wid, checksum))


def run_cpu_arm(steps, warmup, batch, budget_s=120.0):
    """All host cores, one single-threaded process per core; each step = every worker runs its
    share of the batch between two barriers.  Returns (frames_per_step, [seconds per timed step],
    n_workers).  A calibration pass (one frame per worker) sizes the per-step share so that
    steps+warmup passes fit `budget_s`: the share is min(ceil(batch / cores), what fits), >= 1."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    cores = host_cores()
    n_workers = min(cores, batch)
    fpw_max = max(1, math.ceil(batch / n_workers))
    fpw_now = ctx.Value("i", 1)
    barrier = ctx.Barrier(n_workers + 1)
    q = ctx.Queue()
    total_steps = 1 + warmup + steps
    procs = [ctx.Process(target=_cpu_worker, args=(w, fpw_max, fpw_now, total_steps, barrier, q), daemon=True)
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
        """Samples inside [t0, t1] (host monotonic time of the timed region); when the region is
        shorter than the sampling period, the samples of the preceding warm-up (same load) are used."""
        inside = [l for (t, l) in self.lines if t0 is None or (t0 <= t <= t1 + 0.1)]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [l for (_, l) in self.lines], "warm-up + timed region"
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
def make_inputs(batch, sets, rank, torch):
    """`sets` distinct batches of sweeps + heads on the host (numpy / torch CPU)."""
    import sfa_oracle as O  # only its synthetic-input generators are used here
    out = []
    for s in range(sets):
        base = 1_000_000 * rank + 1000 * s
        rng = np.random.default_rng(base)
        b = O.KITTI.boundary
        pts = np.empty((batch, N_POINTS, 4), dtype=np.float32)
        pts[:, :, 0] = rng.uniform(b["minX"], b["maxX"], (batch, N_POINTS))
        pts[:, :, 1] = rng.uniform(b["minY"], b["maxY"], (batch, N_POINTS))
        pts[:, :, 2] = rng.uniform(b["minZ"], b["maxZ"], (batch, N_POINTS))
        pts[:, :, 3] = rng.uniform(0, 1, (batch, N_POINTS))
        heads = O.synth_heads(base + 7, B=batch, C=HEAD_C, h=HEAD_H, w=HEAD_W)
        out.append((pts, heads))
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

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised (workers are forked)
        fps, times, cores = run_cpu_arm(steps=2, warmup=1, batch=args.batch)
        v = fps * len(times) / sum(times)
        cpu_baseline = {"value": round(v, 2), "unit": "frames/s", "cores": cores, "kind": "port",
                        "value_per_core": round(v / cores, 2), "cpu": cpu_model_name(),
                        "sample": "%d frames (2 timed passes of %d after 1 warm-up) of the same workload, one process per "
                                  "core, oracle port of the reference (numpy lexsort+unique BEV, torch max_pool2d+topk decode B=1, "
                                  "post_processing)" % (fps * len(times), fps)}

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pkg("build").build()
    lib = pkg("_lib")
    fast, ev = pkg("fast"), pkg("utils.evaluation_utils")
    geom = pkg("geometry").from_config(pkg("config.kitti_config"))
    B, sets = args.batch, args.sets

    host_sets = make_inputs(B, sets, rank, torch)
    dev_sets = []
    for pts, heads in host_sets:
        dev_sets.append((torch.from_numpy(pts).to(dev).reshape(-1, 4), tuple(t.to(dev) for t in heads)))
    # A step runs on an "engine": the batch is split over `lanes` independent rasterisers, each with
    # its own workspace and CUDA stream, and the decode runs on a stream of its own — sweeps and heads
    # are independent inputs, so inside one graph replay the latency-bound phases of one lane overlap
    # the others'.  `pipelines` engines (each with its own workspaces and output buffers) take the
    # steps in turn on their own launch streams, so step i+1 ramps up while step i drains — what a
    # double-buffered inference loop does.  Every step does the full work; nothing is shared or reused.
    lanes = max(1, min(args.lanes, B))
    lane_frames = [(B * i // lanes, B * (i + 1) // lanes) for i in range(lanes)]
    # every synthetic sweep holds exactly N_POINTS points: the uniform-batch form of the API (offsets=None);
    # --ragged-api passes an explicit offsets array instead (same data, one more dependent load per tile)
    lane_offsets = [torch.arange(b1 - b0 + 1, dtype=torch.int64, device=dev) * N_POINTS if args.ragged_api else None
                    for b0, b1 in lane_frames]

    class Engine:
        def __init__(self):
            self.rasts = [fast.BevRasterizer(geom, max_batch=b1 - b0, max_points=N_POINTS, device=dev)
                          for b0, b1 in lane_frames]
            self.side = [torch.cuda.Stream(device=dev) for _ in range(lanes)]   # lanes 1.. and the decode
            self.launch = torch.cuda.Stream(device=dev)
            self.bev_out = torch.empty((B, 3, BEV_H, BEV_W), dtype=torch.float32, device=dev)
            self.det_out = torch.empty((B, TOPK, 10), dtype=torch.float32, device=dev)
            self.pp_out = (torch.empty((B, TOPK, 8), dtype=torch.float32, device=dev),
                           torch.empty((B, TOPK), dtype=torch.int32, device=dev),
                           torch.empty((B, TOPK), dtype=torch.uint8, device=dev))
            self.dec_ws = fast.DecodeWorkspace(dev, B, HEAD_C, HEAD_H, HEAD_W, TOPK)
            self.graphs = []

        def step_serial(self, s):
            """Same launches as step(), all on the current stream (per-kernel event timing, ncu)."""
            pts, heads = dev_sets[s % sets]
            for i, (b0, b1) in enumerate(lane_frames):
                self.rasts[i](pts[b0 * N_POINTS:b1 * N_POINTS], lane_offsets[i], N_POINTS, out=self.bev_out[b0:b1])
            fast.decode_device(*heads, K=TOPK, out=self.det_out, workspace=self.dec_ws)
            fast.post_process_dense(self.det_out, out=self.pp_out)

        def step(self, s):
            if args.eager:
                return self.step_serial(s)
            pts, heads = dev_sets[s % sets]
            main = torch.cuda.current_stream(dev)
            for st in self.side:
                st.wait_stream(main)
            for i, (b0, b1) in enumerate(lane_frames):
                with torch.cuda.stream(main if i == 0 else self.side[i - 1]):
                    self.rasts[i](pts[b0 * N_POINTS:b1 * N_POINTS], lane_offsets[i], N_POINTS, out=self.bev_out[b0:b1])
            with torch.cuda.stream(self.side[-1]):
                fast.decode_device(*heads, K=TOPK, out=self.det_out, workspace=self.dec_ws)
                fast.post_process_dense(self.det_out, out=self.pp_out)
            for st in self.side:
                main.wait_stream(st)

    class _Eager:
        def __init__(self, eng, s):
            self.eng, self.s = eng, s

        def replay(self):
            self.eng.step(self.s)

    n_pipe = 1 if args.eager else max(1, args.pipelines)
    engines = [Engine() for _ in range(n_pipe)]
    launches_per_step = 0
    cap_stream = torch.cuda.Stream(device=dev)
    for eng in engines:
        for s in range(sets):   # eager warm-up (also loads every kernel)
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
        # warm-up: at least W (>= 3) steps and at least 0.5 s, so clocks settle and get sampled under load
        n_warm, t_w = 0, time.monotonic()
        fork()
        while n_warm < max(args.warmup, 3) or (not args.eager and time.monotonic() - t_w < 0.5):
            replay(n_warm)
            n_warm += 1
            if n_warm % 16 == 0:
                torch.cuda.synchronize()
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

    # ---- per-kernel device time (un-captured pass, events around every library launch) -------------
    torch.cuda.synchronize()
    with lib.profile() as prof:
        # a ~10 ms spin kernel first: every launch and event of the pass queues up behind it, so the
        # event brackets measure back-to-back device time, not the host's launch latency
        torch.cuda._sleep(20_000_000)
        for i in range(max(4, sets)):
            step_serial(i)
        torch.cuda.synchronize()
    hbm_gbs, peak_src = peaks()
    n_prof_steps = max(4, sets)
    kern = {}
    for name, (n_launch, tot_ms) in prof.stats.items():
        per_step_ms = tot_ms / n_prof_steps
        kern[name] = {"launches_per_step": n_launch / n_prof_steps, "ms_per_step": round(per_step_ms, 5),
                      "ms_per_launch": round(tot_ms / n_launch, 5)}
        if name in KERNEL_BYTES:
            gbs = KERNEL_BYTES[name] * B / (per_step_ms * 1e-3) / 1e9
            kern[name]["algorithmic_GBps"] = round(gbs, 1)
            kern[name]["frac_of_hbm_peak"] = round(gbs / hbm_gbs, 4)
    dominant = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    traffic = None
    roofline = None
    if dominant and dominant in KERNEL_BYTES:
        frames_per_launch = B / kern[dominant]["launches_per_step"]
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):   # ncu --set full capture (profiles/README.md), scaled to this run's frames per launch
            with open(tpath) as f:
                ent = json.load(f).get(dominant)
            if ent:
                traffic = int(ent["dram_bytes_per_launch"] * frames_per_launch / ent["frames_per_launch"])
        achieved = KERNEL_BYTES[dominant] * frames_per_launch / (kern[dominant]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": hbm_gbs, "unit": "GB/s",
                    "frac": round(achieved / hbm_gbs, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(KERNEL_BYTES[dominant] * frames_per_launch)}
    path_gbs = value / world * BYTES_FRAME / 1e9
    roofline_path = {"bytes_per_frame": BYTES_FRAME, "achieved": round(path_gbs, 1), "peak": hbm_gbs, "unit": "GB/s",
                     "frac": round(path_gbs / hbm_gbs, 4), "per": "GPU, whole path (all kernels of a step)"}

    # ---- e2e: host buffers through the host-pipeline C ABI ------------------------------------------
    e2e = None
    if not args.no_e2e:
        pts_h, heads_h = host_sets[0]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pts_pin = pin(pts_h.reshape(-1, 4))
        heads_pin = tuple(t.contiguous().pin_memory() for t in heads_h)
        offs_h = (np.arange(B + 1, dtype=np.int64) * N_POINTS)
        # `pipelines` independent workers (host threads), each with a BEV pipeline and a decode pipeline of
        # its own (own streams, staging buffers, pinned outputs), take the steps in turn: within a step the
        # decode's uploads overlap the BEV maps' downloads (PCIe is full duplex; ctypes releases the GIL),
        # and one step's uploads overlap the previous step's downloads.
        from concurrent.futures import ThreadPoolExecutor
        n_work = max(1, args.pipelines)
        pts_np = pts_pin.numpy()
        heads_np = [t.numpy() for t in heads_pin]

        class HostWorker:
            def __init__(self):
                self.bev = fast.HostPipeline(geom, max_frames=B, max_points=N_POINTS, C=0, h=1, w=1, K=1, device=local_rank)
                self.dec = fast.HostPipeline(geom, max_frames=B, max_points=0, C=HEAD_C, h=HEAD_H, w=HEAD_W, K=TOPK,
                                             device=local_rank)
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
        engines[0].step_serial(0)
        torch.cuda.synchronize()
        e2e_ok = bool(torch.equal(engines[0].bev_out.cpu(), bev_pin) and torch.equal(engines[0].det_out.cpu(), det_pin))
        if not e2e_ok:
            raise SystemExit("e2e outputs differ from the device-resident path")
        h2d = pts_pin.numel() * 4 + sum(t.numel() * 4 for t in heads_pin) + offs_h.nbytes
        d2h = bev_pin.numel() * 4 + det_pin.numel() * 4
        e2e = {"value": round(B * e2e_steps * world / dt, 1), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": round(dt / e2e_steps * 1e3, 4), "steps": e2e_steps,
               "outputs_equal_device_path": e2e_ok,
               "api": "sfa_pipeline_bev_host + sfa_pipeline_decode_host (pinned host sweeps/heads in, host BEV maps + "
                      "detections out), %d host workers taking steps in turn" % n_work}
        pool.shutdown()
        for wk in workers:
            wk.close()

    if rank == 0:
        line = {
            "metric": "BEV+decode frames/s", "value": round(value, 1), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": round(ms_total / args.steps, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "batch of %d synthetic KITTI sweeps (%d pts, uniform in the KITTI boundary) -> %d x "
                                   "[3,608,608] BEV + _nms/_topk/decode K=%d on [%d,%d,%d] heads + dense post_processing, per GPU per step"
                                   % (B, N_POINTS, B, TOPK, HEAD_C, HEAD_H, HEAD_W),
                       "frames_per_step_per_gpu": B, "l2_policy": "inputs rotate over %d distinct batches (%.0f MB) > L2" %
                       (sets, sets * B * (16 * N_POINTS + 44 * HEAD_H * HEAD_W) / 1e6),
                       "cuda_graph": not args.eager, "streams": "%d engine(s) alternating steps; per engine %d BEV lanes + 1 decode stream" % (n_pipe, lanes), "sharding": "frames, no collective on the data path"},
            "gpu_launches": int(launches_per_step * args.steps),
            "e2e": e2e, "roofline": roofline, "roofline_path": roofline_path, "kernels": kern,
            "cpu_baseline": cpu_baseline, "clocks": clocks.summary(t_begin, t_end),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own algorithm (oracle port; the reference is Python and does not travel to
    the GPU box) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    fps, times, cores = run_cpu_arm(steps=args.steps, warmup=args.warmup, batch=args.batch)
    total = sum(times)
    v = fps * len(times) / total
    sample = ("each step = %d frames (%d per process x %d processes, one per host core) of the same workload; oracle "
              "port of get_filtered_lidar+makeBEVMap+.float()+decode(B=1,K=%d)+post_processing" %
              (fps, fps // cores, cores, TOPK))
    if fps < args.batch:
        sample += "; bounded below the %d-frame batch so that %d steps fit the time budget" % (args.batch, args.steps + args.warmup)
    line = {
        "impl": "reference", "metric": "BEV+decode frames/s", "value": round(v, 2), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / len(times) * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic KITTI sweeps (%d pts) -> [3,608,608] BEV + decode K=%d on [%d,%d,%d] heads + "
                               "post_processing, reference numpy/torch algorithm on host cores" %
                               (N_POINTS, TOPK, HEAD_C, HEAD_H, HEAD_W), "frames_per_step": fps},
        "gpu_launches": 0,
        "cpu_baseline": {"value": round(v, 2), "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model_name()},
        "e2e": {"value": round(v, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--pipelines", type=int, default=2,
                    help="engines (own workspaces, outputs, streams) that take the steps in turn, so consecutive steps overlap")
    ap.add_argument("--lanes", type=int, default=4, help="independent BEV streams the batch is split over")
    ap.add_argument("--ragged-api", action="store_true", help="pass an offsets array instead of the uniform-batch form")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true",
                    help="profiling aid (ncu): plain launches instead of CUDA-graph replays, exactly W warm-up steps; "
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
        args.steps = 2000 if args.steps is None else args.steps
        args.warmup = 5 if args.warmup is None else args.warmup
        run_b200(args)


if __name__ == "__main__":
    main()
