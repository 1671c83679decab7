#!/usr/bin/env python
"""Memcpy-only ceiling of bench.py's `e2e` leg: the SAME pinned buffers and byte counts per step (sweeps + heads host ->
device, BEV maps + detections device -> host), no kernels.  Run on 1 GPU or under torchrun on N GPUs of one box:

    python tools/pcie_ceiling.py [--steps 60]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py

Every rank copies on two streams (H2D and D2H concurrently, like the pipeline does: PCIe is full duplex) in 16-frame
chunks, waits with a blocking event, and the job's frames/s is taken over the slowest rank.  Rank 0 prints one JSON line
and merges it into profiles/pcie_ceiling.json under the GPU count (bench.py reports e2e as a fraction of it)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, N, H, W, C, h, w, K = 64, 120_000, 608, 608, 3, 152, 152, 50


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--no-write", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    h2d_elems = B * N * 4 + B * (C + 8) * h * w          # sweeps + the five heads (float32)
    d2h_elems = B * 3 * H * W + B * K * 10               # BEV maps + detections
    src = torch.empty(h2d_elems, dtype=torch.float32).pin_memory()
    dst = torch.empty(d2h_elems, dtype=torch.float32).pin_memory()
    src.uniform_()
    d_in = torch.empty(h2d_elems, dtype=torch.float32, device=dev)
    d_out = torch.zeros(d2h_elems, dtype=torch.float32, device=dev)
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    done = [torch.cuda.Event(blocking=True), torch.cuda.Event(blocking=True)]
    n_chunks = B // args.chunk

    def step():
        for c in range(n_chunks):
            a0, a1 = h2d_elems * c // n_chunks, h2d_elems * (c + 1) // n_chunks
            b0, b1 = d2h_elems * c // n_chunks, d2h_elems * (c + 1) // n_chunks
            with torch.cuda.stream(up):
                d_in[a0:a1].copy_(src[a0:a1], non_blocking=True)
            with torch.cuda.stream(down):
                dst[b0:b1].copy_(d_out[b0:b1], non_blocking=True)
        done[0].record(up)
        done[1].record(down)
        done[0].synchronize()
        done[1].synchronize()

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        fps = B * args.steps * world / dt
        line = {"n_gpus": world, "frames_per_s": round(fps, 1), "ms_per_step": round(dt / args.steps * 1e3, 3),
                "h2d_bytes_per_step": h2d_elems * 4, "d2h_bytes_per_step": d2h_elems * 4,
                "h2d_GBps_per_gpu": round(h2d_elems * 4 * args.steps / dt / 1e9, 2),
                "d2h_GBps_per_gpu": round(d2h_elems * 4 * args.steps / dt / 1e9, 2),
                "aggregate_GBps": round((h2d_elems + d2h_elems) * 4 * args.steps * world / dt / 1e9, 1),
                "host_cores": len(os.sched_getaffinity(0)),
                "what": "memcpy only (pinned host <-> device, two streams per rank, %d-frame chunks, blocking event waits); no kernels" % args.chunk}
        print(json.dumps(line))
        if not args.no_write:
            path = os.path.join(ROOT, "profiles", "pcie_ceiling.json")
            table = {}
            if os.path.exists(path):
                with open(path) as f:
                    table = json.load(f)
            table[str(world)] = line
            out_dir = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out_dir, exist_ok=True)
            for p in (path, os.path.join(out_dir, "pcie_ceiling.json")):
                with open(p, "w") as f:
                    json.dump(table, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
