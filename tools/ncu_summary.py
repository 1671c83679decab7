#!/usr/bin/env python
"""Reads an `ncu --set full` report (tools/gpu_profile.sh) and writes the per-kernel summary committed as
profiles/rNN_ncu_full_summary.json plus profiles/roofline_traffic.json (DRAM bytes per launch, read by bench.py).
Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_summary.json profiles/roofline_traffic.json"""
import csv
import json
import subprocess
import sys

rep, out_summary, out_traffic = sys.argv[1:4]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
           "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
seen, summary, traffic = set(), [], {}
FRAMES = {"peak_candidates": 64, "peak_select": 64, "bev_bin": 32, "bev_band": 32}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
    if name in seen:
        continue
    seen.add(name)
    entry = {"kernel": name}
    for m in METRICS:
        if m in col:
            entry[m] = ("%s %s" % (r[col[m]], units[col[m]])).strip()
    summary.append(entry)
    short = {"peak_candidates_kernel": "peak_candidates", "peak_select_kernel": "peak_select",
             "bev_bin_staged_kernel": "bev_bin", "bev_band_kernel": "bev_band"}.get(name)
    if short:
        def mb(metric):
            v, u = float(r[col[metric]].replace(",", "")), units[col[metric]].lower()
            return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}[u]
        rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
        traffic[short] = {"dram_bytes_per_launch": int(round((rd + wr) * 1e6)), "dram_read_MB": rd, "dram_write_MB": wr,
                          "frames_per_launch": FRAMES[short],
                          "note": "ncu --set full, cold L2 at kernel start, --lanes 1 (32 frames per BEV launch); bev_band reads = "
                                  "its records, which a real run finds in L2; part of its writes is still dirty in L2 at kernel end"}
json.dump(summary, open(out_summary, "w"), indent=1)
json.dump(traffic, open(out_traffic, "w"), indent=1)
print("kernels:", [e["kernel"] for e in summary])
