#!/usr/bin/env python
"""Reads an ncu report's source page (SASS) for one kernel and prints the instructions that collect the
most warp-stall samples, with the dominant stall reason.  Usage: tools/ncu_hot.py report.ncu-rep regex [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[col["# Samples"]] != "# Samples":
        body.append(r)
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print("kernel:", rows[0][1][:100]); print("total samples", tot, "warp-instructions", tot_inst)
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print("%5.1f%%  inst=%8s  %-18s %s" % (100.0 * n / max(tot, 1), r[col["Instructions Executed"]], why, r[col["Source"]][:90]))

# ---- contiguous SASS segments with (nearly) the same execution count = loops / phases --------------
print("segments (>= 1% of executed warp-instructions):")
prev, seg_start, acc, segs = None, 0, 0, []
for i, r in enumerate(body):
    c = int(r[col["Instructions Executed"]] or 0)
    if prev is None or abs(c - prev) > max(2000, 0.1 * prev):
        if prev is not None:
            segs.append((seg_start, i - 1, acc))
        seg_start, acc = i, 0
    acc += c
    prev = c
segs.append((seg_start, len(body) - 1, acc))
for s0, e0, a in segs:
    if a / max(tot_inst, 1) > 0.01:
        smp = sum(int(r[col["# Samples"]] or 0) for r in body[s0:e0 + 1])
        print("  sass %4d-%4d  inst %5.1f%%  samples %5.1f%%  x%-8s %s" % (
            s0, e0, 100.0 * a / tot_inst, 100.0 * smp / max(tot, 1), body[s0][col["Instructions Executed"]], body[s0][col["Source"]][:50]))
