#!/usr/bin/env python
"""Warp-stall samples and executed instructions per CUDA SOURCE LINE from an ncu report captured with
--import-source on (kernels built with -lineinfo).  Usage: tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, cur, col, hdr = {}, None, None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; col = {}
        for i, n in enumerate(hdr):
            col.setdefault(n, i)
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (cur, int(r[0]))
        a = agg.setdefault(key, {"src": r[1], "samples": 0, "inst": 0, "stalls": {}})
        try:
            a["samples"] += int(r[col["# Samples"]] or 0); a["inst"] += int(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        for n, i in col.items():
            if n.startswith("stall_") and "Not Issued" not in n:
                try: a["stalls"][n] = a["stalls"].get(n, 0) + int(r[i] or 0)
                except ValueError: pass
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print("total samples %d, warp-instructions %d" % (tot, toti))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    why = max(a["stalls"], key=a["stalls"].get) if a["stalls"] else ""
    print("%5.1f%% smp %5.1f%% inst  %-14s %s:%d  %s" % (100.0 * a["samples"] / max(tot, 1), 100.0 * a["inst"] / max(toti, 1), why, f, ln, a["src"].strip()[:95]))
