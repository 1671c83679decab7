#!/usr/bin/env python
"""Summarises an ncu launch list (gpu__time_duration.sum per launch): per kernel count / mean / total and share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki].split("(")[0].split("::")[-1][:48], []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-50s n=%4d  mean=%9.2f us  total=%10.1f us  share=%5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
