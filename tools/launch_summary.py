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
OURS = ("bev_", "peak_", "post_process", "filter_", "nms_kernel")   # libsfa_b200 kernels; the rest is torch set-up / the spin blocker
tot = sum(sum(v) for k, v in agg.items() if k.startswith(OURS))
print("shares are of the time spent in libsfa_b200 kernels (cold-cache, serialised launches)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = "%5.1f%%" % (100 * sum(v) / tot) if k.startswith(OURS) else "  (not ours)"
    print("%-50s n=%4d  mean=%9.2f us  total=%10.1f us  share=%s" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, share))
