"""Throughput of makeBVFeature on the GPU (sfa_bvfeature_rasterize): 64 sweeps of 250k points
(Argoverse range, 800 x 800 map), inputs resident in HBM, 3 input sets rotating (768 MB > L2).
Prints frames/s, the fraction of the HBM roofline (16 N + 12 H W algorithmic bytes per frame) and
the per-kernel times from the library's event brackets."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfa_b200  # noqa: E402
from sfa_b200 import _lib, fast  # noqa: E402
import sfa_oracle as O  # noqa: E402  (boundary constants only)

B, N, SETS, STEPS = 64, int(os.environ.get("N", 250000)), 3, 30
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
sets = []
for s in range(SETS):
    p = torch.empty((B, N, 4), device=dev)
    p[..., 0] = torch.rand((B, N), generator=g, device=dev) * 80 - 40
    p[..., 1] = torch.rand((B, N), generator=g, device=dev) * 80 - 40
    p[..., 2] = torch.rand((B, N), generator=g, device=dev) * 4 - 3
    p[..., 3] = torch.randint(0, 256, (B, N), generator=g, device=dev).float()
    sets.append(p)
rast = fast.BvFeatureRasterizer(O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY, max_batch=B)
out = torch.empty((B, 3, rast.H, rast.W), device=dev)
for i in range(3):
    rast(sets[i % SETS], out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(STEPS):
    rast(sets[i % SETS], out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
with _lib.profile() as prof:
    rast(sets[0], out=out)
    torch.cuda.synchronize()
bytes_per_frame = 16 * N + 12 * rast.H * rast.W
peak = 6545.6e9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass
fps = B / (ms * 1e-3)
print(json.dumps({"workload": "makeBVFeature 64 x %d pts -> 3x%dx%d" % (N, rast.H, rast.W), "ms_per_batch": round(ms, 4),
                  "frames_per_s": round(fps, 1), "us_per_frame": round(ms * 1e3 / B, 3),
                  "roofline_frac": round(fps * bytes_per_frame / peak, 4),
                  "kernels_ms": {k: [v[0], round(v[1], 4)] for k, v in prof.stats.items()}}))
