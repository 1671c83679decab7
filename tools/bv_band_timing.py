"""clock64 phase breakdown of bv_band (needs a SFA_DEBUG_TIMING=1 build of the library)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfa_b200
from sfa_b200 import _lib, fast
import sfa_oracle as O
B, N = 64, int(os.environ.get("N", 250000))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
p = torch.empty((B, N, 4), device=dev)
p[..., 0] = torch.rand((B, N), generator=g, device=dev) * 80 - 40
p[..., 1] = torch.rand((B, N), generator=g, device=dev) * 80 - 40
p[..., 2] = torch.rand((B, N), generator=g, device=dev) * 4 - 3
p[..., 3] = torch.randint(0, 256, (B, N), generator=g, device=dev).float()
rast = fast.BvFeatureRasterizer(O.ARGO_BV_DISCRETIZATION, O.ARGO_BV_BOUNDARY, max_batch=B)
out = torch.empty((B, 3, rast.H, rast.W), device=dev)
for _ in range(3):
    rast(p, out=out)
torch.cuda.synchronize()
fn = _lib.load().sfa_debug_band_timing
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn(None, 1)
reps = 10
for _ in range(reps):
    rast(p, out=out)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
fn(buf, 0)
names = ["barrier + cursor", "prefetched records + atomics", "remaining records", "barrier", "normalise + store pass"]
items = reps * B * 128
tot = sum(buf[:5])
for k, n in enumerate(names):
    print("%-30s %8.0f cycles/item %5.1f%%" % (n, buf[k] / items, 100.0 * buf[k] / tot))
print("total per item %.0f cycles" % (tot / items))
