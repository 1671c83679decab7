"""HBM bandwidth by direction on this B200: write-only (fill), read-only (sum), copy (read+write)."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 29   # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = n * 4 / 1e9
tw = t(lambda: a.fill_(1.0)); print("write-only  fill_      %.1f GB/s" % (gb / tw * 1e3))
tz = t(lambda: a.zero_());    print("write-only  zero_      %.1f GB/s" % (gb / tz * 1e3))
tr = t(lambda: a.sum());      print("read-only   sum        %.1f GB/s" % (gb / tr * 1e3))
tc = t(lambda: b.copy_(a));   print("copy        read+write %.1f GB/s" % (2 * gb / tc * 1e3))
tm = t(lambda: torch.cuda.memset if False else a.zero_())
