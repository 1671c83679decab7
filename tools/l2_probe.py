"""Does data written by one kernel stay in L2 for the next kernel to read?  (B200, 126 MB L2)"""
import torch
dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        pre = fn(None)
        e0.record(); fn(pre); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for mb in (8, 16, 32, 64, 96, 128, 256):
    buf = torch.empty(mb << 18, dtype=torch.float32, device=dev)
    def warm(pre):
        if pre is None:
            buf.fill_(1.0); return 1
        return buf.sum()
    def cold(pre):
        if pre is None:
            buf.fill_(1.0); flush.fill_(0); return 1
        return buf.sum()
    def readtwice(pre):
        if pre is None:
            flush.fill_(0); buf.sum(); return 1
        return buf.sum()
    tw, tc, tr = t(warm), t(cold), t(readtwice)
    print("%4d MB: read after write %.1f us (%.0f GB/s) | after write+flush %.1f us (%.0f GB/s) | after read %.1f us (%.0f GB/s)" %
          (mb, tw, mb * 1.048576e6 / tw / 1e3, tc, mb * 1.048576e6 / tc / 1e3, tr, mb * 1.048576e6 / tr / 1e3))
