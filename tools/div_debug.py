import sys, ctypes, importlib, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
P = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
L = importlib.import_module(P + "._lib"); lib = L.load()
dev = torch.device("cuda", 0)
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def run(d, first, count):
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    L.check(lib.sfa_selftest_division(float(np.float32(d)), first, count, ctypes.c_void_p(bad.data_ptr()), stream))
    torch.cuda.synchronize()
    return int(bad.item())
bits = lambda v: int(np.float32(v).view(np.uint32))
for d in (50 / 608, 0.1, 1.0, 3.0):
    print("d =", d)
    for e in range(-32, 14, 2):
        n = run(d, bits(2.0 ** e), bits(2.0 ** (e + 2)) - bits(2.0 ** e))
        if n: print("   [2^%d, 2^%d): %d mismatches" % (e, e + 2, n))
    print("   denormal/tiny:", run(d, 0, 1 << 24), " huge:", run(d, bits(2.0 ** 90), 0x7FC00010 - bits(2.0 ** 90)))
