import sys, os, importlib, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import sfa_oracle as O
fast = importlib.import_module("lidar-image_object-detection_-fpn_resnet-yolov8_b200.fast")
dev = torch.device('cuda', 0)
for B in (1, 4, 8, 16, 33, 64, 128, 256):
    heads = [t.to(dev) for t in O.synth_heads(1, B=B)]
    out = torch.empty((B, 50, 10), device=dev)
    for _ in range(5): fast.decode_device(*heads, K=50, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fast.decode_device(*heads, K=50, out=out)
    e1.record(); torch.cuda.synchronize()
    print("slabs=%s B=%3d  %.2f us/launch" % (os.environ.get("SFA_DECODE_SLABS", "auto"), B, e0.elapsed_time(e1) / 50 * 1e3), flush=True)
