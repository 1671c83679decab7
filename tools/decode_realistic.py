"""Decode timing on plateau-heavy heat maps (sigmoid outputs clamped at 1e-4) vs spread scores."""
import sys, importlib, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import sfa_oracle as O
fast = importlib.import_module("lidar-image_object-detection_-fpn_resnet-yolov8_b200.fast")
dev = torch.device('cuda', 0)
B = 64
def run(name, hm):
    _, off, d, z, dim = [t.to(dev) for t in O.synth_heads(1, B=B)]
    hm = hm.to(dev).contiguous()
    out = torch.empty((B, 50, 10), device=dev)
    for _ in range(3): fast.decode_device(hm, off, d, z, dim, K=50, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(5_000_000)
    e0.record()
    for _ in range(20): fast.decode_device(hm, off, d, z, dim, K=50, out=out)
    e1.record(); torch.cuda.synchronize()
    print("%-44s %8.1f us per 64 frames" % (name, e0.elapsed_time(e1) / 20 * 1e3))
g = torch.Generator().manual_seed(0)
run("uniform scores (bench workload)", O.synth_heads(2, B=B)[0])
x = torch.randn(B, 3, 152, 152, generator=g) * 2 - 9.5
run("sigmoid, 97% of cells clamped to 1e-4", torch.clamp(torch.sigmoid(x), 1e-4, 1 - 1e-4))
x = torch.randn(B, 3, 152, 152, generator=g) * 2 - 14
run("sigmoid, all but ~0.1% clamped (sparse scene)", torch.clamp(torch.sigmoid(x), 1e-4, 1 - 1e-4))
run("constant map", torch.full((B, 3, 152, 152), 0.5))
run("7 distinct values", torch.randint(1, 8, (B, 3, 152, 152), generator=g).float() / 8)
