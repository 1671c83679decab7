import sys, ctypes, importlib, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import sfa_oracle as O
P = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
fast = importlib.import_module(P + ".fast"); lib = importlib.import_module(P + "._lib").load()
geom = importlib.import_module(P + ".geometry").from_config(importlib.import_module(P + ".config.kitti_config"))
dev = torch.device("cuda", 0)
B, N = 64, 120000
rng = np.random.default_rng(0)
pts = np.stack([rng.uniform(0, 50, (B, N)), rng.uniform(-25, 25, (B, N)), rng.uniform(-2.73, 1.27, (B, N)), rng.uniform(0, 1, (B, N))], -1).astype(np.float32)
pts = torch.from_numpy(pts).to(dev)
rast = fast.BevRasterizer(geom, max_batch=B, max_points=N, device=dev)
out = torch.empty((B, 3, 608, 608), device=dev)
for _ in range(3): rast.rasterize_uniform(pts, out=out)
torch.cuda.synchronize()
fn = lib.sfa_debug_band_timing; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn(None, 1)
reps = 10
for _ in range(reps): rast.rasterize_uniform(pts, out=out)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)(); fn(buf, 0)
names = ["clear+barrier", "records wait + phase1", "phase2", "phase3+fence", "prefetch+store issue", "TMA read-out wait"]
items = reps * B * 128
tot = sum(buf[:6])
for k, nme in enumerate(names):
    print("%-24s %8.0f cycles/item  %5.1f%%" % (nme, buf[k] / items, 100.0 * buf[k] / tot))
print("total per item %.0f cycles; items per CTA %.1f" % (tot / items, B * 128 / 592))

bnames = ["offsets + load issue + barrier", "points wait + cells + histogram", "scan + global atomics", "stage", "copy-out issue"]
tiles = reps * B * ((N + 2047) // 2048)
tot = sum(buf[8:13])
for k, nme in enumerate(bnames):
    print("bin %-32s %8.0f cycles/tile  %5.1f%%" % (nme, buf[8 + k] / tiles, 100.0 * buf[8 + k] / tot))
print("bin total per tile %.0f cycles" % (tot / tiles))
