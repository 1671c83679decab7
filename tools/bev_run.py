"""Minimal BEV workload for profiling: 64 uniform KITTI sweeps (2 alternating input sets, 246 MB > L2) through
BevRasterizer, `iters` times on one stream; prints us per 64 frames (CUDA events).  Usage: bev_run.py [iters] [algorithm]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
fast = importlib.import_module(P + ".fast"); geometry = importlib.import_module(P + ".geometry")
cnf = importlib.import_module(P + ".config.kitti_config")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(os.environ.get("SFA_N", "120000"))
dev = torch.device("cuda", 0)
B = 64
rng = np.random.default_rng(0)
def sweeps():
    a = np.empty((B, N, 4), np.float32)
    a[..., 0] = rng.uniform(0, 50, (B, N)); a[..., 1] = rng.uniform(-25, 25, (B, N))
    a[..., 2] = rng.uniform(-2.73, 1.27, (B, N)); a[..., 3] = rng.uniform(0, 1, (B, N))
    return torch.from_numpy(a).to(dev)
sets = [sweeps(), sweeps()]
rast = fast.BevRasterizer(geometry.from_config(cnf, algorithm=algo), max_batch=B, max_points=N, device=dev)
out = torch.empty((B, 3, 608, 608), device=dev)
for it in range(3): rast.rasterize_uniform(sets[it & 1], out=out)
torch.cuda.synchronize()
torch.cuda._sleep(5_000_000)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(iters): rast.rasterize_uniform(sets[it & 1], out=out)
e1.record(); torch.cuda.synchronize()
print("%.1f us per 64 frames (N=%d, algorithm %d)" % (e0.elapsed_time(e1) / iters * 1e3, N, algo))
