"""DRAM bytes of ONE whole step as a unit (no per-kernel serialisation, no cache flush between its kernels):
64 KITTI sweeps -> BEV (algorithm from argv) + decode + post-processing on one stream, bracketed by
cudaProfilerStart/Stop.  Run under
    ncu --replay-mode range --profile-from-start off --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum python tools/range_traffic.py <algorithm> [steps]
With engines > 1 the steps are dealt round-robin to that many independent engines (own rasteriser workspace, outputs and
CUDA stream), like bench.py's --pipelines: the range then holds the overlapped schedule the benchmark times.
Usage: range_traffic.py [algorithm=0] [steps_in_range=1] [engines=1]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfa_oracle as O   # synthetic heads only
P = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
fast = importlib.import_module(P + ".fast"); geometry = importlib.import_module(P + ".geometry")
cnf = importlib.import_module(P + ".config.kitti_config")
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_eng = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
B, N = 64, 120000
rng = np.random.default_rng(0)
def sweeps():
    a = np.empty((B, N, 4), np.float32)
    a[..., 0] = rng.uniform(0, 50, (B, N)); a[..., 1] = rng.uniform(-25, 25, (B, N))
    a[..., 2] = rng.uniform(-2.73, 1.27, (B, N)); a[..., 3] = rng.uniform(0, 1, (B, N))
    return torch.from_numpy(a).to(dev)
sets = [sweeps() for _ in range(3)]
heads = [tuple(t.to(dev) for t in O.synth_heads(7 + i, B=B)) for i in range(3)]
class Engine:
    def __init__(self):
        self.rast = fast.BevRasterizer(geometry.from_config(cnf, algorithm=algo), max_batch=B, max_points=N, device=dev)
        self.bev = torch.empty((B, 3, 608, 608), device=dev)
        self.det = torch.empty((B, 50, 10), device=dev)
        self.pp = (torch.empty((B, 50, 8), device=dev), torch.empty((B, 50), dtype=torch.int32, device=dev),
                   torch.empty((B, 50), dtype=torch.uint8, device=dev))
        self.ws = fast.DecodeWorkspace(dev, B, 3, 152, 152, 50)
        self.stream = torch.cuda.Stream(device=dev)

    def step(self, i):
        with torch.cuda.stream(self.stream):
            self.rast.rasterize_uniform(sets[i % 3], out=self.bev)
            fast.decode_device(*heads[i % 3], K=50, out=self.det, workspace=self.ws, post=self.pp)


engines = [Engine() for _ in range(n_eng)]
torch.cuda.synchronize()
for i in range(2 * n_eng + 2): engines[i % n_eng].step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(steps): engines[i % n_eng].step(4 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("range of %d step(s) on %d engine(s) done, algorithm %d" % (steps, n_eng, algo))
