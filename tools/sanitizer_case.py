"""A small pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys, importlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sfa_oracle as O
P = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"
fast = importlib.import_module(P + ".fast"); geometry = importlib.import_module(P + ".geometry")
ev = importlib.import_module(P + ".utils.evaluation_utils"); L = importlib.import_module(P + "._lib")
cnf = importlib.import_module(P + ".config.kitti_config")
dev = torch.device("cuda", 0)
for algo in (L.BEV_TILED, L.BEV_TILED_TWO_KERNEL, L.BEV_GLOBAL_ATOMIC):   # fused persistent kernel, bin + band, global atomics
    geom = geometry.from_config(cnf, algorithm=algo)
    sweeps = [O.synth_sweep(1, 9000, O.KITTI, "zties"), O.synth_sweep(2, 7000, O.KITTI, "outside"),
              O.synth_sweep(3, 9000, O.KITTI, "onecell")]
    lens = [s.shape[0] for s in sweeps]
    pts = torch.from_numpy(np.concatenate(sweeps)).to(dev)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    rast = fast.BevRasterizer(geom, max_batch=3, max_points=max(lens), device=dev)
    got = rast(pts, offsets, max(lens)).cpu().numpy()
    for i, s in enumerate(sweeps):
        assert np.array_equal(got[i], O.make_bev_scatter(s, O.KITTI, True, np.float32)), (algo, i)
# sweep-side extras of the two-kernel path: augmentation prologue + flip, and front + back maps from one read
geom = geometry.from_config(cnf)
sweeps = [O.synth_sweep(5, 6000, O.KITTI, "outside"), O.synth_sweep(6, 4000, O.KITTI, "zties")]
lens = [s.shape[0] for s in sweeps]
pts = torch.from_numpy(np.concatenate(sweeps)).to(dev)
offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
rast = fast.BevRasterizer(geom, max_batch=2, max_points=max(lens), device=dev)
mats = torch.from_numpy(np.stack([np.stack(O.transform_matrices(0, 0, 0, rz=a)) for a in (0.3, -0.2)])).to(dev)
got = rast(pts, offsets, max(lens), mats=mats, scales=torch.tensor([1.02, 0.97], device=dev), hflip=torch.tensor([1, 0], dtype=torch.uint8, device=dev))
pair = fast.FrontBackRasterizer(max_batch=2, max_points=max(lens), device=dev)
front, back = pair(pts, offsets, max(lens))
for i, s in enumerate(sweeps):
    assert np.array_equal(front[i].cpu().numpy(), O.make_bev_scatter(s, O.KITTI, True, np.float32))
    assert np.array_equal(back[i].cpu().numpy(), O.make_bev_scatter(s, O.KITTI_BACK, True, np.float32))
heads = O.synth_heads(3, B=2, tie_free=True)
det = ev.decode(*[t.to(dev) for t in heads], K=50)
assert np.array_equal(det.cpu().numpy(), O.decode(*[t.clone() for t in heads], K=50).numpy())
hm = torch.full((1, 3, 152, 152), 0.5)
ev.decode(hm.to(dev), *[t[:1].to(dev) for t in heads[1:]], K=50)          # plateau: long list, tie-break paths
fast.post_process_dense(det, real=True)
post = (torch.empty((2, 50, 8), device=dev), torch.empty((2, 50), dtype=torch.int32, device=dev), torch.empty((2, 50), dtype=torch.uint8, device=dev))
fast.decode_device(*[t.to(dev) for t in heads], K=50, post=post)            # sfa_decode_post
ev._nms(heads[0].to(dev)); ev._topk(heads[0].to(dev), K=20)
f = importlib.import_module(P + ".data_process.kitti_data_utils").get_filtered_lidar(O.synth_sweep(4, 5000, O.KITTI, "outside"), O.KITTI.boundary)
torch.cuda.synchronize()
print("sanitizer case ok", f.shape)
