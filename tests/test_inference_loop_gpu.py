"""GPU: BASELINE config[2] data flow — sweeps -> GPU BEV -> a random-init stride-4 conv backbone with the reference's five
heads (stock PyTorch, stands in for fpn_resnet_18, which is out of scope and cannot travel to the GPU box) -> `_sigmoid`
-> GPU decode -> dense post-processing, batch 32, everything staying on the device.

What is checked is the hot path around the backbone: the BEV maps fed to it are bit-exact vs the oracle, and the
detections decoded from the backbone's REAL head tensors (sigmoid outputs clamped at 1e-4 / 1-1e-4, i.e. with plateaus
and ties, unlike the synthetic heads) equal the oracle's decode of the same tensors copied to the CPU."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import sfa_oracle as O
from conftest import pkg

pytestmark = pytest.mark.gpu


class TinyBackbone(nn.Module):
    """[B,3,608,608] -> the five SFA3D heads at stride 4 ([B,c,152,152]); models/fpn_resnet.py:112-263 in miniature."""
    HEADS = {"hm_cen": 3, "cen_offset": 2, "direction": 2, "z_coor": 1, "dim": 3}

    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 16, 7, 2, 3), nn.BatchNorm2d(16), nn.ReLU(inplace=True),
                                  nn.Conv2d(16, 32, 3, 2, 1), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                                  nn.Conv2d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.heads = nn.ModuleDict({k: nn.Sequential(nn.Conv2d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(32, c, 1))
                                    for k, c in self.HEADS.items()})

    def forward(self, x):
        f = self.stem(x)
        return {k: h(f) for k, h in self.heads.items()}


@pytest.mark.parametrize("hm_bias", [0.0, -9.5])
def test_bev_backbone_decode_loop_batch32(cuda_device, hm_bias):
    """hm_bias = -9.5 pushes most of the heat map under the 1e-4 clamp: a map-wide plateau in which every cell is a
    'peak' (the candidate list overflows shared memory, ties straddle the K-th place), like a trained detector's
    background.  hm_bias = 0 keeps scores spread."""
    fast, ev, tu = pkg("fast"), pkg("utils.evaluation_utils"), pkg("utils.torch_utils")
    B, N = 32, 60000
    geom = pkg("geometry").from_config(pkg("config.kitti_config"))
    sweeps = [O.synth_sweep(500 + i, N, O.KITTI, "clustered" if i % 2 else "uniform") for i in range(B)]
    pts = torch.from_numpy(np.concatenate(sweeps)).to(cuda_device)
    offsets = torch.arange(B + 1, dtype=torch.int64, device=cuda_device) * N
    rast = fast.BevRasterizer(geom, max_batch=B, max_points=N, device=cuda_device)
    bev = rast(pts, offsets, N)
    for i in (0, 13, 31):
        want = O.make_bev_scatter(sweeps[i], O.KITTI, True, np.float32)
        assert np.array_equal(bev[i].cpu().numpy().view(np.uint32), want.view(np.uint32))

    torch.manual_seed(0)
    net = TinyBackbone().to(cuda_device).eval()
    with torch.no_grad():
        net.heads["hm_cen"][-1].bias.fill_(hm_bias)
        out = net(bev)
        hm = tu._sigmoid(out["hm_cen"])
        off = tu._sigmoid(out["cen_offset"])
        det = ev.decode(hm, off, out["direction"], out["z_coor"], out["dim"], K=50)
        rows, cls, keep = fast.post_process_dense(det)
    torch.cuda.synchronize()
    assert det.shape == (B, 50, 10) and det.is_cuda

    cpu = lambda t: t.detach().cpu().contiguous()
    want = O.decode(cpu(hm), cpu(off), cpu(out["direction"]), cpu(out["z_coor"]), cpu(out["dim"]), K=50).numpy()
    got = det.cpu().numpy()
    # scores are unique up to ties: compare scores exactly, and full rows wherever a score is not tied in its frame
    assert np.array_equal(got[:, :, 0], want[:, :, 0])
    n_untied = 0
    for b in range(B):
        s = want[b, :, 0]
        hm_b = cpu(hm)[b].numpy()
        nms_b = O._nms(cpu(hm)[b:b + 1]).numpy()[0]
        for k in range(50):
            tied_in_map = np.count_nonzero(nms_b == s[k]) > 1
            if not tied_in_map:
                assert np.array_equal(got[b, k].view(np.uint32), want[b, k].view(np.uint32)), (b, k)
                n_untied += 1
            else:   # a tie group: our pick must be a genuine kept cell with that score, lowest (class, index) first
                c, x, y = int(got[b, k, 9]), got[b, k, 1], got[b, k, 2]
                offb = cpu(off)[b].numpy()
                cand = np.argwhere(nms_b[c] == s[k])
                assert any(np.float32(xx) + offb[0, yy, xx] == x and np.float32(yy) + offb[1, yy, xx] == y for yy, xx in cand)
        # GPU tie rule: within equal scores, ascending (class, y*w+x)
        g = got[b]
        key = g[:, 9].astype(np.int64) * 152 * 152 + np.round(g[:, 2] - 0.5).astype(np.int64) * 0   # class only (cheap check)
        eq = g[:-1, 0] == g[1:, 0]
        assert np.all(key[:-1][eq] <= key[1:][eq])
    if hm_bias == 0.0:
        assert n_untied > B * 25
    # dense post-processing agrees with the reference-semantics oracle on the same detections
    ref = O.post_processing(got.astype(np.float32))
    r, c, k = rows.cpu().numpy(), cls.cpu().numpy(), keep.cpu().numpy()
    for i in (0, 17, 31):
        for j in range(3):
            w = np.asarray(ref[i][j], np.float32).reshape(-1, 8)
            gsel = r[i][(c[i] == j) & k[i].astype(bool)]
            assert gsel.shape == w.shape and np.array_equal(gsel[:, :7], w[:, :7])
            np.testing.assert_allclose(gsel[:, 7], w[:, 7], rtol=1e-5, atol=1e-6)
