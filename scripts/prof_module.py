"""cProfile of MonoConDetector.batch_eval (the reference's call surface) on the GPU: where the host time of a call goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                         # noqa: E402
import torch                                               # noqa: E402
import monocon_pytorch_b200 as M                           # noqa: E402
from oracle import fixtures as FX                          # noqa: E402

B, H, W = 16, 384, 1280
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False, precision='fp32', max_batch=B).to(dev).eval()


class _Calib:
    def __init__(self, p):
        self.P2 = p


P2_np = FX.kitti_p2(B, 19)
calibs = [_Calib(p) for p in P2_np]
metas = {'pad_shape': [(H, W)] * B, 'ori_shape': [(H, W)] * B, 'sample_idx': list(range(B))}
img = (torch.randn(B, 3, H, W) * 0.01).to(dev)
data = {'img': img, 'img_metas': metas, 'calib': calibs}
for _ in range(3):
    model.batch_eval(data, get_vis_format=False)
model.freeze_engine(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    model.batch_eval(data, get_vis_format=False)
torch.cuda.synchronize()
print('ms per call (frames already on the device):', (time.perf_counter() - t0) / 30 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    model.batch_eval(data, get_vis_format=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
