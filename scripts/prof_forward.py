"""Two eager forward+decode passes of the bench workload (batch 16, 384x1280, bf16) -- the target of the ncu passes.
The first pass is the warm-up; profile the second one (ncu -s <launches of pass 1>)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
import bench                                   # noqa: E402
from monocon_pytorch_b200 import engine as E   # noqa: E402

B = int(os.environ.get('PROF_BATCH', '16'))
dev = torch.device('cuda', 0)
sd = bench.synthetic_state_dict()
eng = E.Engine(dev, B, bench.H, bench.W, 'bf16')
eng.load_state_dict(sd)
img = bench.synthetic_frames(B, 1).to(dev)
P2 = torch.from_numpy(bench.kitti_p2(B)).to(dev)
invP = E.inverse_viewpad(bench.kitti_p2(B)).to(dev)
for _ in range(int(os.environ.get('PROF_PASSES', '2'))):
    eng.infer_device(img, P2, invP)
    torch.cuda.synchronize()
print('kernels per pass:', eng.kernel_launches)
