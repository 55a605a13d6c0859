"""One profiled eager forward+decode pass of the bench workload (batch 16, 384x1280) -- the target of the ncu passes.
Warm-up (scale calibration + one pass) runs outside the profiled range: start ncu with `--profile-from-start off`.
PROF_PRECISION = fp32 (default: the fp32-accurate tensor-core mode) | bf16 | fp32_simt; PROF_BATCH; PROF_PASSES."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
import bench                                   # noqa: E402
from monocon_pytorch_b200 import engine as E   # noqa: E402

B = int(os.environ.get('PROF_BATCH', '16'))
precision = os.environ.get('PROF_PRECISION', 'fp32')
dev = torch.device('cuda', 0)
sd = bench.synthetic_state_dict()
eng = E.Engine(dev, B, bench.H, bench.W, precision)
eng.load_state_dict(sd)
img = bench.synthetic_frames(B, 1).to(dev)
P2 = torch.from_numpy(bench.kitti_p2(B)).to(dev)
invP = E.inverse_viewpad(bench.kitti_p2(B)).to(dev)
if eng.tensor_core_fp32:
    eng.calibrate_scales(img)
eng.infer_device(img, P2, invP)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(int(os.environ.get('PROF_PASSES', '1'))):
    eng.infer_device(img, P2, invP)
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('precision', precision, 'kernels per pass:', eng.kernel_launches)
