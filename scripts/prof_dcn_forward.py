"""One eager forward + decode of the DCN detector at the bench shape (B = 16, 384x1280) for an ncu launch list.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/dcn_launches.csv python scripts/prof_dcn_forward.py [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocon_pytorch_b200 import engine as E      # noqa: E402
from oracle import fixtures as FX                 # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
dev = torch.device('cuda', 0)
B, H, W = 16, 384, 1280
sd = FX.make_state_dict(0, use_dcn=True)
img = FX.make_images(B, H, W, seed=100).to(dev)
P2h = FX.kitti_p2(B, 7)
P2, invP = torch.from_numpy(P2h).to(dev), E.inverse_viewpad(P2h).to(dev)
eng = E.Engine(dev, B, H, W, precision, use_dcn=True)
eng.load_state_dict(sd)
if eng.tensor_core_fp32:
    eng.calibrate_scales(img)
torch.cuda.synchronize()
print('PROFILE_BEGIN', flush=True)
eng.infer_device(img, P2, invP)
torch.cuda.synchronize()
print('PROFILE_END', flush=True)
eng.close()
