"""One deformable block at the size of the largest DCN layers of the plan (ida_2 node: 16 x 96 x 320 pixels, 2 x 64 channels in,
64 out) through the operator entry -- the launch ncu captures for the columns kernel.   python scripts/prof_dcn_columns.py [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocon_pytorch_b200 import engine as E      # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device='cpu').manual_seed(1)
B, C, H, W, Cout = 16, 128, 96, 320, 64
x = torch.randn(B, C, H, W, generator=g).to(dev)
off = (torch.randn(B, 18, H, W, generator=g) * 0.5).to(dev)
mask = torch.sigmoid(torch.randn(B, 9, H, W, generator=g)).to(dev)
w = (torch.randn(Cout, C, 3, 3, generator=g) * 0.03).to(dev)
for precision in sys.argv[1:] or ['fp32', 'bf16']:
    for _ in range(2):
        y = E.deform_conv2d(x, off, mask, w, None, split=2, precision=precision)
    torch.cuda.synchronize()
    print(precision, float(y.abs().max()))
