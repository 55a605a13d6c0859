"""Generates tests/golden/dcn.npz with torchvision's own operator (torchvision.ops.deform_conv2d, the pinned third-party algorithm
of oracle/dcn_oracle.py).  Run in the build container: python tests/golden/gen_dcn_golden.py"""
import os
import sys

import numpy as np
import torch
import torchvision
from torchvision.ops import deform_conv2d

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dcn_oracle as D      # noqa: E402  (the case generator only)

out = {'torchvision_version': np.array(torchvision.__version__)}
for n, (B, C, H, W, Cout, seed) in enumerate(D.GOLDEN_CASES):
    x, offset, mask, w, b = D.make_case(B, C, H, W, Cout, seed)
    y = deform_conv2d(x.double(), offset.double(), w.double(), b.double(), stride=1, padding=1, dilation=1, mask=mask.double())
    out[f'y{n}'] = y.numpy()
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'dcn.npz'), **out)
print('wrote dcn.npz', {k: getattr(v, 'shape', None) for k, v in out.items()})


# ---- the whole detector with the DCN neck: the UNMODIFIED reference model with the 3x3 convolution of every IDAUp Conv2dBlock
# (model/backbone/dla_neck.py:21-26) swapped for a DCNv2 pack built on torchvision's operator -> tests/golden/dcn_model.npz.
# load_state_dict(strict=True) proves the key layout of the variant (oracle/fixtures.py, use_dcn=True).
sys.path.insert(0, '/root/reference')
from model import MonoConDetector                      # noqa: E402  (the reference)
from model.backbone.dla_neck import Conv2dBlock        # noqa: E402  (the reference)
from utils.tensor_ops import get_local_maximum, get_topk_from_heatmap   # noqa: E402  (the reference)
from oracle import fixtures as FX                      # noqa: E402


class DCNv2Pack(torch.nn.Module):
    """mmcv ModulatedDeformConv2dPack / CenterNet DCN, 3x3 stride 1 pad 1, one offset group, no bias, on torchvision's operator."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.zeros(cout, cin, 3, 3))
        self.conv_offset = torch.nn.Conv2d(cin, 27, 3, 1, 1, bias=True)

    def forward(self, x):
        om = self.conv_offset(x)
        return deform_conv2d(x, om[:, :18], self.weight, None, stride=1, padding=1, dilation=1, mask=torch.sigmoid(om[:, 18:]))


model = MonoConDetector(34, pretrained_backbone=False).eval()
for m in model.modules():
    if isinstance(m, Conv2dBlock):
        m.conv = DCNv2Pack(m.conv.in_channels, m.conv.out_channels)
sd = FX.make_state_dict(0, use_dcn=True)
print(model.load_state_dict(sd, strict=True))
B, H, W = 2, 128, 256
img = FX.make_images(B, H, W, seed=1)


class _Calib:
    def __init__(self, p2):
        self.P2 = p2


data = {'img': img, 'img_metas': {'pad_shape': [(H, W)] * B}, 'calib': [_Calib(p) for p in FX.kitti_p2(B, seed=1)]}
with torch.no_grad():
    pred = model(data)
    nms = get_local_maximum(pred['center_heatmap_pred'].clone(), kernel=3)
    scores, inds, clses, ys, xs = get_topk_from_heatmap(nms, k=31)
out = {'pred/' + k: v.numpy().astype(np.float32) for k, v in pred.items()}
out['topk/scores'] = scores.numpy()
out['topk/inds'] = inds.numpy()
out['hw'] = np.array([H, W])
out['img_seed'] = np.array(1)
out['keys'] = np.array([f'{k} {tuple(v.shape)}' for k, v in model.state_dict().items() if 'conv_offset' in k])
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'dcn_model.npz'), **out)
print('wrote dcn_model.npz', {k: v.shape for k, v in out.items() if k.startswith('pred/')}, 'min top-31 score gap',
      float((scores[:, :-1] - scores[:, 1:]).min()))
