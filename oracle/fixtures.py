"""Seeded synthetic fixtures for the parity tests.  TEST INFRASTRUCTURE ONLY (see monocon_oracle.py).

There is no dataset, checkpoint or network access, and the reference ships no
golden vectors, so every test input is generated from a seed:

* ``param_table()``      -- (key, shape, dtype) of the reference state_dict (242 parameter
                            tensors + BN buffers; checked against the real reference in
                            tests/golden/gen_golden.py and against the product module in tests/).
* ``make_state_dict()``  -- a seeded, *calibrated* random network: He-normal conv weights,
                            random BN affines, and BN running statistics taken from one
                            calibration batch, so that every BN fold is non-trivial, the
                            activations stay O(1) at every depth and the centre heat-map is
                            not saturated (SURVEY.md §0 fact 3 explains why the reference's
                            own random init + N(0,1) input is unusable for top-k parity).
* ``make_images()``, ``kitti_p2()`` -- seeded inputs / a typical KITTI projection matrix.

Everything is generated on the CPU with ``torch.Generator`` so that the build
container and the GPU box produce the same tensors.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import monocon_oracle as O

HEAD_OUT = {'heatmap_head': 3, 'wh_head': 2, 'offset_head': 2, 'center2kpt_offset_head': 18,
            'kpt_heatmap_head': 9, 'kpt_heatmap_offset_head': 2, 'dim_head': 3, 'depth_head': 2}


def _bn_entries(prefix: str, c: int, affine: bool = True):
    out = []
    if affine:
        out += [(prefix + '.weight', (c,), torch.float32), (prefix + '.bias', (c,), torch.float32)]
    out += [(prefix + '.running_mean', (c,), torch.float32), (prefix + '.running_var', (c,), torch.float32),
            (prefix + '.num_batches_tracked', (), torch.int64)]
    return out


def _block_entries(prefix: str, cin: int, cout: int):
    return ([(prefix + '.conv1.weight', (cout, cin, 3, 3), torch.float32)] + _bn_entries(prefix + '.bn1', cout) +
            [(prefix + '.conv2.weight', (cout, cout, 3, 3), torch.float32)] + _bn_entries(prefix + '.bn2', cout))


def _tree_entries(prefix: str, levels: int, cin: int, cout: int, level_root: bool, root_dim: int = 0):
    """Parameter registration order of Tree.__init__ (dla.py:148-185)."""
    if root_dim == 0:
        root_dim = 2 * cout
    if level_root:
        root_dim += cin
    out = []
    if levels == 1:
        out += _block_entries(prefix + '.tree1', cin, cout)
        out += _block_entries(prefix + '.tree2', cout, cout)
        out += [(prefix + '.root.conv.weight', (cout, root_dim, 1, 1), torch.float32)]
        out += _bn_entries(prefix + '.root.bn', cout)
    else:
        out += _tree_entries(prefix + '.tree1', levels - 1, cin, cout, False, 0)
        out += _tree_entries(prefix + '.tree2', levels - 1, cout, cout, False, root_dim + cout)
    if cin != cout:
        out += [(prefix + '.project.0.weight', (cout, cin, 1, 1), torch.float32)]
        out += _bn_entries(prefix + '.project.1', cout)
    return out


def param_table(use_dcn: bool = False) -> List[Tuple[str, tuple, torch.dtype]]:
    ch, lv = O.DLA34_CHANNELS, O.DLA34_LEVELS
    t = [('backbone.base_layer.0.weight', (16, 3, 7, 7), torch.float32)] + _bn_entries('backbone.base_layer.1', 16)
    t += [('backbone.level0.0.weight', (16, 16, 3, 3), torch.float32)] + _bn_entries('backbone.level0.1', 16)
    t += [('backbone.level1.0.weight', (32, 16, 3, 3), torch.float32)] + _bn_entries('backbone.level1.1', 32)
    for l in range(2, 6):
        t += _tree_entries(f'backbone.level{l}', lv[l], ch[l - 1], ch[l], level_root=(l != 2))
    # DLAUp over channels (64,128,256,512): ida_0 out 256 (1 step), ida_1 out 128 (2), ida_2 out 64 (3)
    for i, (cout, cins) in enumerate(((256, (512,)), (128, (256, 256)), (64, (128, 128, 128)))):
        for j, cin in enumerate(cins, start=1):
            p = f'neck.ida_{i}'
            t += [(f'{p}.proj_{j}.conv.weight', (cout, cin, 3, 3), torch.float32)] + _bn_entries(f'{p}.proj_{j}.bn1', cout)
            t += [(f'{p}.up_{j}.weight', (cout, 1, 4, 4), torch.float32)]
            t += [(f'{p}.node_{j}.conv.weight', (cout, 2 * cout, 3, 3), torch.float32)] + _bn_entries(f'{p}.node_{j}.bn1', cout)
            if use_dcn:          # DCNv2 pack: the offset / mask convolution of each block (27 = 18 offsets + 9 mask logits)
                t += [(f'{p}.proj_{j}.conv.conv_offset.weight', (27, cin, 3, 3), torch.float32), (f'{p}.proj_{j}.conv.conv_offset.bias', (27,), torch.float32)]
                t += [(f'{p}.node_{j}.conv.conv_offset.weight', (27, 2 * cout, 3, 3), torch.float32), (f'{p}.node_{j}.conv.conv_offset.bias', (27,), torch.float32)]
    for name in O.HEAD_STEMS:
        p = f'head.{name}'
        t += [(p + '.0.weight', (64, 64, 3, 3), torch.float32), (p + '.0.bias', (64,), torch.float32)]
        t += [(p + '.1.weight_', (10, 64), torch.float32), (p + '.1.bias_', (10, 64), torch.float32)]
        t += _bn_entries(p + '.1', 64, affine=False)
        t += [(p + '.1.attn_weights.attention.0.weight', (10, 64, 1, 1), torch.float32)]
        t += _bn_entries(p + '.1.attn_weights.attention.1', 10)
        if name in HEAD_OUT:
            t += [(p + '.3.weight', (HEAD_OUT[name], 64, 1, 1), torch.float32), (p + '.3.bias', (HEAD_OUT[name],), torch.float32)]
    for name in ('dir_cls', 'dir_reg'):
        t += [(f'head.{name}.0.weight', (12, 64, 1, 1), torch.float32), (f'head.{name}.0.bias', (12,), torch.float32)]
    return t


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_images(batch: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """Seeded N(0,1) frames, (B,3,H,W) fp32 -- the distribution of a Normalize()d KITTI image."""
    g = _gen(f'img{batch}x{h}x{w}', seed)
    return torch.randn(batch, 3, h, w, generator=g, dtype=torch.float32)


def kitti_p2(batch: int, seed: int = 0) -> np.ndarray:
    """Typical KITTI P2, slightly perturbed per image so per-image calibration is exercised."""
    base = np.array([[721.5377, 0., 609.5593, 44.85728],
                     [0., 721.5377, 172.854, 0.2163791],
                     [0., 0., 1., 0.002745884]], dtype=np.float32)
    rng = np.random.RandomState(1000 + seed)
    out = np.repeat(base[None], batch, 0)
    out[:, 0, 0] += rng.uniform(-5, 5, batch).astype(np.float32)
    out[:, 1, 1] = out[:, 0, 0]
    out[:, 0, 2] += rng.uniform(-8, 8, batch).astype(np.float32)
    out[:, 1, 2] += rng.uniform(-4, 4, batch).astype(np.float32)
    return out.astype(np.float32)


def make_state_dict(seed: int = 0, calibrate: bool = True, calib_hw: Tuple[int, int] = (128, 256), use_dcn: bool = False,
                    dcn_offset_gain: float = 0.1) -> Dict[str, torch.Tensor]:
    """``use_dcn``: also the conv_offset parameters of the DCN neck variant: He-normal times ``dcn_offset_gain`` plus N(0, 0.2)
    biases.  The default gain gives offsets of a few tenths of a pixel and mask logits of the same size (a DCNv2 pack starts from
    ZERO offset weights, and a zero-initialised pack would test nothing).  Gain 1 -- pixel-scale offsets computed from white-noise
    random-init features -- makes every deformable block amplify an input perturbation 2-5x (d column / d offset = the spatial
    gradient of noise): the strict fp32 FFMA twin, whose operators sit 1e-6 from torch, then ends 1e-2 from the reference at
    384x1280 (profiles/r02_dcn_error_growth.txt), so that fixture only serves the stage-wise tests."""
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, dtype in param_table(use_dcn):
        g = _gen(key, seed)
        if dtype == torch.int64:
            sd[key] = torch.tensor(1, dtype=torch.int64)
        elif key.endswith('running_mean'):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith('running_var'):
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith('.weight_'):                      # AttnBN mixture banks, attentive_norm.py:150-152
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith('.bias_'):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif '.up_' in key:                                 # bilinear 4x4 (dla_neck.py:83-92), perturbed: it is trainable
            k1 = torch.tensor([0.25, 0.75, 0.75, 0.25])
            w = (k1[:, None] * k1[None, :]).expand(shape).clone()
            sd[key] = w * (1.0 + 0.2 * torch.randn(shape, generator=g))
        elif len(shape) == 4:                               # conv weights: He-normal on fan-in
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * float(np.sqrt(2.0 / fan_in))
            if key.endswith('.conv_offset.weight'):
                sd[key] = sd[key] * dcn_offset_gain
        elif key.endswith('.weight'):                       # BN gamma
            sd[key] = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif key.endswith('.bias'):                         # BN beta / conv bias
            sd[key] = 0.2 * torch.randn(shape, generator=g)
        else:
            raise KeyError(key)
    # final 1x1 convs: AttnBN's mixture gain is ~5 at init (a ~ 0.5 over 10 banks of N(1,0.1)), so scale
    # the output convs down to get O(1) maps; heat-maps keep the reference's prior bias -log(9)
    # (monocon_heads.py:134-137) with logits spread enough to cross test_thres=0.4 but rarely clamp.
    for key in list(sd.keys()):
        if key.startswith('head.') and key.endswith('.weight') and sd[key].dim() == 4 and sd[key].shape[-1] == 1 \
                and 'attn_weights' not in key:
            sd[key] = sd[key] * 0.25
    for name in ('heatmap_head', 'kpt_heatmap_head'):
        sd[f'head.{name}.3.bias'] = torch.full_like(sd[f'head.{name}.3.bias'], -2.1972246)
        sd[f'head.{name}.3.weight'] = sd[f'head.{name}.3.weight'] * 1.5
    sd['head.dim_head.3.bias'] = sd['head.dim_head.3.bias'] + torch.tensor([1.6, 1.6, 3.9])
    sd['head.depth_head.3.weight'] = sd['head.depth_head.3.weight'] * 0.5
    if calibrate:
        img = make_images(2, calib_hw[0], calib_hw[1], seed=seed + 7919)
        O.forward(sd, img, calibrate=True)
        # output calibration: per-class heat logits ~ N(-3, 0.7) (peaks of a 384x1280 frame land around
        # 0.4-0.6, clear of the clamp ceiling, so top-k is tie-free) and uncertainty channel ~ N(0.3, 0.3)
        # (sigma = exp(-d1) ~ 0.75, so score*sigma straddles test_thres = 0.4).
        pred = O.forward(sd, img)

        def renorm(conv: str, rows, values, mean: float, std: float) -> None:
            w, b = sd[conv + '.weight'], sd[conv + '.bias']
            for r, v in zip(rows, values):
                m, s = float(v.mean()), float(v.std())
                a = std / max(s, 1e-6)
                w[r] = w[r] * a
                b[r] = mean - a * (m - float(b[r]))

        for conv, key in (('head.heatmap_head.3', 'center_heatmap_pred'), ('head.kpt_heatmap_head.3', 'kpt_heatmap_pred')):
            p = pred[key].clamp(1e-4, 1 - 1e-4)
            logit = torch.log(p / (1 - p))
            renorm(conv, range(logit.shape[1]), [logit[:, c] for c in range(logit.shape[1])], -3.0, 0.7)
        renorm('head.depth_head.3', [1], [pred['depth_pred'][:, 1]], 0.3, 0.3)
    return {k: v.contiguous() for k, v in sd.items()}
