"""CPU oracle for the BACKWARD half of the training step, written without autograd.  TEST INFRASTRUCTURE ONLY.

``oracle.monocon_oracle.train_step`` gets its gradients from torch autograd, which pins *what* the gradients are (it is
checked against the unmodified reference's own step, tests/golden/train_step.npz) but says nothing about *how* a kernel
computes them.  This file is the other half: every layer kind of the detector has an explicit backward formula here
(`*_backward` functions, one per CUDA kernel the training engine needs), and a tape replays them in reverse over the
train-mode forward of reference ``MonoConDetector.forward`` (model/detector/monocon_detector.py:53-61).  The gradients it
produces are checked against the same reference digests as the autograd oracle (tests/test_backward_oracle.py), so each
formula below is a pinned per-kernel checker for DESIGN.md §9 item 1.

Layer kinds and where the reference uses them (all file:line relative to /root/reference):
  conv2d (+bias)            model/backbone/dla.py:22-31, 117-121, 228-236; dla_neck.py:24-31; monocon_heads.py:114-131
  BatchNorm2d, train mode   dla.py:24,30,119,185,233,295; dla_neck.py:27
  ReLU / residual add / cat dla.py:34-51, 124-132
  MaxPool2d(s, stride=s)    dla.py:176-177, 193
  depthwise ConvTranspose2d dla_neck.py:58-65  (k = 4, s = 2, p = 1 for every use in DLAUp)
  AttnBatchNorm2d           model/norm/attentive_norm.py:79-91, 154-164; monocon_heads.py:117
  output transforms         monocon_heads.py:168-170 (sigmoid + clamp), :183 (depth)

Only ``tests/`` may import this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .monocon_oracle import DLA34_CHANNELS, DLA34_LEVELS, EPS, HEAD_STEMS, PRED_KEYS


# --------------------------------------------------------------------------------------
# per-kernel backward formulas
# --------------------------------------------------------------------------------------
def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor, x_hw: Tuple[int, int], stride: int, padding: int) -> torch.Tensor:
    """dL/dx of y = conv2d(x, w, stride, padding):  dx[n,ci,iy,ix] = sum_{co,ky,kx} dy[n,co,oy,ox] * w[co,ci,ky,kx] over the
    (oy, ox) with oy*stride + ky - padding == iy (same for x).  Written as a gather over the output so that the CUDA
    kernel is a plain convolution of the (zero-dilated for stride 2) dy with the 180-degree-rotated, transposed filter."""
    n, co, oh, ow = dy.shape
    _, ci, k, _ = w.shape
    ih, iw = x_hw
    if stride > 1:                                            # zero-dilate dy to the input grid
        d = dy.new_zeros(n, co, (oh - 1) * stride + 1, (ow - 1) * stride + 1)
        d[:, :, ::stride, ::stride] = dy
    else:
        d = dy
    wt = w.flip(2, 3).transpose(0, 1).contiguous()            # (ci, co, k, k), rotated
    lo = k - 1 - padding
    hi_h = ih - (d.shape[2] + 2 * lo - (k - 1)) + lo          # rows the strided forward never reached stay zero
    hi_w = iw - (d.shape[3] + 2 * lo - (k - 1)) + lo
    d = F.pad(d, (lo, hi_w, lo, hi_h))
    return F.conv2d(d, wt)


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, k: int, stride: int, padding: int) -> torch.Tensor:
    """dL/dw[co,ci,ky,kx] = sum_{n,oy,ox} dy[n,co,oy,ox] * xpad[n,ci,oy*stride+ky, ox*stride+kx]  (a GEMM with the
    reduction over n*oh*ow: M = co, N = ci*k*k)."""
    n, ci = x.shape[:2]
    co = dy.shape[1]
    cols = F.unfold(x, k, padding=padding, stride=stride)      # (n, ci*k*k, oh*ow)
    return torch.einsum('nol,nkl->ok', dy.reshape(n, co, -1), cols).reshape(co, ci, k, k)


def batchnorm_train_backward(x: torch.Tensor, dy: torch.Tensor, gamma: Optional[torch.Tensor], eps: float):
    """Train-mode BatchNorm2d: with mu, var the (biased) batch statistics over N = B*H*W, xhat = (x - mu) * rsqrt(var + eps):
        dbeta = sum(dy), dgamma = sum(dy * xhat),
        dx = gamma * rsqrt(var + eps) / N * (N * dy - dbeta - xhat * dgamma).
    Two passes over HBM on the GPU: one reduction (dbeta, dgamma per channel), one elementwise."""
    dims = (0, 2, 3)
    cnt = x.numel() // x.shape[1]
    var, mean = torch.var_mean(x, dim=dims, unbiased=False, keepdim=True)
    inv = (var + eps).rsqrt()
    xhat = (x - mean) * inv
    dbeta = dy.sum(dims)
    dgamma = (dy * xhat).sum(dims)
    g = gamma.view(1, -1, 1, 1) if gamma is not None else 1.0
    dx = g * inv / cnt * (cnt * dy - dbeta.view(1, -1, 1, 1) - xhat * dgamma.view(1, -1, 1, 1))
    return dx, dgamma, dbeta


def maxpool_backward(x: torch.Tensor, dy: torch.Tensor, s: int) -> torch.Tensor:
    """MaxPool2d(s, stride=s): the gradient goes to the FIRST maximum of each s x s window in (ky, kx) scan order (ATen's
    forward keeps an element only when it is strictly greater than the running maximum).  Ties are common: the input is
    post-ReLU."""
    n, c, h, w = x.shape
    win = x.view(n, c, h // s, s, w // s, s).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // s, w // s, s * s)
    arg = win.argmax(dim=-1, keepdim=True)                     # first occurrence
    dwin = torch.zeros_like(win).scatter_(-1, arg, dy.unsqueeze(-1))
    return dwin.view(n, c, h // s, w // s, s, s).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h, w)


def upsample2_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor):
    """Depthwise ConvTranspose2d(k=4, s=2, p=1, groups=C):  y[n,c,2i-1+ky,2j-1+kx] += x[n,c,i,j] * w[c,0,ky,kx].
        dx[n,c,i,j]  = sum_{ky,kx} dy[n,c,2i-1+ky,2j-1+kx] * w[c,0,ky,kx]
        dw[c,0,ky,kx] = sum_{n,i,j} x[n,c,i,j] * dy[n,c,2i-1+ky,2j-1+kx]        (the weight is a trained parameter,
    dla_neck.py:58-65 + fill_up_weights)."""
    n, c, h, wd = x.shape
    d = F.pad(dy, (1, 1, 1, 1))                                # index (2i-1+ky)+1 = 2i+ky
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for ky in range(4):
        for kx in range(4):
            tap = d[:, :, ky:ky + 2 * h:2, kx:kx + 2 * wd:2]
            dx += tap * w[:, 0, ky, kx].view(1, -1, 1, 1)
            dw[:, 0, ky, kx] = (x * tap).sum((0, 2, 3))
    return dx, dw


def attn_batchnorm_forward(x, wa, g10, b10, weight_, bias_):
    """Train-mode AttnBatchNorm2d forward with everything the backward needs (attentive_norm.py:79-91, 154-164)."""
    b, c, h, w = x.shape
    var, mean = torch.var_mean(x, dim=(0, 2, 3), unbiased=False, keepdim=True)
    xhat = (x - mean) * (var + 1e-3).rsqrt()
    iv, im = torch.var_mean(x, dim=(2, 3))                      # unbiased instance statistics  (:84)
    r = (iv + 1e-3).rsqrt()
    y = im * r                                                  # (:85)
    a0 = y @ wa.t()                                             # 1x1 conv C -> K on a 1x1 map
    v10, m10 = torch.var_mean(a0, dim=0, unbiased=False, keepdim=True)
    a0hat = (a0 - m10) * (v10 + 1e-5).rsqrt()
    a1 = a0hat * g10 + b10
    a = F.relu6(a1 + 3.) / 6.                                   # hsigmoid (:20)
    wt, bs = a @ weight_, a @ bias_
    out = wt[:, :, None, None] * xhat + bs[:, :, None, None]
    return out, dict(xhat=xhat, var=var, im=im, iv=iv, r=r, y=y, a0hat=a0hat, v10=v10, a1=a1, a=a, wt=wt)


def attn_batchnorm_backward(x, dout, wa, g10, weight_, bias_, s):
    """Backward of the above.  One HBM reduction pass gives, per (image, channel), sum(dout) and sum(dout * xhat); the K = 10
    mixture algebra is then a few hundred flops per image; one elementwise pass writes dx, which has three parts:
    the affine-free batch norm, the instance mean and the instance variance."""
    b, c, h, w = x.shape
    hw, cnt = h * w, b * h * w
    xhat, a = s['xhat'], s['a']
    dbs = dout.sum((2, 3))                                      # (B, C)
    dwt = (dout * xhat).sum((2, 3))
    grads = {'weight_': a.t() @ dwt, 'bias_': a.t() @ dbs}
    da = dwt @ weight_.t() + dbs @ bias_.t()                    # (B, K)
    da1 = da * ((s['a1'] > -3.) & (s['a1'] < 3.)).float() / 6.  # hardtanh backward is strict at both ends
    dg10 = (da1 * s['a0hat']).sum(0)
    db10 = da1.sum(0)
    da0 = g10 * (s['v10'] + 1e-5).rsqrt() / b * (b * da1 - db10 - s['a0hat'] * dg10)
    grads.update({'attn.1.weight': dg10, 'attn.1.bias': db10, 'attn.0.weight': (da0.t() @ s['y']).view(-1, c, 1, 1)})
    dy = da0 @ wa                                               # (B, C)
    dm = dy * s['r']
    dv = dy * s['im'] * (-0.5) * s['r'] ** 3
    dxhat = dout * s['wt'][:, :, None, None]
    sum1 = dxhat.sum((0, 2, 3), keepdim=True)
    sum2 = (dxhat * xhat).sum((0, 2, 3), keepdim=True)
    dx = (s['var'] + 1e-3).rsqrt() / cnt * (cnt * dxhat - sum1 - xhat * sum2)
    dx = dx + dm[:, :, None, None] / hw + dv[:, :, None, None] * 2. * (x - s['im'][:, :, None, None]) / (hw - 1)
    return dx, grads


def sigmoid_clamp_backward(p: torch.Tensor, dp: torch.Tensor) -> torch.Tensor:
    """p = clamp(sigmoid(z), 1e-4, 1 - 1e-4) (monocon_heads.py:168-170): dz = dp * p * (1 - p) inside the clamp, 0 where it
    clamped.  Only p is kept by the forward, so "clamped" is read as p sitting exactly on a bound; ATen's mask is on the
    unclamped sigmoid and would pass a value that equals the bound without being clamped -- a measure-zero difference."""
    inside = (p > 1e-4) & (p < 1. - 1e-4)
    return dp * p * (1. - p) * inside.float()


def depth_transform_backward(z0: torch.Tensor, dd0: torch.Tensor) -> torch.Tensor:
    """d = 1 / (sigmoid(z) + EPS) - 1 (monocon_heads.py:183): dz = -dd * s * (1 - s) / (s + EPS)^2."""
    sg = torch.sigmoid(z0)
    return -dd0 * sg * (1. - sg) / (sg + EPS) ** 2


# --------------------------------------------------------------------------------------
# tape
# --------------------------------------------------------------------------------------
class Tape:
    """Forward ops append (output, backward closure); ``backward`` replays them in reverse, accumulating activation
    gradients per tensor and parameter gradients per state_dict key.  This is the order a stage list is walked on the GPU."""

    def __init__(self, sd: Dict[str, torch.Tensor]):
        self.sd = sd
        self.ops: List[Tuple[torch.Tensor, Callable]] = []
        self.act: Dict[int, torch.Tensor] = {}
        self.param: Dict[str, torch.Tensor] = {}
        self.kernels: List[str] = []                           # backward kernel launches in execution order

    def acc(self, t: torch.Tensor, g: torch.Tensor):
        k = id(t)
        self.act[k] = g if k not in self.act else self.act[k] + g

    def accp(self, key: str, g: torch.Tensor):
        self.param[key] = g if key not in self.param else self.param[key] + g

    def backward(self, seeds: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
        for t, g in seeds:
            self.acc(t, g)
        for out, fn in reversed(self.ops):
            g = self.act.pop(id(out), None)
            if g is not None:
                fn(g)

    # ---- layers -----------------------------------------------------------------------
    def conv(self, x, key, stride=1, padding=0, bias=False, need_dx=True):
        w = self.sd[key + '.weight']
        y = F.conv2d(x, w, self.sd[key + '.bias'] if bias else None, stride=stride, padding=padding)

        def bw(dy):
            self.kernels.append(f'wgrad {key}')
            self.accp(key + '.weight', conv2d_wgrad(x, dy, w.shape[-1], stride, padding))
            if bias:
                self.accp(key + '.bias', dy.sum((0, 2, 3)))
            if need_dx:
                self.kernels.append(f'dgrad {key}')
                self.acc(x, conv2d_dgrad(dy, w, x.shape[2:], stride, padding))
        self.ops.append((y, bw))
        return y

    def bn(self, x, prefix, eps=1e-5):
        gamma, beta = self.sd[prefix + '.weight'], self.sd[prefix + '.bias']
        y = F.batch_norm(x, None, None, gamma, beta, True, 0.1, eps)

        def bw(dy):
            self.kernels.append(f'bn_backward {prefix}')
            dx, dg, db = batchnorm_train_backward(x, dy, gamma, eps)
            self.accp(prefix + '.weight', dg)
            self.accp(prefix + '.bias', db)
            self.acc(x, dx)
        self.ops.append((y, bw))
        return y

    def relu(self, x):
        y = x.clamp_min(0.)
        self.ops.append((y, lambda dy: self.acc(x, dy * (y > 0).float())))
        return y

    def add(self, a, b):
        y = a + b

        def bw(dy):
            self.acc(a, dy)
            self.acc(b, dy)
        self.ops.append((y, bw))
        return y

    def cat(self, xs):
        y = torch.cat(list(xs), 1)

        def bw(dy):
            o = 0
            for t in xs:
                self.acc(t, dy[:, o:o + t.shape[1]])
                o += t.shape[1]
        self.ops.append((y, bw))
        return y

    def maxpool(self, x, s):
        y = F.max_pool2d(x, s, stride=s)

        def bw(dy):
            self.kernels.append('maxpool_backward')
            self.acc(x, maxpool_backward(x, dy, s))
        self.ops.append((y, bw))
        return y

    def up(self, x, key):
        w = self.sd[key + '.weight']
        assert w.shape[-1] == 4
        y = F.conv_transpose2d(x, w, None, stride=2, padding=1, groups=w.shape[0])

        def bw(dy):
            self.kernels.append(f'upsample2_backward {key}')
            dx, dw = upsample2_backward(x, w, dy)
            self.accp(key + '.weight', dw)
            self.acc(x, dx)
        self.ops.append((y, bw))
        return y

    def attn_bn(self, x, prefix):
        sd = self.sd
        wa = sd[prefix + '.attn_weights.attention.0.weight'].flatten(1)
        g10, b10 = sd[prefix + '.attn_weights.attention.1.weight'], sd[prefix + '.attn_weights.attention.1.bias']
        y, saved = attn_batchnorm_forward(x, wa, g10, b10, sd[prefix + '.weight_'], sd[prefix + '.bias_'])

        def bw(dy):
            self.kernels.append(f'attn_bn_backward {prefix}')
            dx, g = attn_batchnorm_backward(x, dy, wa, g10, sd[prefix + '.weight_'], sd[prefix + '.bias_'], saved)
            self.accp(prefix + '.weight_', g['weight_'])
            self.accp(prefix + '.bias_', g['bias_'])
            self.accp(prefix + '.attn_weights.attention.0.weight', g['attn.0.weight'])
            self.accp(prefix + '.attn_weights.attention.1.weight', g['attn.1.weight'])
            self.accp(prefix + '.attn_weights.attention.1.bias', g['attn.1.bias'])
            self.acc(x, dx)
        self.ops.append((y, bw))
        return y


# --------------------------------------------------------------------------------------
# the detector on the tape (same graph as monocon_oracle.dla34_forward / dlaup_forward / heads_forward)
# --------------------------------------------------------------------------------------
def _block(t: Tape, x, prefix, stride, residual=None):
    residual = x if residual is None else residual
    out = t.relu(t.bn(t.conv(x, prefix + '.conv1', stride, 1), prefix + '.bn1'))
    out = t.bn(t.conv(out, prefix + '.conv2', 1, 1), prefix + '.bn2')
    return t.relu(t.add(out, residual))


def _tree(t: Tape, x, prefix, levels, cin, cout, stride, level_root, children=None):
    children = [] if children is None else children
    bottom = t.maxpool(x, stride) if stride > 1 else x
    if levels == 1:          # an outer Tree's projected residual is never consumed (dla.py:187-205): no gradient, not computed
        residual = t.bn(t.conv(bottom, prefix + '.project.0'), prefix + '.project.1') if cin != cout else bottom
    if level_root:
        children.append(bottom)
    if levels == 1:
        x1 = _block(t, x, prefix + '.tree1', stride, residual)
        x2 = _block(t, x1, prefix + '.tree2', 1)
        return t.relu(t.bn(t.conv(t.cat([x2, x1, *children]), prefix + '.root.conv'), prefix + '.root.bn'))
    x1 = _tree(t, x, prefix + '.tree1', levels - 1, cin, cout, stride, False)
    children.append(x1)
    return _tree(t, x1, prefix + '.tree2', levels - 1, cout, cout, 1, False, children=children)


def forward_on_tape(t: Tape, img: torch.Tensor):
    """Returns (pred, seeds): the reference's pred_dict and, for each entry, the pre-transform tensor the backward starts at."""
    ch = DLA34_CHANNELS
    x = t.relu(t.bn(t.conv(img, 'backbone.base_layer.0', 1, 3, need_dx=False), 'backbone.base_layer.1'))
    x = t.relu(t.bn(t.conv(x, 'backbone.level0.0', 1, 1), 'backbone.level0.1'))
    maps = [x]
    x = t.relu(t.bn(t.conv(x, 'backbone.level1.0', 2, 1), 'backbone.level1.1'))
    maps.append(x)
    for lvl in range(2, 6):
        x = _tree(t, x, f'backbone.level{lvl}', DLA34_LEVELS[lvl], ch[lvl - 1], ch[lvl], 2, lvl != 2)
        maps.append(x)
    layers = list(maps[2:])
    for i in range(len(layers) - 1):                           # DLAUp.forward, dla_neck.py:136-143
        sub = layers[-i - 2:]
        p = f'neck.ida_{i}'
        for j in range(1, len(sub)):                           # IDAUp.forward, dla_neck.py:94-106
            u = t.relu(t.bn(t.conv(sub[j], f'{p}.proj_{j}.conv', 1, 1), f'{p}.proj_{j}.bn1'))
            u = t.up(u, f'{p}.up_{j}')
            sub[j] = t.relu(t.bn(t.conv(t.cat([sub[j - 1], u]), f'{p}.node_{j}.conv', 1, 1), f'{p}.node_{j}.bn1'))
        layers[-i - 2:] = sub
    feat = layers[-1]
    stems = {n: t.relu(t.attn_bn(t.conv(feat, f'head.{n}.0', 1, 1, bias=True), f'head.{n}.1')) for n in HEAD_STEMS}
    raw = {k: t.conv(stems[s], 'head.' + c, bias=True) for k, s, c in PRED_KEYS}
    pred = dict(raw)
    for k in ('center_heatmap_pred', 'kpt_heatmap_pred'):
        pred[k] = torch.clamp(torch.sigmoid(raw[k]), 1e-4, 1. - 1e-4)
    d = raw['depth_pred']
    pred['depth_pred'] = torch.cat([(1. / (torch.sigmoid(d[:, 0:1]) + EPS)) - 1., d[:, 1:2]], 1)
    return pred, raw


def pred_grad_to_raw(pred, raw, dpred):
    """Chain dL/dpred (what ``mc_losses`` returns) through the output transforms of monocon_heads.py:168-170,183."""
    draw = dict(dpred)
    for k in ('center_heatmap_pred', 'kpt_heatmap_pred'):
        draw[k] = sigmoid_clamp_backward(pred[k], dpred[k])
    dd = dpred['depth_pred']
    draw['depth_pred'] = torch.cat([depth_transform_backward(raw['depth_pred'][:, 0:1], dd[:, 0:1]), dd[:, 1:2]], 1)
    return draw


def manual_train_step(sd: Dict[str, torch.Tensor], img: torch.Tensor, label: Dict[str, np.ndarray], pad_hw):
    """The same step as ``monocon_oracle.train_step`` with the network's backward done by the formulas above.  The only
    autograd use is d(total)/d(pred) through ``train_oracle.losses`` -- on the GPU that is ``mc_losses`` (already parity-tested
    by tests/test_gpu_train_ops.py)."""
    from . import train_oracle as TO
    with torch.no_grad():
        t = Tape({k: v.clone() for k, v in sd.items()})
        pred, raw = forward_on_tape(t, img.float())
    leaves = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
    fh, fw = pred['center_heatmap_pred'].shape[2:]
    tgt = TO.generate_targets(label, pad_hw, (fh, fw))
    losses = TO.losses(leaves, {k: torch.from_numpy(v) for k, v in tgt.items()})
    total = sum(losses.values())
    total.backward()
    dpred = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    with torch.no_grad():
        draw = pred_grad_to_raw(pred, raw, dpred)
        t.backward([(raw[k], draw[k]) for k in raw])
    return {'losses': {k: float(v.detach()) for k, v in losses.items()}, 'total': float(total.detach()), 'grads': t.param,
            'kernels': t.kernels, 'dpred': dpred}
