"""CPU oracle for the MonoCon forward + decode hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, state_dict-driven restatement of the reference's
inference path (2gunsu/monocon-pytorch).  It is the *checker* for the CUDA
engine: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``monocon_pytorch_b200/`` imports it, and the product path has no CPU fallback.

Parity pinning: the reference publishes no tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against the reference's own PyTorch
modules run in the build container: ``tests/golden/gen_golden.py`` imports
``/root/reference``, loads the seeded fixture state_dict (``oracle/fixtures.py``)
into the reference ``MonoConDetector`` and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
those files on every run (CPU, no reference needed).

Arithmetic: the reference's convolutions / batch-norms are PyTorch ATen ops on
fp32 with TF32 off (reference test.py:30-33); the same ATen CPU ops are used
here (torch is the third-party dependency that holds that arithmetic, SURVEY.md
§8c).  The decode (NMS / top-k / gather / lifting) is restated in numpy with an
explicit tie-break (lowest flat index first), which is what the CUDA kernel
implements; on tie-free inputs it equals ``torch.topk``.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PI = float(np.pi)          # model/dense_heads/monocon_heads.py:26
EPS = 1e-12                # model/dense_heads/monocon_heads.py:25

# DLA-34 arch settings: model/backbone/dla.py:211
DLA34_LEVELS = (1, 1, 1, 2, 2, 1)
DLA34_CHANNELS = (16, 32, 64, 128, 256, 512)

# registration order of the nine 3x3 stems and their 1x1 outputs
# (model/dense_heads/monocon_heads.py:74-88)
HEAD_STEMS = ('heatmap_head', 'wh_head', 'offset_head', 'center2kpt_offset_head',
              'kpt_heatmap_head', 'kpt_heatmap_offset_head', 'dim_head', 'depth_head', 'dir_feat')

# pred_dict key -> (stem, key of the 1x1 conv)   (monocon_heads.py:165-200)
PRED_KEYS = (
    ('center_heatmap_pred', 'heatmap_head', 'heatmap_head.3'),
    ('kpt_heatmap_pred', 'kpt_heatmap_head', 'kpt_heatmap_head.3'),
    ('wh_pred', 'wh_head', 'wh_head.3'),
    ('offset_pred', 'offset_head', 'offset_head.3'),
    ('kpt_heatmap_offset_pred', 'kpt_heatmap_offset_head', 'kpt_heatmap_offset_head.3'),
    ('center2kpt_offset_pred', 'center2kpt_offset_head', 'center2kpt_offset_head.3'),
    ('dim_pred', 'dim_head', 'dim_head.3'),
    ('depth_pred', 'depth_head', 'depth_head.3'),
    ('alpha_cls_pred', 'dir_feat', 'dir_cls.0'),
    ('alpha_offset_pred', 'dir_feat', 'dir_reg.0'),
)
PRED_NAMES = tuple(k for k, _, _ in PRED_KEYS)


class _Ctx:
    """Carries the state_dict and the optional BN-calibration switch."""

    def __init__(self, sd: Dict[str, torch.Tensor], calibrate: bool = False, emulate_bf16: bool = False, train: bool = False):
        self.sd = sd
        self.calibrate = calibrate
        self.emulate_bf16 = emulate_bf16
        self.train = train          # nn.Module.train(): batch statistics + running-stat update in every BatchNorm
        self.trace = None           # dict: the neck blocks record their outputs by prefix (stage-wise comparisons)

    def q(self, x: torch.Tensor) -> torch.Tensor:
        """bf16 storage emulation: where the CUDA engine's throughput mode stores an activation or a
        convolution weight as bf16, round it (round-to-nearest-even); identity in the fp32 oracle."""
        return x.bfloat16().float() if self.emulate_bf16 else x


# --------------------------------------------------------------------------------------
# primitive layers
# --------------------------------------------------------------------------------------
def _bn(ctx: _Ctx, x: torch.Tensor, prefix: str, eps: float = 1e-5, affine: bool = True) -> torch.Tensor:
    """Eval-mode nn.BatchNorm2d (running statistics).  dla.py:24,30,119,185,233,295; dla_neck.py:27.

    ``calibrate`` (fixture construction only, never used for parity): overwrite the
    running statistics with this batch's statistics first, so that a seeded random
    network gets realistic, non-identity BN folds.
    """
    sd = ctx.sd
    if ctx.calibrate and 'attn_weights' not in prefix:
        dims = (0, 2, 3)
        var, mean = torch.var_mean(x, dim=dims, unbiased=True) if x.numel() // x.shape[1] > 1 else \
            (torch.ones(x.shape[1]), x.mean(dim=dims))
        sd[prefix + '.running_mean'] = mean.detach().clone()
        sd[prefix + '.running_var'] = var.detach().clone().clamp_min(1e-4)
    w = sd[prefix + '.weight'] if affine else None
    b = sd[prefix + '.bias'] if affine else None
    if ctx.train:
        # train mode: batch statistics, running statistics updated in place (momentum 0.1; 0.03 for the AttnBN base BN,
        # monocon_heads.py:117), num_batches_tracked + 1 (torch.nn.modules.batchnorm._BatchNorm.forward)
        momentum = 0.03 if ('head.' in prefix and 'attn_weights' not in prefix) else 0.1
        if prefix + '.num_batches_tracked' in sd:
            sd[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
        return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], w, b, True, momentum, eps)
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], w, b, False, 0.0, eps)


def _conv(ctx: _Ctx, x: torch.Tensor, key: str, stride: int = 1, padding: int = 0, bias: bool = False,
          quant_w: bool = True) -> torch.Tensor:
    b = ctx.sd[key + '.bias'] if bias else None
    w = ctx.sd[key + '.weight']
    return F.conv2d(x, ctx.q(w) if quant_w else w, b, stride=stride, padding=padding)


# --------------------------------------------------------------------------------------
# backbone: DLA-34   (model/backbone/dla.py)
# --------------------------------------------------------------------------------------
def _basic_block(ctx: _Ctx, x, prefix: str, stride: int, residual=None):
    """BasicBlock.forward, dla.py:34-51."""
    if residual is None:
        residual = x
    out = _conv(ctx, x, prefix + '.conv1', stride=stride, padding=1)
    out = ctx.q(F.relu(_bn(ctx, out, prefix + '.bn1')))
    out = _conv(ctx, out, prefix + '.conv2', stride=1, padding=1)
    out = _bn(ctx, out, prefix + '.bn2')
    return ctx.q(F.relu(out + residual))


def _root(ctx: _Ctx, xs: Sequence[torch.Tensor], prefix: str):
    """Root.forward with residual=False (DLA-34), dla.py:124-132."""
    x = _conv(ctx, torch.cat(list(xs), 1), prefix + '.conv')
    return ctx.q(F.relu(_bn(ctx, x, prefix + '.bn')))


def _tree(ctx: _Ctx, x, prefix: str, levels: int, cin: int, cout: int, stride: int,
          level_root: bool, residual=None, children: Optional[list] = None):
    """Tree.forward, dla.py:187-205 (including its quirk: an outer Tree's ``residual``
    is computed and handed to ``tree1`` which, being a Tree, ignores it)."""
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride=stride) if stride > 1 else x          # dla.py:193
    if cin != cout:                                                                # dla.py:194
        residual = ctx.q(_bn(ctx, _conv(ctx, bottom, prefix + '.project.0'), prefix + '.project.1'))
    else:
        residual = bottom
    if level_root:
        children.append(bottom)                                                    # dla.py:196-197
    if levels == 1:
        x1 = _basic_block(ctx, x, prefix + '.tree1', stride, residual)             # dla.py:198
        x2 = _basic_block(ctx, x1, prefix + '.tree2', 1)                           # dla.py:200
        return _root(ctx, [x2, x1, *children], prefix + '.root')                   # dla.py:201
    x1 = _tree(ctx, x, prefix + '.tree1', levels - 1, cin, cout, stride, False, residual=residual)
    children.append(x1)                                                            # dla.py:203
    return _tree(ctx, x1, prefix + '.tree2', levels - 1, cout, cout, 1, False, children=children)


def dla34_forward(ctx: _Ctx, img: torch.Tensor) -> List[torch.Tensor]:
    """DLA.forward, dla.py:273-278 -> six maps."""
    ch = DLA34_CHANNELS
    x = _conv(ctx, ctx.q(img), 'backbone.base_layer.0', stride=1, padding=3)       # dla.py:231-234
    x = ctx.q(F.relu(_bn(ctx, x, 'backbone.base_layer.1')))
    outs = []
    x = ctx.q(F.relu(_bn(ctx, _conv(ctx, x, 'backbone.level0.0', 1, 1), 'backbone.level0.1')))   # dla.py:236
    outs.append(x)
    x = ctx.q(F.relu(_bn(ctx, _conv(ctx, x, 'backbone.level1.0', 2, 1), 'backbone.level1.1')))   # dla.py:237
    outs.append(x)
    for lvl in range(2, 6):                                                        # dla.py:238-241
        x = _tree(ctx, x, f'backbone.level{lvl}', DLA34_LEVELS[lvl], ch[lvl - 1], ch[lvl], 2,
                  level_root=(lvl != 2))
        outs.append(x)
    return outs


# --------------------------------------------------------------------------------------
# neck: DLAUp / IDAUp   (model/backbone/dla_neck.py)
# --------------------------------------------------------------------------------------
def _conv_block(ctx: _Ctx, x, prefix: str):
    """Conv2dBlock (3x3, no bias, BN, ReLU), dla_neck.py:34-38.  A state_dict that carries ``<prefix>.conv.conv_offset.*`` selects the
    DCN variant of the block (north_star; absent from the reference, see oracle/dcn_oracle.py): the 3x3 convolution is a DCNv2 pack."""
    if prefix + '.conv.conv_offset.weight' in ctx.sd:
        from . import dcn_oracle as D
        y = D.dcn_pack(x, ctx.sd[prefix + '.conv.weight'], ctx.sd[prefix + '.conv.conv_offset.weight'],
                       ctx.sd[prefix + '.conv.conv_offset.bias'], round_fn=ctx.q if ctx.emulate_bf16 else None)
        out = ctx.q(F.relu(_bn(ctx, y, prefix + '.bn1')))
    else:
        out = ctx.q(F.relu(_bn(ctx, _conv(ctx, x, prefix + '.conv', 1, 1), prefix + '.bn1')))
    if ctx.trace is not None:
        ctx.trace[prefix] = out
    return out


def _ida_up(ctx: _Ctx, layers: List[torch.Tensor], prefix: str) -> List[torch.Tensor]:
    """IDAUp.forward, dla_neck.py:94-106.  ``up_i`` is a depthwise ConvTranspose2d
    (k=2f, s=f, p=f//2, groups=C, no bias), dla_neck.py:58-65; every use in DLAUp has f=2."""
    for i in range(1, len(layers)):
        w = ctx.sd[f'{prefix}.up_{i}.weight']
        f = w.shape[-1] // 2
        x = _conv_block(ctx, layers[i], f'{prefix}.proj_{i}')
        x = ctx.q(F.conv_transpose2d(x, w, None, stride=f, padding=f // 2, groups=w.shape[0]))
        layers[i] = _conv_block(ctx, torch.cat([layers[i - 1], x], 1), f'{prefix}.node_{i}')
    return layers


def dlaup_forward(ctx: _Ctx, maps: Sequence[torch.Tensor]) -> torch.Tensor:
    """DLAUp.forward with start_level=2, dla_neck.py:136-143 (list-slice mutation semantics)."""
    layers = list(maps[2:])
    for i in range(len(layers) - 1):
        layers[-i - 2:] = _ida_up(ctx, layers[-i - 2:], f'neck.ida_{i}')
    return layers[-1]


# --------------------------------------------------------------------------------------
# heads   (model/dense_heads/monocon_heads.py, model/norm/attentive_norm.py)
# --------------------------------------------------------------------------------------
def attn_batchnorm(ctx: _Ctx, x: torch.Tensor, prefix: str) -> torch.Tensor:
    """AttnBatchNorm2d.forward (attentive_norm.py:154-164) + AttnWeights.forward (:79-91).

    base BN: affine-free, eps 1e-3 (monocon_heads.py:117);  instance statistic
    y = mean * rsqrt(unbiased var + 1e-3);  a = hsigmoid(BN10(conv1x1(y)));
    out = (a @ weight_) * BN(x) + (a @ bias_).
    """
    sd = ctx.sd
    out = _bn(ctx, x, prefix, eps=1e-3, affine=False)
    b, c = x.shape[:2]
    var, mean = torch.var_mean(x, dim=(2, 3), keepdim=True)                        # attentive_norm.py:84
    y = mean * (var + 1e-3).rsqrt()                                                # attentive_norm.py:85
    a = F.conv2d(y, sd[prefix + '.attn_weights.attention.0.weight'])               # attentive_norm.py:51 (fp32 on the GPU too)
    a = _bn(ctx, a, prefix + '.attn_weights.attention.1')                          # attentive_norm.py:52
    a = (F.relu6(a + 3.) / 6.).view(b, -1)                                         # attentive_norm.py:20,91
    weight = a @ sd[prefix + '.weight_']                                           # attentive_norm.py:159
    bias = a @ sd[prefix + '.bias_']                                               # attentive_norm.py:160
    return weight[:, :, None, None] * out + bias[:, :, None, None]


def heads_forward(ctx: _Ctx, feat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """MonoConDenseHeads._get_predictions, monocon_heads.py:165-200."""
    stems = {}
    for name in HEAD_STEMS:                                                        # monocon_heads.py:114-131
        x = ctx.q(_conv(ctx, feat, f'head.{name}.0', 1, 1, bias=True))             # stored pre-norm stem output
        # bf16 emulation: the engine's tensor-core head (csrc/head_tc.cu) feeds the 1x1 convs bf16 operands
        stems[name] = ctx.q(F.relu(attn_batchnorm(ctx, x, f'head.{name}.1')))
    pred = {}
    for key, stem, conv in PRED_KEYS:
        pred[key] = _conv(ctx, stems[stem], 'head.' + conv, bias=True, quant_w=True)
    for key in ('center_heatmap_pred', 'kpt_heatmap_pred'):                        # monocon_heads.py:168-170
        pred[key] = torch.clamp(torch.sigmoid(pred[key]), 1e-4, 1. - 1e-4)
    d = pred['depth_pred']                                                         # monocon_heads.py:183
    if ctx.train:                                                                  # autograd-safe form of the in-place write
        pred['depth_pred'] = torch.cat([((1. / (torch.sigmoid(d[:, 0:1]) + EPS)) - 1.), d[:, 1:2]], 1)
    else:
        d[:, 0] = (1. / (torch.sigmoid(d[:, 0]) + EPS)) - 1.
    return pred


def forward(sd: Dict[str, torch.Tensor], img: torch.Tensor, calibrate: bool = False,
            return_intermediates: bool = False, emulate_bf16: bool = False):
    """MonoConDetector.forward in eval mode (monocon_detector.py:53-65,85-87).

    ``emulate_bf16`` rounds exactly the tensors that the engine's throughput mode stores as bf16
    (activations between layers and convolution weights; accumulation, BN folds, AttnBN and the 1x1
    output convolutions stay fp32).  It is the checker for that mode: the reference itself is fp32."""
    ctx = _Ctx(sd, calibrate, emulate_bf16)
    if return_intermediates:
        ctx.trace = {}
    with torch.no_grad():
        maps = dla34_forward(ctx, img.float())
        feat = dlaup_forward(ctx, maps)
        pred = heads_forward(ctx, feat)
    if return_intermediates:
        return pred, {'backbone': maps, 'feat': feat, 'neck': ctx.trace}
    return pred


def train_step(sd: Dict[str, torch.Tensor], img: torch.Tensor, label: Dict[str, np.ndarray], pad_hw) -> Dict[str, object]:
    """One training forward + backward as the reference runs it (engine/monocon_engine.py:80-91 up to ``backward``):
    ``MonoConDetector.forward`` in train mode (batch-statistic BatchNorm everywhere, monocon_detector.py:53-61) ->
    ``TargetGenerator`` -> ``_get_losses`` -> plain sum (utils/engine_utils.py:79-80) -> autograd.

    Returns the ten losses, the total, d(total)/d(parameter) for every floating-point parameter that received a gradient, and
    the updated BatchNorm buffers.  `sd` is not modified.  Checker for the future GPU training step (configs[2]); pinned to
    the reference by tests/golden/train_step.npz."""
    from . import train_oracle as TO
    work = {k: v.clone() for k, v in sd.items()}
    params = [k for k, v in work.items() if v.is_floating_point() and not k.endswith(('running_mean', 'running_var'))]
    for k in params:
        work[k].requires_grad_(True)
    ctx = _Ctx(work, train=True)
    maps = dla34_forward(ctx, img.float())
    feat = dlaup_forward(ctx, maps)
    pred = heads_forward(ctx, feat)
    fh, fw = feat.shape[2:]
    tgt = TO.generate_targets(label, pad_hw, (fh, fw))
    losses = TO.losses(pred, {k: torch.from_numpy(v) for k, v in tgt.items()})
    total = sum(losses.values())
    total.backward()
    grads = {k: work[k].grad.detach().clone() for k in params if work[k].grad is not None}
    buffers = {k: v.detach().clone() for k, v in work.items() if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))}
    return {'losses': {k: float(v.detach()) for k, v in losses.items()}, 'total': float(total.detach()), 'grads': grads,
            'buffers': buffers, 'pred': {k: v.detach() for k, v in pred.items()}}


# --------------------------------------------------------------------------------------
# decode   (utils/tensor_ops.py, model/dense_heads/monocon_heads.py:313-558) -- numpy
# --------------------------------------------------------------------------------------
def local_maximum(heat: np.ndarray, kernel: int = 3) -> np.ndarray:
    """get_local_maximum, tensor_ops.py:17-21: keep = (maxpool3x3(h) == h); h * keep."""
    pad = (kernel - 1) // 2
    B, C, H, W = heat.shape
    padded = np.full((B, C, H + 2 * pad, W + 2 * pad), -np.inf, dtype=heat.dtype)
    padded[:, :, pad:pad + H, pad:pad + W] = heat
    hmax = heat.copy()
    for dy in range(kernel):
        for dx in range(kernel):
            np.maximum(hmax, padded[:, :, dy:dy + H, dx:dx + W], out=hmax)
    return heat * (hmax == heat).astype(heat.dtype)


def topk_from_heatmap(scores: np.ndarray, k: int):
    """get_topk_from_heatmap, tensor_ops.py:24-31.  Sorted descending; equal scores are
    ordered by ascending flat index (torch.topk leaves that order unspecified)."""
    B, C, H, W = scores.shape
    flat = scores.reshape(B, -1)
    order = np.argsort(-flat, axis=1, kind='stable')[:, :k]
    topk_scores = np.take_along_axis(flat, order, axis=1)
    topk_inds = order.astype(np.int64)
    clses = topk_inds // (H * W)
    inds = topk_inds % (H * W)
    ys = inds // W
    xs = (inds % W).astype(np.int32).astype(np.float32)
    return topk_scores, inds, clses, ys, xs


def gather_rows(feat: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """transpose_and_gather_feat, tensor_ops.py:55-59: (B,C,H,W),(B,K) -> (B,K,C)."""
    B, C, H, W = feat.shape
    f = feat.reshape(B, C, H * W)
    return np.stack([f[b][:, ind[b]].T for b in range(B)], axis=0)


def decode_alpha(alpha_cls: np.ndarray, alpha_offset: np.ndarray, num_bins: int = 12) -> np.ndarray:
    """decode_alpha, monocon_heads.py:379-396 (first max on ties, one wrap pass)."""
    cls = np.argmax(alpha_cls, axis=-1)[..., None]
    off = np.take_along_axis(alpha_offset, cls, axis=2)
    angle_per_class = np.float32((2 * PI) / float(num_bins))
    alpha = (cls.astype(np.float32) * angle_per_class + off).astype(np.float32)
    alpha = np.where(alpha > PI, alpha - np.float32(2 * PI), alpha)
    alpha = np.where(alpha < -PI, alpha + np.float32(2 * PI), alpha)
    return alpha.astype(np.float32)


def inverse_viewpad(P2: np.ndarray) -> np.ndarray:
    """4x4 inverse of the padded projection, computed on the CPU in fp32 exactly like
    convert_pts2D_to_pts3D does (monocon_heads.py:543-546).  Returns inv(viewpad) (not transposed)."""
    out = []
    for p in np.asarray(P2, dtype=np.float32).reshape(-1, 3, 4):
        viewpad = torch.eye(4)
        viewpad[:3, :4] = torch.from_numpy(p.copy())
        out.append(torch.inverse(viewpad).numpy())
    return np.stack(out, 0).astype(np.float32)


def decode(pred: Dict[str, np.ndarray], P2: np.ndarray, img_hw: Tuple[int, int], topk: int = 30,
           thres: float = 0.4, kernel: int = 3, num_bins: int = 12, invP: Optional[np.ndarray] = None):
    """decode_heatmap (monocon_heads.py:399-482) + the origin shift of _get_bboxes (:313-329),
    with fixed-shape outputs: box2d (B,K,5), box3d (B,K,7), labels (B,K) int64,
    inds (B,K) int64, valid (B,K) bool.  ``P2``: (B,3,4) float32."""
    f32 = np.float32
    p = {k: np.asarray(v, dtype=f32) for k, v in pred.items()}
    img_h, img_w = img_hw
    heat = p['center_heatmap_pred']
    B, _, fh, fw = heat.shape
    P2 = np.asarray(P2, dtype=f32).reshape(B, 3, 4)
    if invP is None:
        invP = inverse_viewpad(P2)
    heat = local_maximum(heat, kernel)
    scores, inds, labels, ys, xs = topk_from_heatmap(heat, topk)
    ysf = ys.astype(f32)

    wh = gather_rows(p['wh_pred'], inds)
    offset = gather_rows(p['offset_pred'], inds)
    tx = xs + offset[..., 0]
    ty = ysf + offset[..., 1]
    sx, sy = f32(img_w / fw), f32(img_h / fh)
    x1 = (tx - wh[..., 0] / f32(2.)) * sx
    y1 = (ty - wh[..., 1] / f32(2.)) * sy
    x2 = (tx + wh[..., 0] / f32(2.)) * sx
    y2 = (ty + wh[..., 1] / f32(2.)) * sy

    alpha = decode_alpha(gather_rows(p['alpha_cls_pred'], inds),
                         gather_rows(p['alpha_offset_pred'], inds), num_bins)      # (B,K,1)
    depth_pred = gather_rows(p['depth_pred'], inds)
    sigma = np.exp(-depth_pred[..., 1]).astype(f32)                                # monocon_heads.py:440
    score = (scores * sigma).astype(f32)
    box2d = np.stack([x1, y1, x2, y2, score], axis=2).astype(f32)

    c2k = gather_rows(p['center2kpt_offset_pred'], inds)[..., -2:]                 # monocon_heads.py:443-446
    cu = ((c2k[..., 0] + xs) * sx).astype(f32)
    cv = ((c2k[..., 1] + ysf) * sy).astype(f32)

    # calculate_roty, monocon_heads.py:485-515
    fx = P2[:, 0, 0][:, None]
    cx = P2[:, 0, 2][:, None]
    rot_y = (alpha[..., 0] + np.arctan2((cu - cx).astype(f32), np.broadcast_to(fx, cu.shape).astype(f32))).astype(f32)
    while (rot_y > PI).any():
        rot_y = np.where(rot_y > PI, rot_y - f32(2 * PI), rot_y).astype(f32)
    while (rot_y < -PI).any():
        rot_y = np.where(rot_y < -PI, rot_y + f32(2 * PI), rot_y).astype(f32)

    # convert_pts2D_to_pts3D, monocon_heads.py:518-558
    d = depth_pred[..., 0]
    homo = np.stack([cu * d, cv * d, d, np.ones_like(d)], axis=-1).astype(f32)     # (B,K,4)
    xyz = np.einsum('bkj,bij->bki', homo, invP.astype(f32)).astype(f32)[..., :3]   # homo @ inv^T

    dim = gather_rows(p['dim_pred'], inds)
    box3d = np.concatenate([xyz, dim, rot_y[..., None]], axis=-1).astype(f32)
    valid = box2d[..., 4] > f32(thres)                                             # monocon_heads.py:465
    box3d[..., 1] += box3d[..., 4] * f32(0.5)                                      # monocon_heads.py:320-328
    return {'box2d': box2d, 'box3d': box3d, 'labels': labels.astype(np.int64),
            'inds': inds.astype(np.int64), 'scores_raw': scores.astype(f32), 'valid': valid}


def to_ragged(dec: Dict[str, np.ndarray]):
    """Fixed-shape decode -> the reference's per-image ragged lists (monocon_heads.py:467-480)."""
    out2d, out3d, outl = [], [], []
    for b in range(dec['valid'].shape[0]):
        m = dec['valid'][b]
        out2d.append(dec['box2d'][b][m])
        out3d.append(dec['box3d'][b][m])
        outl.append(dec['labels'][b][m])
    return out2d, out3d, outl


def forward_and_decode(sd, img: torch.Tensor, P2: np.ndarray, topk: int = 30, thres: float = 0.4):
    pred = forward(sd, img)
    pred_np = {k: v.numpy() for k, v in pred.items()}
    dec = decode(pred_np, P2, tuple(img.shape[-2:]), topk=topk, thres=thres)
    return pred_np, dec


# --------------------------------------------------------------------------------------------
# input pipeline: Normalize -> Pad(32) -> ToTensor of the reference's test transforms
# (transforms/default_transforms.py:376-431 with dataset/monocon_dataset.py:38-42).  Pinned by tests/golden/input.npz.
# --------------------------------------------------------------------------------------------
INPUT_MEAN = (123.675, 116.28, 103.53)
INPUT_STD = (58.395, 57.12, 57.375)


def preprocess_u8(frames, size_divisor: int = 32, mean=INPUT_MEAN, std=INPUT_STD) -> np.ndarray:
    """frames: list of (h, w, 3) uint8 arrays (sizes may differ) -> (B, 3, H, W) float32 with every frame normalised in
    float64 like numpy does ((img.astype(float32) - mean) / std, then torch.Tensor -> float32), zero-padded bottom / right to
    the largest per-frame multiple of `size_divisor`."""
    m = np.array(mean).reshape(1, 1, -1)
    s = np.array(std).reshape(1, 1, -1)
    outs = []
    for f in frames:
        x = ((f.astype(np.float32) - m) / s)
        h, w = x.shape[:2]
        ph = int(np.ceil(h / size_divisor)) * size_divisor
        pw = int(np.ceil(w / size_divisor)) * size_divisor
        canvas = np.zeros((ph, pw, 3), dtype=x.dtype)
        canvas[:h, :w] = x
        outs.append(canvas.astype(np.float32).transpose(2, 0, 1))
    H = max(o.shape[1] for o in outs)
    W = max(o.shape[2] for o in outs)
    batch = np.zeros((len(outs), 3, H, W), np.float32)
    for i, o in enumerate(outs):
        batch[i, :, :o.shape[1], :o.shape[2]] = o
    return batch
