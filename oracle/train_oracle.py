"""CPU restatement of the reference's training-side hot-path pieces.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by anything under monocon_pytorch_b200/.

* ``generate_targets``  -- ``TargetGenerator.__call__`` (utils/target_generator.py:30-138) with
  ``gaussian_radius`` / ``gaussian2D`` / ``generate_gaussian_target`` (utils/tensor_ops.py:62-125),
  numpy, float32 arithmetic in the reference's operation order.
* ``losses``            -- ``MonoConDenseHeads._get_losses`` (model/dense_heads/monocon_heads.py:203-310)
  with losses/{focal_loss,l1_loss,dim_loss,depth_loss,cross_entropy_loss}.py, written with torch CPU
  ops so that autograd yields the gradients w.r.t. the ten prediction maps (what ``loss.backward()``
  hands to the head convolutions, engine/monocon_engine.py:85-91).
* ``clip_adamw_step``   -- ``clip_grad_norm_(35, 2)`` + ``AdamW.step`` as called by
  engine/monocon_engine.py:94-100 (third-party torch.optim arithmetic, restated in numpy float32).

Pinned: tests/golden/train_small.npz is produced by tests/golden/gen_train_golden.py from the UNMODIFIED
reference modules (TargetGenerator, MonoConDenseHeads._get_losses with autograd, torch.optim.AdamW);
tests/test_train_oracle.py holds this file to it (integer outputs exact, floats <= 1e-6).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

F32 = np.float32
EPS = 1e-12
PI = np.pi


# --------------------------------------------------------------------------------------------
# targets
# --------------------------------------------------------------------------------------------
def gaussian_radius(h: F32, w: F32) -> F32:
    """utils/tensor_ops.py:76-98 with min_overlap = 0.3: float32 tensor arithmetic, math.sqrt in double."""
    mo = 0.3
    b1 = F32(h + w)
    c1 = F32(F32(F32(w * h) * F32(1 - mo)) / F32(1 + mo))
    sq1 = math.sqrt(float(F32(F32(b1 * b1) - F32(F32(4) * c1))))
    r1 = F32(F32(b1 - F32(sq1)) / F32(2))
    b2 = F32(F32(2) * F32(h + w))
    c2 = F32(F32(F32(1 - mo) * w) * h)
    sq2 = math.sqrt(float(F32(F32(b2 * b2) - F32(F32(16) * c2))))
    r2 = F32(F32(b2 - F32(sq2)) / F32(8))
    a3 = 4 * mo
    b3 = F32(F32(-2 * mo) * F32(h + w))
    c3 = F32(F32(F32(mo - 1) * w) * h)
    sq3 = math.sqrt(float(F32(F32(b3 * b3) - F32(F32(4 * a3) * c3))))
    r3 = F32(F32(b3 + F32(sq3)) / F32(2 * a3))
    return min(r1, r2, r3)


def splat(canvas: np.ndarray, cx: int, cy: int, radius: int) -> None:
    """generate_gaussian_target (utils/tensor_ops.py:101-125) + gaussian2D (:62-73): max-splat in place."""
    diameter = 2 * radius + 1
    sigma = diameter / 6
    x = np.arange(-radius, radius + 1, dtype=F32)[None, :]
    y = np.arange(-radius, radius + 1, dtype=F32)[:, None]
    g = np.exp((-(x * x + y * y) / F32(2 * sigma * sigma)).astype(F32)).astype(F32)
    g[g < np.finfo(F32).eps * g.max()] = 0
    height, width = canvas.shape
    left, right = min(cx, radius), min(width - cx, radius + 1)
    top, bottom = min(cy, radius), min(height - cy, radius + 1)
    sl = canvas[cy - top:cy + bottom, cx - left:cx + right]
    np.maximum(sl, g[radius - top:radius + bottom, radius - left:radius + right], out=sl)


def angle_to_class(alpha: F32, bins: int = 12):
    """TargetGenerator._convert_angle_to_class (utils/target_generator.py:141-149)."""
    two_pi = F32(2 * PI)
    def rem(a, b):                       # torch.remainder on float32
        m = F32(math.fmod(float(a), float(b)))
        if m != 0 and ((b < 0) != (m < 0)):
            m = F32(m + b)
        return m
    angle = rem(F32(alpha), two_pi)
    apc = 2 * PI / float(bins)
    shifted = rem(F32(angle + F32(apc / 2)), two_pi)
    cls = int(F32(shifted / F32(apc)))
    res = F32(shifted - F32(cls * apc + apc / 2))
    return cls, res


def generate_targets(label: Dict[str, np.ndarray], pad_hw, feat_hw, num_classes=3, max_objs=30, num_kpt=9,
                     bins=12) -> Dict[str, np.ndarray]:
    """label: gt_bboxes (B,M,4) f32, gt_labels (B,M) u8, gt_bboxes_3d (B,M,7) f32, depths (B,M) f32,
    gt_kpts_2d (B,M,18) f32, gt_kpts_valid_mask (B,M,9) u8, mask (B,M) bool."""
    B = label['mask'].shape[0]
    fh, fw = feat_hw
    h_ratio, w_ratio = F32(fh / pad_hw[0]), F32(fw / pad_hw[1])
    M, K = max_objs, num_kpt
    t = {'center_heatmap_target': np.zeros((B, num_classes, fh, fw), F32), 'wh_target': np.zeros((B, M, 2), F32),
         'offset_target': np.zeros((B, M, 2), F32), 'dim_target': np.zeros((B, M, 3), F32),
         'alpha_cls_target': np.zeros((B, M, 1), F32), 'alpha_offset_target': np.zeros((B, M, 1), F32),
         'depth_target': np.zeros((B, M, 1), F32), 'center2kpt_offset_target': np.zeros((B, M, 2 * K), F32),
         'kpt_heatmap_target': np.zeros((B, K, fh, fw), F32), 'kpt_heatmap_offset_target': np.zeros((B, M, 2 * K), F32),
         'indices': np.zeros((B, M), np.int64), 'indices_kpt': np.zeros((B, M, K), np.int64),
         'mask_target': np.zeros((B, M), bool), 'mask_center2kpt_offset': np.zeros((B, M, 2 * K), F32),
         'mask_kpt_heatmap_offset': np.zeros((B, M, 2 * K), F32)}
    for b in range(B):
        m = label['mask'][b].astype(bool)
        boxes = label['gt_bboxes'][b][m].astype(F32)
        if len(boxes) < 1:
            continue
        labels = label['gt_labels'][b][m].astype(np.int64)
        kpts = label['gt_kpts_2d'][b][m].astype(F32).reshape(-1, K, 2).copy()
        kpts[:, :, 0] = kpts[:, :, 0] * w_ratio
        kpts[:, :, 1] = kpts[:, :, 1] * h_ratio
        kmask = label['gt_kpts_valid_mask'][b][m]
        b3d = label['gt_bboxes_3d'][b][m].astype(F32)
        depth = label['depths'][b][m].astype(F32)
        for o in range(len(boxes)):
            ctx = F32(F32(F32(boxes[o, 0] + boxes[o, 2]) * w_ratio) / F32(2))
            cty = F32(F32(F32(boxes[o, 1] + boxes[o, 3]) * h_ratio) / F32(2))
            cxi, cyi = int(ctx), int(cty)
            fbh = F32(F32(boxes[o, 3] - boxes[o, 1]) * h_ratio)
            fbw = F32(F32(boxes[o, 2] - boxes[o, 0]) * w_ratio)
            radius = max(0, int(gaussian_radius(fbh, fbw)))
            splat(t['center_heatmap_target'][b, labels[o]], cxi, cyi, radius)
            t['indices'][b, o] = cyi * fw + cxi
            t['wh_target'][b, o] = (fbw, fbh)
            t['offset_target'][b, o] = (F32(ctx - F32(cxi)), F32(cty - F32(cyi)))
            t['dim_target'][b, o] = b3d[o, 3:6]
            t['depth_target'][b, o] = depth[o]
            cls, res = angle_to_class(b3d[o, 6], bins)
            t['alpha_cls_target'][b, o] = cls
            t['alpha_offset_target'][b, o] = res
            t['mask_target'][b, o] = True
            for k in range(K):
                kx, ky = kpts[o, k]
                kxi, kyi = int(kx), int(ky)
                if kmask[o, k] < 1:
                    continue
                t['center2kpt_offset_target'][b, o, 2 * k] = F32(kx - F32(cxi))
                t['center2kpt_offset_target'][b, o, 2 * k + 1] = F32(ky - F32(cyi))
                t['mask_center2kpt_offset'][b, o, 2 * k:2 * k + 2] = 1
                if not (0 <= kxi < fw and 0 <= kyi < fh):
                    continue
                splat(t['kpt_heatmap_target'][b, k], kxi, kyi, radius)
                t['indices_kpt'][b, o, k] = kyi * fw + kxi
                t['kpt_heatmap_offset_target'][b, o, 2 * k] = F32(kx - F32(kxi))
                t['kpt_heatmap_offset_target'][b, o, 2 * k + 1] = F32(ky - F32(kyi))
                t['mask_kpt_heatmap_offset'][b, o, 2 * k:2 * k + 2] = 1
    t['indices_kpt'] = t['indices_kpt'].reshape(B, -1)
    return t


# --------------------------------------------------------------------------------------------
# losses (torch CPU ops; autograd gives d(sum of the ten losses) / d(prediction maps))
# --------------------------------------------------------------------------------------------
LOSS_NAMES = ['loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_dim', 'loss_center2kpt_offset', 'loss_kpt_heatmap',
              'loss_kpt_heatmap_offset', 'loss_alpha_cls', 'loss_alpha_reg', 'loss_depth']


def _gather(feat: torch.Tensor, ind: torch.Tensor) -> torch.Tensor:
    """transpose_and_gather_feat (utils/tensor_ops.py:34-59): (B,C,H,W), (B,n) -> (B,n,C)."""
    B, C = feat.shape[:2]
    f = feat.permute(0, 2, 3, 1).reshape(B, -1, C)
    return f.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))


def _focal(p: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """GaussianFocalLoss (losses/focal_loss.py:21-44), gamma 2, beta 4, eps 1e-12."""
    pos = t.eq(1).float()
    neg = t.lt(1).float()
    num_pos = pos.sum()
    pos_loss = (torch.log(p + EPS) * torch.pow(1 - p, 2.0) * pos).sum()
    neg_loss = (torch.log((1 - p) + EPS) * torch.pow(p, 2.0) * torch.pow(1 - t, 4.0) * neg).sum()
    return -neg_loss if num_pos == 0 else -(pos_loss + neg_loss) / num_pos


def losses(pred: Dict[str, torch.Tensor], tgt: Dict[str, torch.Tensor], max_objs=30, num_kpts=9, bins=12) -> Dict[str, torch.Tensor]:
    ind, ind_k = tgt['indices'], tgt['indices_kpt']
    B = ind.shape[0]
    mt = tgt['mask_target'].bool()
    n = int(mt.sum())
    assert n > 0, 'the reference asserts on an empty batch (losses/l1_loss.py:15)'
    ex = lambda key: _gather(pred[key], ind)[mt]
    out = {}
    out['loss_offset'] = (ex('offset_pred') - tgt['offset_target'][mt]).abs().mean()
    out['loss_wh'] = 0.1 * (ex('wh_pred') - tgt['wh_target'][mt]).abs().mean()
    dp, dt = ex('dim_pred'), tgt['dim_target'][mt]
    l = (dp - dt).abs() / dp.detach()
    with torch.no_grad():
        comp = (dp - dt).abs().mean() / l.mean()
    out['loss_dim'] = (l * comp).mean()
    dpr = ex('depth_pred')
    d, s = dpr[:, 0], dpr[:, 1]
    out['loss_depth'] = (1.4142 * torch.exp(-s) * (d - tgt['depth_target'][mt].flatten()).abs() + s).mean()
    out['loss_center_heatmap'] = _focal(pred['center_heatmap_pred'], tgt['center_heatmap_target'])
    out['loss_kpt_heatmap'] = _focal(pred['kpt_heatmap_pred'], tgt['kpt_heatmap_target'])
    mk = tgt['mask_center2kpt_offset'][mt]
    out['loss_center2kpt_offset'] = (ex('center2kpt_offset_pred') * mk - tgt['center2kpt_offset_target'][mt]).abs().sum() / (mk.sum() + EPS)
    kp = _gather(pred['kpt_heatmap_offset_pred'], ind_k).reshape(B, max_objs, num_kpts * 2)[mt]
    mko = tgt['mask_kpt_heatmap_offset'][mt]
    out['loss_kpt_heatmap_offset'] = (kp - tgt['kpt_heatmap_offset_target'][mt]).abs().sum() / (mko.sum() + EPS)
    onehot = torch.zeros(n, bins).scatter_(1, tgt['alpha_cls_target'][mt].long().view(-1, 1), 1.0)
    out['loss_alpha_cls'] = torch.nn.functional.binary_cross_entropy_with_logits(ex('alpha_cls_pred'), onehot, reduction='none').mean()
    ao = (ex('alpha_offset_pred') * onehot).sum(1, keepdim=True)
    out['loss_alpha_reg'] = (ao - tgt['alpha_offset_target'][mt]).abs().mean()
    return out


# --------------------------------------------------------------------------------------------
# optimiser step
# --------------------------------------------------------------------------------------------
def clip_adamw_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, beta1: float, beta2: float, eps: float = 1e-8,
                    weight_decay: float = 1e-5, max_norm: float = 35.0):
    """torch.nn.utils.clip_grad_norm_(params, max_norm, 2) followed by torch.optim.AdamW.step (single-tensor path:
    mul_(1 - lr wd), lerp_, mul_/addcmul_, sqrt / sqrt(bc2) + eps, addcdiv_), numpy float32, in place.
    `step` is the 1-based step count after this update.  Returns the total gradient norm (float32)."""
    norms = np.array([np.sqrt(np.sum(g.astype(F32) * g.astype(F32), dtype=F32)) for g in grads], dtype=F32)
    total = F32(np.sqrt(np.sum(norms * norms, dtype=F32)))
    coef = min(F32(1.0), F32(F32(max_norm) / F32(total + F32(1e-6))))
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = (g * coef).astype(F32)
        p *= F32(1 - lr * weight_decay)
        m += F32(1 - beta1) * (g - m)
        v *= F32(beta2)
        v += F32(1 - beta2) * g * g
        denom = (np.sqrt(v) / F32(bc2_sqrt) + F32(eps)).astype(F32)
        p += F32(-step_size) * (m / denom)
    return total
