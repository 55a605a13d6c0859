"""Seeded synthetic labels / prediction maps / optimiser tensors for the training-side parity tests.
TEST INFRASTRUCTURE ONLY (see train_oracle.py).  Label layout = MonoConDataset._create_empty_labels
(dataset/monocon_dataset.py:160-171) after collation: leading batch dimension, max_objs = 30 rows per image."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

MAX_OBJS = 30
NUM_KPT = 9
PRED_CH = {'center_heatmap_pred': 3, 'kpt_heatmap_pred': 9, 'wh_pred': 2, 'offset_pred': 2, 'kpt_heatmap_offset_pred': 2,
           'center2kpt_offset_pred': 18, 'dim_pred': 3, 'depth_pred': 2, 'alpha_cls_pred': 12, 'alpha_offset_pred': 12}

# (lr, beta1) per optimiser step, as CyclicScheduler (solver/cyclic_scheduler.py:36-71) rewrites them every iteration
OPT_SCHEDULE = ((2.25e-4, 0.95), (3.1e-4, 0.943), (4.4e-4, 0.931))
OPT_SHAPES = ((64, 32, 3, 3), (64,), (10, 64), (128, 64, 1, 1), (3,))


def make_labels(B: int, pad_hw: Tuple[int, int], seed: int, empty_images: Sequence[int] = (), min_objs: int = 3,
                max_objs_per_image: int = 8) -> Dict[str, np.ndarray]:
    """Random KITTI-like labels.  Valid rows are NOT contiguous (the reference compacts them with the mask,
    utils/target_generator.py:48-52); some key-points fall outside the image / feature map and carry visibility 0 / 1 / 2."""
    rng = np.random.RandomState(seed)
    H, W = pad_hw
    M, K = MAX_OBJS, NUM_KPT
    lab = {'gt_bboxes': np.zeros((B, M, 4), np.float32), 'gt_labels': np.zeros((B, M), np.uint8),
           'gt_bboxes_3d': np.zeros((B, M, 7), np.float32), 'depths': np.zeros((B, M), np.float32),
           'gt_kpts_2d': np.zeros((B, M, 2 * K), np.float32), 'gt_kpts_valid_mask': np.zeros((B, M, K), np.uint8),
           'mask': np.zeros((B, M), bool)}
    for b in range(B):
        if b in empty_images:
            continue
        n = rng.randint(min_objs, max_objs_per_image + 1)
        rows = np.sort(rng.choice(M, n, replace=False))
        for i, r in enumerate(rows):
            bw = rng.uniform(0.04, 0.35) * W
            bh = rng.uniform(0.06, 0.45) * H
            cx = rng.uniform(bw / 2 + 1, W - bw / 2 - 1)
            cy = rng.uniform(bh / 2 + 1, H - bh / 2 - 1)
            if i == 1:                         # second object shares the first one's centre cell (overlapping splats)
                cx, cy = lab['gt_bboxes'][b, rows[0], [0, 1]] + lab['gt_bboxes'][b, rows[0], [2, 3]]
                cx, cy = cx / 2 + 0.3, cy / 2 + 0.2
                bw, bh = min(bw, 2 * min(cx, W - cx) - 2), min(bh, 2 * min(cy, H - cy) - 2)
            lab['gt_bboxes'][b, r] = (cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2)
            lab['gt_labels'][b, r] = rng.randint(0, 3)
            lab['gt_bboxes_3d'][b, r] = np.concatenate([rng.uniform(-20, 20, 3), rng.uniform(0.5, 4.5, 3), rng.uniform(-7, 7, 1)])
            lab['depths'][b, r] = rng.uniform(3, 60)
            k = np.stack([cx + rng.uniform(-0.8, 0.8, K) * bw, cy + rng.uniform(-0.8, 0.8, K) * bh], 1)
            k[rng.randint(0, K)] += (W, 0)       # one key-point far outside the image
            k[rng.randint(0, K)] = (-3.0, cy)    # one slightly left of it (negative coordinate: int() truncates toward 0)
            lab['gt_kpts_2d'][b, r] = k.reshape(-1)
            vis = rng.choice([0, 1, 2], K, p=[0.2, 0.1, 0.7]).astype(np.uint8)
            lab['gt_kpts_valid_mask'][b, r] = vis
            lab['mask'][b, r] = True
    return lab


def make_pred(B: int, feat_hw: Tuple[int, int], seed: int) -> Dict[str, np.ndarray]:
    """Prediction maps in the value ranges the heads produce (heat-maps clamped sigmoids, positive dims)."""
    rng = np.random.RandomState(seed)
    h, w = feat_hw
    out = {}
    for k, c in PRED_CH.items():
        x = rng.randn(B, c, h, w).astype(np.float32)
        if k.endswith('heatmap_pred'):
            x = np.clip(1.0 / (1.0 + np.exp(-(x - 2.0))), 1e-4, 1 - 1e-4).astype(np.float32)
        elif k == 'dim_pred':
            x = (1.6 + 0.3 * x).astype(np.float32)
        elif k == 'depth_pred':
            x[:, 0] = (20.0 + 8.0 * x[:, 0])
            x[:, 1] = 0.5 * x[:, 1]
        out[k] = np.ascontiguousarray(x)
    return out


def make_opt_tensors(seed: int):
    rng = np.random.RandomState(seed)
    params = [rng.randn(*s).astype(np.float32) * 0.1 for s in OPT_SHAPES]
    grads = [[(rng.randn(*s) * (40.0 if step == 1 else 0.1)).astype(np.float32) for s in OPT_SHAPES]
             for step in range(len(OPT_SCHEDULE))]      # step 1 exceeds max_norm = 35: the clip is active there
    return params, grads
