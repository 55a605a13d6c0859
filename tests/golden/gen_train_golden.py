"""Generate tests/golden/train_small.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_train_golden.py

Training-side rows of the hot-path table:
* targets : the reference ``TargetGenerator`` (utils/target_generator.py:30-138) on seeded synthetic labels
            (B = 3, pad shape 128x256 -> 32x64 maps; one image without objects, key-points inside / outside the map
            and with visibility 0 / 1 / 2, objects of all three classes, overlapping centres)
* losses  : ``MonoConDenseHeads._get_losses`` (model/dense_heads/monocon_heads.py:203-310) on seeded prediction maps with
            autograd: the ten loss values and d(sum)/d(prediction map) for all ten maps
* adamw   : three steps of ``clip_grad_norm_(35, 2)`` + ``torch.optim.AdamW`` (engine/monocon_engine.py:39-53,94-100)
            on five small tensors with per-step lr / beta1 as the cyclic scheduler would set them
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import train_fixtures as TF                              # noqa: E402
from model.dense_heads.monocon_heads import MonoConDenseHeads       # noqa: E402  (the reference)
from utils.target_generator import TargetGenerator                  # noqa: E402  (the reference)


def main():
    torch.set_num_threads(1)
    out = {}
    pad_hw, feat_hw, B = (128, 256), (32, 64), 3
    label = TF.make_labels(B, pad_hw, seed=3, empty_images=(1,))
    tl = {k: torch.from_numpy(v) for k, v in label.items()}
    data = {'img': torch.zeros(B, 3, *pad_hw), 'img_metas': {'pad_shape': [pad_hw] * B}, 'label': tl}
    tgt = TargetGenerator()(data, feat_shape=(B, 64, *feat_hw))
    for k, v in label.items():
        out['label/' + k] = v
    for k, v in tgt.items():
        out['target/' + k] = v.numpy()

    pred = TF.make_pred(B, feat_hw, seed=5)
    pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in pred.items()}
    head = MonoConDenseHeads()
    loss = head._get_losses(pt, tgt)
    total = sum(loss.values())
    total.backward()
    for k, v in pred.items():
        out['pred/' + k] = v
        out['grad/' + k] = pt[k].grad.numpy()
    for k, v in loss.items():
        out['loss/' + k] = np.float32(float(v))

    # optimiser
    ps, gs = TF.make_opt_tensors(seed=7)
    params = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in ps]
    opt = torch.optim.AdamW(params, lr=2.25e-4, weight_decay=1e-5, betas=(0.95, 0.99))
    sched = TF.OPT_SCHEDULE
    for step, (lr, b1) in enumerate(sched):
        for g in opt.param_groups:
            g['lr'] = lr
            g['betas'] = (b1, 0.99)
        for p, g in zip(params, gs[step]):
            p.grad = torch.from_numpy(g.copy())
        tn = torch.nn.utils.clip_grad_norm_(params, max_norm=35, norm_type=2)
        opt.step()
        out[f'opt/norm{step}'] = np.float32(float(tn))
        for i, p in enumerate(params):
            out[f'opt/p{step}_{i}'] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'train_small.npz'), **out)
    print('wrote train_small.npz:', len(out), 'arrays;', {k: float(v) for k, v in loss.items()})


if __name__ == '__main__':
    main()
