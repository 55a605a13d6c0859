"""Generate tests/golden/train_step.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_train_step_golden.py

One training forward + backward of the reference ``MonoConDetector`` in ``train()`` mode (engine/monocon_engine.py:80-91):
seeded fixture weights (oracle/fixtures.py), B = 2 frames of 128x256, seeded labels (oracle/train_fixtures.py).  Stored: the
ten losses, and for every parameter tensor the L2 norm, the sum and 8 sampled entries of its gradient (None-gradient
tensors are listed), plus the same digests of the updated BatchNorm buffers."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import fixtures as FX, train_fixtures as TF          # noqa: E402
from model import MonoConDetector                                # noqa: E402  (the reference)

B, HW = 2, (128, 256)


def sample_positions(key: str, numel: int) -> np.ndarray:
    import zlib
    return np.random.RandomState(zlib.crc32(key.encode()) & 0x7fffffff).randint(0, max(1, numel), size=8)


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    model.load_state_dict(FX.make_state_dict(0), strict=True)
    model.train()
    img = FX.make_images(B, *HW, seed=31)
    label = TF.make_labels(B, HW, seed=32)
    data = {'img': img, 'img_metas': {'pad_shape': [HW] * B}, 'label': {k: torch.from_numpy(v) for k, v in label.items()}}
    pred, loss = model(data)
    total = sum(loss.values())
    total.backward()
    out = {'total': np.float64(float(total))}
    for k, v in loss.items():
        out['loss/' + k] = np.float64(float(v))
    nograd = []
    for k, p in model.named_parameters():
        if p.grad is None:
            nograd.append(k)
            continue
        g = p.grad.detach().double().reshape(-1)
        pos = sample_positions(k, g.numel())
        out['grad/' + k] = np.concatenate([[float(g.norm()), float(g.sum())], g[pos].numpy()])
    for k, b in model.named_buffers():
        f = b.detach().double().reshape(-1)
        pos = sample_positions(k, f.numel())
        out['buf/' + k] = np.concatenate([[float(f.norm()), float(f.sum())], f[pos].numpy()])
    out['nograd'] = np.array(nograd)
    np.savez_compressed(os.path.join(HERE, 'train_step.npz'), **out)
    print('wrote train_step.npz:', len(out), 'entries; total loss', float(total), '; tensors without gradient:', nograd)


if __name__ == '__main__':
    main()
