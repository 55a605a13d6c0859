"""Generate tests/golden/input.npz from the UNMODIFIED reference transforms (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_input_golden.py

Normalize + Pad(32) + ToTensor (transforms/default_transforms.py:376-431) as composed for testing in
dataset/monocon_dataset.py:38-42, on three seeded uint8 frames of different KITTI-like sizes (scaled down)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from transforms.default_transforms import Normalize, Pad, ToTensor          # noqa: E402  (the reference)

SIZES = ((29, 90), (30, 94), (32, 96))


def frames(seed=11):
    rng = np.random.RandomState(seed)
    return [rng.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in SIZES]


def main():
    out = {}
    for i, f in enumerate(frames()):
        d = {'img': f, 'img_metas': {}}
        for t in (Normalize(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375]), Pad(size_divisor=32), ToTensor()):
            d = t(d)
        out[f'frame{i}'] = f
        out[f'tensor{i}'] = d['img'].numpy()
        out[f'pad_shape{i}'] = np.array(d['img_metas']['pad_shape'])
    np.savez_compressed(os.path.join(HERE, 'input.npz'), **out)
    print('wrote input.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
