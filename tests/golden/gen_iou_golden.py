"""Generate tests/golden/iou.npz from the UNMODIFIED reference evaluation kernels (run in the build container only).

    NUMBA_ENABLE_CUDASIM=1 PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_iou_golden.py

rotate_iou_gpu_eval (engine/kitti_eval/rotate_iou.py:337-379, a numba.cuda kernel: executed by numba's CUDA simulator
here, there is no GPU in the build container) and d3_box_overlap (engine/kitti_eval/eval.py:159-164) on seeded boxes:
random rotated boxes plus identical, contained, disjoint, axis-aligned and quarter-turn pairs."""
import os
import sys

import numpy as np

assert os.environ.get('NUMBA_ENABLE_CUDASIM') == '1', 'run with NUMBA_ENABLE_CUDASIM=1'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference/engine')
from kitti_eval.rotate_iou import rotate_iou_gpu_eval          # noqa: E402  (the reference)
from kitti_eval.eval import d3_box_overlap                      # noqa: E402  (the reference)


def boxes_bev(seed, n):
    rng = np.random.RandomState(seed)
    b = np.stack([rng.uniform(-8, 8, n), rng.uniform(2, 18, n), rng.uniform(0.6, 5, n), rng.uniform(0.6, 5, n),
                  rng.uniform(-3.2, 3.2, n)], 1).astype(np.float32)
    return b


def main():
    b = boxes_bev(1, 24)
    q = boxes_bev(2, 18)
    q[0] = b[0]                                   # identical
    q[1] = b[1]; q[1, 2:4] *= 0.5                 # contained
    q[2] = b[2]; q[2, 0] += 100                   # disjoint
    b[3, 4] = 0; q[3] = b[3]; q[3, 0] += 0.5      # axis-aligned, shifted
    q[4] = b[4]; q[4, 4] += np.float32(np.pi / 2)  # quarter turn about the same centre
    out = {'boxes': b, 'qboxes': q}
    for c in (-1, 0, 1, 2):
        out[f'riou{c}'] = rotate_iou_gpu_eval(b, q, c)
    rng = np.random.RandomState(3)
    def cam(bev, n):
        return np.stack([bev[:, 0], rng.uniform(1.0, 2.2, n), bev[:, 1], bev[:, 2], rng.uniform(1.2, 2.0, n), bev[:, 3], bev[:, 4]], 1)
    b3, q3 = cam(b, len(b)).astype(np.float64), cam(q, len(q)).astype(np.float64)
    q3[0] = b3[0]
    out['boxes3d'], out['qboxes3d'] = b3, q3
    for c in (-1, 0, 1):
        out[f'd3_{c}'] = d3_box_overlap(b3, q3, c)
    np.savez_compressed(os.path.join(HERE, 'iou.npz'), **out)
    print('wrote iou.npz', {k: v.shape for k, v in out.items()}, float(out['riou-1'][0, 0]), float(out['d3_-1'][0, 0]))


if __name__ == '__main__':
    main()
