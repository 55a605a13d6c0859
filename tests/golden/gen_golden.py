"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py

Imports /root/reference (read-only, never copied), loads the seeded fixture state_dict
(oracle/fixtures.py) into the reference ``MonoConDetector`` with ``strict=True`` (which also
proves key/shape compatibility of the fixture table), runs the reference's own
``forward`` (monocon_detector.py:53-65) and ``_get_bboxes`` (monocon_heads.py:313-329) on the CPU
and stores what the parity tests need:

* small.npz : B=2, 128x256  -- all ten prediction maps in full + decode outputs
* full.npz  : B=2, 384x1280 -- decode outputs, the top-k tuple, per-map moments and 1024 sampled
                               values per map (the full maps would be 16 MB)
* keys.txt  : the reference's state_dict keys / shapes / dtypes
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import fixtures as FX                      # noqa: E402
from model import MonoConDetector                      # noqa: E402  (the reference)
from utils.tensor_ops import get_local_maximum, get_topk_from_heatmap   # noqa: E402  (the reference)

SEED = 0
THRES = (0.4, 1.0)


class _Calib:                                          # the decode reads only .P2 (monocon_heads.py:501,543)
    def __init__(self, p2):
        self.P2 = p2


def sample_positions(n_elem: int, n: int, seed: int) -> np.ndarray:
    return np.random.RandomState(seed).randint(0, n_elem, size=n).astype(np.int64)


def run_case(model, name: str, batch: int, h: int, w: int, img_seed: int, full_maps: bool):
    torch.set_num_threads(os.cpu_count())
    img = FX.make_images(batch, h, w, seed=img_seed)
    P2 = FX.kitti_p2(batch, seed=img_seed)
    data = {'img': img, 'img_metas': {'pad_shape': [(h, w)] * batch}, 'calib': [_Calib(p) for p in P2]}
    with torch.no_grad():
        pred = model(data)
    out = {'P2': P2, 'hw': np.array([h, w]), 'img_seed': np.array(img_seed)}
    for k, v in pred.items():
        a = v.numpy().astype(np.float32)
        if full_maps:
            out['pred/' + k] = a
        else:
            pos = sample_positions(a.size, 1024, 1234)
            out['pos/' + k] = pos
            out['val/' + k] = a.reshape(-1)[pos]
            out['mom/' + k] = np.array([a.mean(dtype=np.float64), a.std(dtype=np.float64), np.abs(a).max()], dtype=np.float64)
    with torch.no_grad():
        nms = get_local_maximum(pred['center_heatmap_pred'].clone(), kernel=3)
        scores, inds, clses, ys, xs = get_topk_from_heatmap(nms, k=31)
    out['topk/scores'] = scores.numpy()                # 31 so that the gap to the first loser is known
    out['topk/inds'] = inds.numpy()
    out['topk/clses'] = clses.numpy()
    for t in THRES:
        model.head.test_thres = t
        with torch.no_grad():
            b2, b3, lb = model.head._get_bboxes(data, {k: v.clone() for k, v in pred.items()})
        for b in range(batch):
            out[f'dec{t}/box2d/{b}'] = b2[b].numpy()
            out[f'dec{t}/box3d/{b}'] = b3[b].numpy()
            out[f'dec{t}/labels/{b}'] = lb[b].numpy()
    model.head.test_thres = 0.4
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith('dec0.4/box2d')})


def main():
    sd = FX.make_state_dict(SEED)
    model = MonoConDetector(34, pretrained_backbone=False).eval()
    print(model.load_state_dict(sd, strict=True))
    with open(os.path.join(HERE, 'keys.txt'), 'w') as f:
        for k, v in model.state_dict().items():
            f.write(f'{k} {tuple(v.shape)} {v.dtype}\n')
    # a digest of the fixture weights, so that a drifting generator is caught on the GPU box
    digest = {k: float(v.double().sum()) for k, v in sd.items() if v.dtype == torch.float32}
    np.savez_compressed(os.path.join(HERE, 'weights_digest.npz'),
                        keys=np.array(list(digest.keys())), sums=np.array(list(digest.values()), dtype=np.float64))
    run_case(model, 'small', 2, 128, 256, img_seed=1, full_maps=True)
    run_case(model, 'full', 2, 384, 1280, img_seed=2, full_maps=False)
    gen_kitti_golden()



def gen_kitti_golden():
    """KITTI-format conversion of the reference (utils/kitti_convert_utils.py) on the decoded boxes stored in full.npz."""
    from utils.kitti_convert_utils import convert_to_kitti_2d, convert_to_kitti_3d, CLASSES
    g = np.load(os.path.join(HERE, 'full.npz'))

    class _FullCalib:
        def __init__(self, p2):
            self.P2 = p2
            self.P0 = p2.copy()
            self.V2C = np.eye(4, dtype=np.float32)[:3]

    res3d, res2d = [], []
    for b in range(2):
        b2 = torch.from_numpy(g[f'dec0.4/box2d/{b}'])
        b3 = torch.from_numpy(g[f'dec0.4/box3d/{b}'])
        lb = torch.from_numpy(g[f'dec0.4/labels/{b}'])
        res3d.append(dict(boxes_3d=b3, scores_3d=b2[:, -1], labels_3d=lb))
        res2d.append([b2.numpy()[lb.numpy() == c] for c in range(3)])
    metas = {'sample_idx': [7, 11], 'ori_shape': [(375, 1242), (370, 1224)], 'scale_hw': [(1.0, 1.0)]}
    calibs = [_FullCalib(p) for p in g['P2']]
    k3 = convert_to_kitti_3d(res3d, metas, calibs)
    k2 = convert_to_kitti_2d(res2d, metas)
    out = {}
    for tag, ks in (('k3', k3), ('k2', k2)):
        for b, anno in enumerate(ks):
            for key, val in anno.items():
                if key == 'name':
                    val = np.array([CLASSES.index(n) for n in val], dtype=np.int64)
                out[f'{tag}/{b}/{key}'] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, 'kitti.npz'), **out)
    print('kitti', {k: v.shape for k, v in out.items() if k.endswith('bbox')})



if __name__ == '__main__':
    main()
