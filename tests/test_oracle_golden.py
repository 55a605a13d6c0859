"""CPU: pin the oracle (oracle/monocon_oracle.py) against outputs of the UNMODIFIED reference.

tests/golden/*.npz were produced by tests/golden/gen_golden.py, which runs the reference's own
MonoConDetector.forward and MonoConDenseHeads._get_bboxes on the seeded fixture.  Tolerances: the
reference is not bit-stable against itself across thread counts / batch composition (SURVEY.md §8c:
~2e-6 relative), so maps are compared at 2e-5 relative-to-max; indices / labels are exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import monocon_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_fixture_table_matches_reference_keys():
    ref = [l.strip().split(' ', 1) for l in open(os.path.join(GOLDEN, 'keys.txt'))]
    mine = [(k, f'{tuple(s)} {d}') for k, s, d in FX.param_table()]
    assert len(mine) == len(ref) == 449
    for (k, s), (rk, rs) in zip(mine, ref):
        assert k == rk and s == rs


def test_fixture_weights_digest(fixture_sd):
    """The seeded generator reproduces the tensors the goldens were made with (container == GPU box)."""
    d = np.load(os.path.join(GOLDEN, 'weights_digest.npz'))
    for k, s in zip(d['keys'], d['sums']):
        got = float(fixture_sd[str(k)].double().sum())
        assert abs(got - s) <= 1e-4 * max(1.0, abs(s)), k


def _rel_to_max(a, b):
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def test_oracle_forward_small(fixture_sd, golden_small):
    g = golden_small
    h, w = [int(v) for v in g['hw']]
    img = FX.make_images(2, h, w, seed=int(g['img_seed']))
    pred = O.forward(fixture_sd, img)
    for k in O.PRED_NAMES:
        assert pred[k].shape == g['pred/' + k].shape
        assert _rel_to_max(pred[k].numpy(), g['pred/' + k]) < 2e-5, k


@pytest.mark.parametrize('thres', [0.4, 1.0])
def test_oracle_decode_small(golden_small, thres):
    """Decode restatement on the reference's own prediction maps: indices / labels exact."""
    g = golden_small
    h, w = [int(v) for v in g['hw']]
    pred = {k: g['pred/' + k] for k in O.PRED_NAMES}
    dec = O.decode(pred, g['P2'], (h, w), topk=30, thres=thres)
    assert np.array_equal(dec['inds'], g['topk/inds'][:, :30])
    assert np.array_equal(dec['labels'], g['topk/clses'][:, :30])
    assert np.array_equal(dec['scores_raw'], g['topk/scores'][:, :30])
    b2, b3, lb = O.to_ragged(dec)
    for b in range(2):
        assert np.array_equal(lb[b], g[f'dec{thres}/labels/{b}'])
        np.testing.assert_allclose(b2[b], g[f'dec{thres}/box2d/{b}'], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(b3[b], g[f'dec{thres}/box3d/{b}'], rtol=1e-5, atol=2e-5)


def test_oracle_full_size(fixture_sd, golden_full):
    """384x1280 (the BASELINE.json geometry): sampled map values, moments, top-k and boxes."""
    g = golden_full
    h, w = [int(v) for v in g['hw']]
    img = FX.make_images(2, h, w, seed=int(g['img_seed']))
    pred_np, dec = O.forward_and_decode(fixture_sd, img, g['P2'], topk=30, thres=0.4)
    for k in O.PRED_NAMES:
        a = pred_np[k]
        vals = a.reshape(-1)[g['pos/' + k]]
        scale = max(1e-12, float(g['mom/' + k][2]))
        assert np.abs(vals - g['val/' + k]).max() / scale < 2e-5, k
        assert abs(a.mean(dtype=np.float64) - g['mom/' + k][0]) < 1e-5 * scale
    # top-k: exact wherever the reference's score gap to the next candidate is not a near-tie
    ref_scores, ref_inds = g['topk/scores'], g['topk/inds']
    gaps = ref_scores[:, :-1] - ref_scores[:, 1:]
    safe = np.minimum(np.concatenate([np.full((2, 1), 1.0), gaps[:, :-1]], 1), gaps) > 5e-6
    assert safe.mean() > 0.8
    assert np.array_equal(dec['inds'][safe], ref_inds[:, :30][safe])
    b2, b3, lb = O.to_ragged(dec)
    if safe.all():
        for b in range(2):
            assert np.array_equal(lb[b], g[f'dec0.4/labels/{b}'])
            np.testing.assert_allclose(b2[b], g[f'dec0.4/box2d/{b}'], rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(b3[b], g[f'dec0.4/box3d/{b}'], rtol=1e-4, atol=2e-4)


def test_oracle_topk_tie_break_and_padding():
    """Edge cases the decode must define itself (torch.topk leaves ties unspecified): plateaus and
    maps with fewer local maxima than k."""
    heat = np.full((1, 3, 8, 8), 0.5, dtype=np.float32)             # one big plateau: every cell is a local max
    s, inds, cls, ys, xs = O.topk_from_heatmap(O.local_maximum(heat), 30)
    assert np.array_equal(cls[0], np.zeros(30)) and np.array_equal(inds[0], np.arange(30))
    heat = np.full((1, 3, 8, 8), 1e-4, dtype=np.float32)
    heat[0, 2, 3, 4] = 0.9
    heat[0, 1, 0, 0] = 0.8
    nms = O.local_maximum(heat)
    s, inds, cls, ys, xs = O.topk_from_heatmap(nms, 5)
    assert list(cls[0][:2]) == [2, 1] and list(inds[0][:2]) == [3 * 8 + 4, 0]
    assert s[0][0] == np.float32(0.9) and s[0][1] == np.float32(0.8)


def test_input_pipeline_oracle_matches_reference_transforms():
    """Normalize + Pad(32) + ToTensor restated in oracle.preprocess_u8 vs the reference's own transform classes
    (tests/golden/gen_input_golden.py): bit-identical float32 tensors, frames of different sizes."""
    g = dict(np.load(os.path.join(GOLDEN, 'input.npz')))
    frames = [g[f'frame{i}'] for i in range(3)]
    batch = O.preprocess_u8(frames)
    assert batch.shape == (3, 3, 32, 96)
    for i in range(3):
        ref = g[f'tensor{i}']
        assert tuple(g[f'pad_shape{i}']) == ref.shape[1:]
        assert np.array_equal(batch[i, :, :ref.shape[1], :ref.shape[2]], ref)
        assert not batch[i, :, ref.shape[1]:].any() and not batch[i, :, :, ref.shape[2]:].any()
