"""CPU: the training-side oracle (oracle/train_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/gen_train_golden.py): targets (utils/target_generator.py), the ten losses and their gradients
(model/dense_heads/monocon_heads.py:203-310 under autograd) and the clip + AdamW step."""
import os

import numpy as np
import pytest
import torch

from oracle import train_fixtures as TF
from oracle import train_oracle as TO

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'train_small.npz')
INT_KEYS = ('indices', 'indices_kpt', 'mask_target', 'alpha_cls_target', 'mask_center2kpt_offset', 'mask_kpt_heatmap_offset')


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


def _label(gold):
    return {k[len('label/'):]: v for k, v in gold.items() if k.startswith('label/')}


def test_fixture_labels_are_the_ones_the_golden_was_made_from(gold):
    lab = TF.make_labels(3, (128, 256), seed=3, empty_images=(1,))
    for k, v in lab.items():
        assert np.array_equal(v, gold['label/' + k]), k


def test_targets_match_reference(gold):
    tgt = TO.generate_targets(_label(gold), (128, 256), (32, 64))
    assert set(tgt) == {k[len('target/'):] for k in gold if k.startswith('target/')}
    for k, v in tgt.items():
        ref = gold['target/' + k]
        assert v.shape == ref.shape, k
        if k in INT_KEYS:
            assert np.array_equal(v.astype(np.float64), ref.astype(np.float64)), k          # bit-exact integer outputs
        else:
            np.testing.assert_allclose(v, ref, rtol=0, atol=1e-6, err_msg=k)
    # the Gaussian centres are exact ones (num_pos of the focal loss counts them, losses/focal_loss.py:24)
    assert (tgt['center_heatmap_target'] == 1).sum() == (gold['target/center_heatmap_target'] == 1).sum() > 0


def test_losses_and_gradients_match_reference(gold):
    pred = {k[len('pred/'):]: torch.from_numpy(v).requires_grad_(True) for k, v in gold.items() if k.startswith('pred/')}
    tgt = {k[len('target/'):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith('target/')}
    loss = TO.losses(pred, tgt)
    assert set(loss) == set(TO.LOSS_NAMES)
    for k in TO.LOSS_NAMES:
        assert abs(float(loss[k].detach()) - float(gold['loss/' + k])) <= 1e-6 * max(1.0, abs(float(gold['loss/' + k]))), k
    sum(loss.values()).backward()
    for k, p in pred.items():
        ref = gold['grad/' + k]
        np.testing.assert_allclose(p.grad.numpy(), ref, rtol=1e-5, atol=1e-7 * max(1.0, np.abs(ref).max()), err_msg=k)


def test_empty_batch_asserts_like_the_reference(gold):
    pred = {k[len('pred/'):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith('pred/')}
    tgt = {k[len('target/'):]: torch.from_numpy(v.copy()) for k, v in gold.items() if k.startswith('target/')}
    tgt['mask_target'][:] = False
    with pytest.raises(AssertionError):
        TO.losses(pred, tgt)


def test_clip_adamw_matches_torch(gold):
    ps, gs = TF.make_opt_tensors(seed=7)
    m = [np.zeros_like(p) for p in ps]
    v = [np.zeros_like(p) for p in ps]
    for step, (lr, b1) in enumerate(TF.OPT_SCHEDULE):
        tn = TO.clip_adamw_step(ps, [g.copy() for g in gs[step]], m, v, step + 1, lr, b1, 0.99)
        assert abs(float(tn) - float(gold[f'opt/norm{step}'])) <= 1e-5 * float(gold[f'opt/norm{step}'])
        for i, p in enumerate(ps):
            np.testing.assert_allclose(p, gold[f'opt/p{step}_{i}'], rtol=2e-6, atol=1e-8, err_msg=f'step {step} tensor {i}')
    assert float(gold['opt/norm1']) > 35.0 > float(gold['opt/norm0'])      # the fixture exercises both clip branches
