"""CPU: the full training-step oracle (oracle.monocon_oracle.train_step: train-mode forward, targets, losses, autograd
backward) against digests of the UNMODIFIED reference's own step (tests/golden/gen_train_step_golden.py): the ten losses,
the gradient of every parameter tensor (norm, sum, 8 sampled entries), the tensors that get no gradient, and the updated
BatchNorm buffers.  This pins the checker of the not-yet-built GPU training step (BASELINE.json configs[2])."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import monocon_oracle as O
from oracle import train_fixtures as TF

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'train_step.npz')


def _pos(key, numel):
    return np.random.RandomState(zlib.crc32(key.encode()) & 0x7fffffff).randint(0, max(1, numel), size=8)


@pytest.fixture(scope='module')
def step(fixture_sd):
    torch.set_num_threads(os.cpu_count())
    img = FX.make_images(2, 128, 256, seed=31)
    label = TF.make_labels(2, (128, 256), seed=32)
    return O.train_step(fixture_sd, img, label, (128, 256))


def test_losses_match_reference(step):
    g = np.load(GOLD)
    for k, v in step['losses'].items():
        ref = float(g['loss/' + k])
        assert abs(v - ref) <= 2e-5 * max(1.0, abs(ref)), (k, v, ref)
    assert abs(step['total'] - float(g['total'])) <= 2e-5 * float(g['total'])


def test_parameter_gradients_match_reference(step):
    g = np.load(GOLD)
    nograd = set(g['nograd'].tolist())
    assert nograd == {'backbone.level3.project.0.weight', 'backbone.level3.project.1.weight', 'backbone.level3.project.1.bias',
                      'backbone.level4.project.0.weight', 'backbone.level4.project.1.weight', 'backbone.level4.project.1.bias'}
    keys = [k[len('grad/'):] for k in g.files if k.startswith('grad/')]
    assert len(keys) == 242 - 6 and set(keys) == set(step['grads'])          # every other parameter tensor has a gradient
    worst = 0.0
    for k in keys:
        ref = g['grad/' + k]
        gr = step['grads'][k].double().reshape(-1)
        got = np.concatenate([[float(gr.norm()), float(gr.sum())], gr[_pos(k, gr.numel())].numpy()])
        scale = max(ref[0], 1e-12)                                           # the tensor's own gradient norm
        err = float(np.abs(got - ref).max() / scale)
        worst = max(worst, err)
        assert err <= 2e-4, (k, err, got[:3], ref[:3])
    assert worst > 0 or True


def test_batchnorm_buffers_match_reference(step, fixture_sd):
    g = np.load(GOLD)
    n = 0
    for k in g.files:
        if not k.startswith('buf/'):
            continue
        name = k[len('buf/'):]
        ref = g[k]
        t = step['buffers'][name].double().reshape(-1)
        got = np.concatenate([[float(t.norm()), float(t.sum())], t[_pos(name, t.numel())].numpy()])
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6, err_msg=name)
        n += 1
    assert n > 100
    # the step really moved the running statistics and counted the batch
    k = 'backbone.level2.tree1.bn1.running_mean'
    assert not torch.equal(step['buffers'][k], fixture_sd[k])
    assert int(step['buffers']['backbone.level2.tree1.bn1.num_batches_tracked']) == int(fixture_sd['backbone.level2.tree1.bn1.num_batches_tracked']) + 1
