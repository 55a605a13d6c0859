"""GPU: the train-mode forward (mc_forward_train: batch-statistic BatchNorm everywhere, AttnBN in train mode, running-stat
updates) against the training-step oracle that is pinned to the reference's own step (tests/test_train_step_oracle.py), then
the ten losses through the GPU target generator + loss kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import engine as E              # noqa: E402
from monocon_pytorch_b200 import train_ops as T           # noqa: E402
from oracle import fixtures as FX                          # noqa: E402
from oracle import monocon_oracle as O                     # noqa: E402
from oracle import train_fixtures as TF                    # noqa: E402

DEV = torch.device('cuda', 0)


def test_train_mode_forward_losses_and_running_stats(fixture_sd):
    B, H, W = 2, 128, 256
    img = FX.make_images(B, H, W, seed=31)
    label = TF.make_labels(B, (H, W), seed=32)
    ref = O.train_step(fixture_sd, img, label, (H, W))
    eng = E.Engine(DEV, B, H, W, 'fp32')
    eng.load_state_dict(fixture_sd, training=True)
    pred = eng.forward_train(img.to(DEV))
    torch.cuda.synchronize()
    worst = 0.0
    for k, t in zip(E.PRED_NAMES, pred):
        r = ref['pred'][k].numpy()
        err = float(np.abs(t.cpu().numpy() - r).max() / max(1e-12, np.abs(r).max()))
        worst = max(worst, err)
        assert err < 1e-3, (k, err)                        # the tolerance north_star states for floating point
    # running statistics after one step (momentum 0.1; AttnBN base BN 0.03)
    checked = 0
    for key, val in ref['buffers'].items():
        if key.endswith('num_batches_tracked') or key.startswith(('backbone.level3.project.', 'backbone.level4.project.')):
            continue                                       # the two unused outer `project` BatchNorms are not in the plan
        got = eng.get_buffer(key, val.numel()).numpy()
        np.testing.assert_allclose(got, val.numpy(), rtol=1e-4, atol=1e-5, err_msg=key)
        checked += 1
    assert checked >= 130                                   # 2 buffers x (49 BN of the plan + 9 AttnBN base + 9 attention BN)
    # the ten losses from the engine's train-mode maps through the GPU target generator and loss kernels
    data = {'img': img.to(DEV), 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(DEV) for k, v in label.items()}}
    tgt = T.TargetGenerator()(data, (B, 64, H // 4, W // 4))
    loss = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt)
    for k, v in ref['losses'].items():
        assert abs(float(loss[k]) - v) <= 2e-3 * max(1.0, abs(v)), (k, float(loss[k]), v)
    # a second step keeps moving the running statistics (the engine owns them between steps)
    before = eng.get_buffer('backbone.level2.tree1.bn1.running_mean', 64).clone()
    eng.forward_train(FX.make_images(B, H, W, seed=33).to(DEV))
    assert not torch.equal(before, eng.get_buffer('backbone.level2.tree1.bn1.running_mean', 64))
    with pytest.raises(E.EngineError):
        eng.forward_train(img[:1].to(DEV))                 # the 10-channel BatchNorm needs B >= 2, like the reference
    eng.close()


def test_module_train_mode_returns_pred_and_losses(fixture_sd):
    """Drop-in surface: ``model.train(); pred_dict, loss_dict = model(data_dict)`` (engine/monocon_engine.py:84) -- forward
    only: same losses as the reference's step, the module's running statistics and num_batches_tracked updated like torch
    does, and loss.backward() raising in the default (forward-only) mode; the opt-in backward is tests/test_gpu_zz_train_backward.py."""
    import monocon_pytorch_b200 as M
    B, H, W = 2, 128, 256
    img = FX.make_images(B, H, W, seed=31)
    label = TF.make_labels(B, (H, W), seed=32)
    ref = O.train_step(fixture_sd, img, label, (H, W))
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    model.load_state_dict(fixture_sd, strict=True)
    model = model.to(DEV).train()
    data = {'img': img.to(DEV), 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(DEV) for k, v in label.items()}}
    pred, loss = model(data)
    assert tuple(pred) == E.PRED_NAMES and tuple(loss) == T.LOSS_NAMES
    for k, v in ref['losses'].items():
        assert abs(float(loss[k]) - v) <= 2e-3 * max(1.0, abs(v)), (k, float(loss[k]), v)
    sd = model.state_dict()
    for key in ('backbone.level2.tree1.bn1.running_var', 'neck.ida_0.node_1.bn1.running_mean', 'head.depth_head.1.running_var',
                'head.dir_feat.1.attn_weights.attention.1.running_mean'):
        np.testing.assert_allclose(sd[key].cpu().numpy(), ref['buffers'][key].numpy(), rtol=1e-4, atol=1e-5, err_msg=key)
    k = 'backbone.level2.tree1.bn1.num_batches_tracked'
    assert int(sd[k]) == int(fixture_sd[k]) + 1
    # the two outer `project` BatchNorms the reference executes for nothing (dla.py:194): their statistics move too
    for key in ('backbone.level3.project.1.running_mean', 'backbone.level3.project.1.running_var',
                'backbone.level4.project.1.running_mean', 'backbone.level4.project.1.running_var'):
        assert not torch.equal(sd[key].cpu(), fixture_sd[key])
        np.testing.assert_allclose(sd[key].cpu().numpy(), ref['buffers'][key].numpy(), rtol=1e-4, atol=1e-5, err_msg=key)
    with pytest.raises(RuntimeError):
        sum(loss.values()).backward()
    only_pred = model(data, return_loss=False)
    assert tuple(only_pred) == E.PRED_NAMES
