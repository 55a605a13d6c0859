"""GPU parity of the training-side kernels (csrc/train_ops.cu) through the C ABI:
targets vs the reference's golden vectors and vs the oracle at the bench size, the ten losses and their gradients, and
the fused clip + AdamW step.  Integer outputs bit-exact; float tolerances are written at each assertion."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import train_ops as T           # noqa: E402
from oracle import train_fixtures as TF                    # noqa: E402
from oracle import train_oracle as TO                      # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'train_small.npz')
INT_KEYS = ('indices', 'indices_kpt', 'mask_target', 'alpha_cls_target', 'mask_center2kpt_offset', 'mask_kpt_heatmap_offset')
DEV = torch.device('cuda', 0)


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


def _data_dict(label, B, pad_hw):
    return {'img': torch.zeros(B, 3, 4, 4, device=DEV), 'img_metas': {'pad_shape': [pad_hw] * B},
            'label': {k: torch.from_numpy(v).to(DEV) for k, v in label.items()}}


def _check_targets(out, ref, heat_tol):
    assert set(out) == set(ref)
    for k, v in out.items():
        a, r = v.cpu().numpy(), ref[k]
        assert a.shape == r.shape, k
        if k in INT_KEYS:
            assert np.array_equal(a.astype(np.float64), r.astype(np.float64)), k                  # bit-exact
        elif k.endswith('heatmap_target'):
            np.testing.assert_allclose(a, r, rtol=0, atol=heat_tol, err_msg=k)                     # expf vs the CPU's exp
            assert np.array_equal(a == 1, r == 1), k                                               # centres are exact ones
            assert np.array_equal(a == 0, r == 0), k                                               # same support
        else:
            np.testing.assert_allclose(a, r, rtol=0, atol=1e-6, err_msg=k)


def test_targets_match_reference_golden(gold):
    label = {k[len('label/'):]: v for k, v in gold.items() if k.startswith('label/')}
    out = T.TargetGenerator()(_data_dict(label, 3, (128, 256)), (3, 64, 32, 64))
    _check_targets(out, {k[len('target/'):]: v for k, v in gold.items() if k.startswith('target/')}, 1e-6)
    assert out['mask_target'].dtype == torch.bool and out['indices'].dtype == torch.int64


@pytest.mark.parametrize('B,seed,empty', [(32, 11, (5, 17)), (1, 12, ()), (4, 13, (0, 1, 2, 3))])
def test_targets_match_oracle_at_bench_size(B, seed, empty):
    """BASELINE.json configs[2] geometry: 384x1280 frames -> 96x320 maps, B = 32; also B = 1 and an all-empty batch."""
    pad_hw, feat_hw = (384, 1280), (96, 320)
    label = TF.make_labels(B, pad_hw, seed=seed, empty_images=empty, min_objs=3, max_objs_per_image=30)
    ref = TO.generate_targets(label, pad_hw, feat_hw)
    out = T.TargetGenerator()(_data_dict(label, B, pad_hw), (B, 64, *feat_hw))
    _check_targets(out, ref, 1e-6)
    # idempotence: a second call on recycled memory gives the same tensors (everything is re-zeroed)
    out2 = T.TargetGenerator()(_data_dict(label, B, pad_hw), (B, 64, *feat_hw))
    for k in out:
        assert torch.equal(out[k], out2[k]), k


def _loss_inputs(gold):
    pred = {k[len('pred/'):]: torch.from_numpy(v).to(DEV) for k, v in gold.items() if k.startswith('pred/')}
    tgt = {k[len('target/'):]: torch.from_numpy(v).to(DEV) for k, v in gold.items() if k.startswith('target/')}
    return pred, tgt


def test_losses_and_gradients_match_reference_golden(gold):
    pred, tgt = _loss_inputs(gold)
    loss, grad = T.get_losses(pred, tgt, with_grad=True)
    assert tuple(loss) == T.LOSS_NAMES == tuple(TO.LOSS_NAMES)
    for k in T.LOSS_NAMES:
        ref = float(gold['loss/' + k])
        assert abs(float(loss[k]) - ref) <= 2e-5 * max(1.0, abs(ref)), (k, float(loss[k]), ref)      # fp32 sums, different order
    for k, g in grad.items():
        ref = gold['grad/' + k]
        np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=2e-4, atol=2e-6 * max(1.0, float(np.abs(ref).max())), err_msg=k)
    loss2 = T.get_losses(pred, tgt, with_grad=False)
    for k in T.LOSS_NAMES:
        assert float(loss2[k]) == float(loss[k]), k


def test_losses_match_oracle_at_bench_size():
    B, pad_hw, feat_hw = 32, (384, 1280), (96, 320)
    label = TF.make_labels(B, pad_hw, seed=21, empty_images=(3,), max_objs_per_image=30)
    tgt_np = TO.generate_targets(label, pad_hw, feat_hw)
    pred_np = TF.make_pred(B, feat_hw, seed=22)
    pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in pred_np.items()}
    ref = TO.losses(pt, {k: torch.from_numpy(v) for k, v in tgt_np.items()})
    sum(ref.values()).backward()
    loss, grad = T.get_losses({k: torch.from_numpy(v).to(DEV) for k, v in pred_np.items()},
                              {k: torch.from_numpy(v).to(DEV) for k, v in tgt_np.items()}, with_grad=True)
    for k in T.LOSS_NAMES:
        r = float(ref[k].detach())
        assert abs(float(loss[k]) - r) <= 5e-5 * max(1.0, abs(r)), (k, float(loss[k]), r)
    for k, g in grad.items():
        r = pt[k].grad.numpy()
        np.testing.assert_allclose(g.cpu().numpy(), r, rtol=5e-4, atol=5e-6 * max(1.0, float(np.abs(r).max())), err_msg=k)


def test_losses_empty_batch_asserts_like_the_reference(gold):
    pred, tgt = _loss_inputs(gold)
    tgt['mask_target'] = torch.zeros_like(tgt['mask_target'])
    with pytest.raises(AssertionError):
        T.get_losses(pred, tgt)


def test_clip_adamw_matches_reference_golden(gold):
    ps, gs = TF.make_opt_tensors(seed=7)
    params = [torch.nn.Parameter(torch.from_numpy(p.copy()).to(DEV)) for p in ps]
    opt = T.ClipAdamW(params, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    for step, (lr, b1) in enumerate(TF.OPT_SCHEDULE):
        opt.param_groups[0]['lr'] = lr
        opt.param_groups[0]['betas'] = (b1, 0.99)
        for p, g in zip(params, gs[step]):
            p.grad = torch.from_numpy(g.copy()).to(DEV)
        tn = opt.step()
        ref_n = float(gold[f'opt/norm{step}'])
        assert abs(float(tn) - ref_n) <= 1e-5 * ref_n
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f'opt/p{step}_{i}'], rtol=2e-6, atol=1e-8, err_msg=f'step {step} tensor {i}')
    opt.close()


def test_clip_adamw_matches_torch_on_detector_parameters():
    """All 242 parameter tensors of the detector (19.62 M elements), three steps against torch's own clip + AdamW on the
    same device; the six dead `project` tensors carry no gradient and must stay untouched."""
    import monocon_pytorch_b200 as M
    torch.manual_seed(0)
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    names = [n for n, _ in model.named_parameters()]
    ours = [torch.nn.Parameter(p.detach().clone().to(DEV).contiguous()) for _, p in model.named_parameters()]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    dead = [i for i, n in enumerate(names) if n.startswith(('backbone.level3.project', 'backbone.level4.project'))]
    assert len(dead) == 6 and len(ours) == 242
    opt = T.ClipAdamW(ours, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    ref = torch.optim.AdamW(theirs, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5)
    g = torch.Generator(device='cpu').manual_seed(1)
    for step in range(3):
        scale = 1e-2 if step != 1 else 1.0               # step 1: norm ~ 4400 >> 35, the clip is active
        for i, (a, b) in enumerate(zip(ours, theirs)):
            if i in dead:
                a.grad = b.grad = None
                continue
            gr = (torch.randn(a.shape, generator=g) * scale).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        tn_ref = torch.nn.utils.clip_grad_norm_(theirs, max_norm=35, norm_type=2)
        ref.step()
        tn = opt.step()
        assert abs(float(tn) - float(tn_ref)) <= 1e-5 * float(tn_ref)
    worst = 0.0
    for i, (a, b) in enumerate(zip(ours, theirs)):
        d = (a.detach() - b.detach()).abs().max().item() / max(1e-6, b.detach().abs().max().item())
        worst = max(worst, d)
        if i in dead:
            assert torch.equal(a.detach(), b.detach())
    assert worst <= 2e-6, worst
    opt.close()
