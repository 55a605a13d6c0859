"""CPU: the autograd plumbing of the opt-in training backward (detector._EngineTrainStep / _LossStep) with a stand-in engine and
the loss oracle in place of the CUDA kernels: argument / gradient arity, the dead `project` tensors left without a gradient, the
equal-weights rule of the fused loss node.  The arithmetic behind the two nodes is checked elsewhere (the kernels on the CPU host
shim and on the GPU); this file only makes sure `sum(loss.values()).backward()` reaches every parameter the way the reference's
does (engine/monocon_engine.py:88-91)."""
import pytest
import torch

from monocon_pytorch_b200 import detector as D
from monocon_pytorch_b200 import engine as E
from monocon_pytorch_b200 import train_ops as T
from oracle import train_fixtures as TF
from oracle import train_oracle as TO


class FakeEngine:
    def __init__(self):
        self.calls = []
        self.train_generation = 0

    def forward_train(self, img):
        self.train_generation += 1
        B, _, H, W = img.shape
        g = torch.Generator().manual_seed(1)
        maps = [torch.randn(B, c, H // 4, W // 4, generator=g) for c in E.PRED_CHANNELS]
        maps[0], maps[1] = torch.sigmoid(maps[0]).clamp(1e-4, 1 - 1e-4), torch.sigmoid(maps[1]).clamp(1e-4, 1 - 1e-4)
        return maps

    def backward_train(self, pred, dpred):
        assert len(pred) == len(dpred) == 10 and all(d.shape == p.shape and d.is_contiguous() for p, d in zip(pred, dpred))
        self.calls.append([d.clone() for d in dpred])

    def get_grad(self, key, shape):
        return torch.full(tuple(shape), 2.0)


def _oracle_get_losses(pred_dict, target_dict, max_objs=30, with_grad=False, check_empty=True):
    with torch.enable_grad():                                    # Function.forward runs with grad mode off
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in pred_dict.items()}
        loss = TO.losses(leaves, target_dict)
        if not with_grad:
            return {k: v.detach() for k, v in loss.items()}
        sum(loss.values()).backward()
    return {k: v.detach() for k, v in loss.items()}, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


def test_loss_backward_reaches_every_live_parameter(monkeypatch):
    monkeypatch.setattr(T, 'get_losses', _oracle_get_losses)
    B, H, W = 2, 64, 128
    names = ['backbone.level2.tree1.conv1.weight', 'backbone.level3.project.0.weight', 'head.wh_head.3.bias', 'backbone.level4.project.1.bias']
    params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2)),
              torch.nn.Parameter(torch.zeros(5))]
    eng = FakeEngine()
    maps = D._EngineTrainStep.apply(eng, torch.zeros(B, 3, H, W), names, *params)
    assert len(maps) == 10 and all(m.requires_grad for m in maps)
    label = TF.make_labels(B, (H, W), seed=5)
    tgt = {k: torch.from_numpy(v) for k, v in TO.generate_targets(label, (H, W), (H // 4, W // 4)).items()}
    losses = D._LossStep.apply(tgt, 30, *maps)
    assert len(losses) == len(T.LOSS_NAMES)
    ref_loss, ref_grad = _oracle_get_losses(dict(zip(E.PRED_NAMES, [m.detach() for m in maps])), tgt, with_grad=True)
    for k, v in zip(T.LOSS_NAMES, losses):
        assert float(v.detach()) == float(ref_loss[k])
    sum(losses).backward()
    assert len(eng.calls) == 1                                   # one engine backward per step
    for d, k in zip(eng.calls[0], E.PRED_NAMES):
        assert torch.equal(d, ref_grad[k]), k                    # dL/dpred of the plain sum arrives unchanged
    assert torch.equal(params[0].grad, torch.full((3, 2), 2.0)) and torch.equal(params[2].grad, torch.full((2,), 2.0))
    assert params[1].grad is None and params[3].grad is None     # the dead `project` tensors, as in the reference


def test_unequal_loss_weights_are_refused(monkeypatch):
    monkeypatch.setattr(T, 'get_losses', _oracle_get_losses)
    B, H, W = 2, 64, 128
    eng = FakeEngine()
    p = torch.nn.Parameter(torch.zeros(1))
    maps = D._EngineTrainStep.apply(eng, torch.zeros(B, 3, H, W), ['head.wh_head.3.bias'], p)
    label = TF.make_labels(B, (H, W), seed=5)
    tgt = {k: torch.from_numpy(v) for k, v in TO.generate_targets(label, (H, W), (H // 4, W // 4)).items()}
    losses = D._LossStep.apply(tgt, 30, *maps)
    with pytest.raises(NotImplementedError):
        (losses[0] * 2 + sum(losses[1:])).backward()
    eng2 = FakeEngine()
    maps = D._EngineTrainStep.apply(eng2, torch.zeros(B, 3, H, W), ['head.wh_head.3.bias'], p)
    losses = D._LossStep.apply(tgt, 30, *maps)
    (0.5 * sum(losses)).backward()                               # a common factor is fine
    assert len(eng2.calls) == 1


def test_backward_through_a_stale_forward_is_refused():
    """The engine keeps the activations of ONE train-mode forward: back-propagating an older graph must raise, not silently use the
    newer batch's activations."""
    eng = FakeEngine()
    names = ['backbone.level2.tree1.conv1.weight']
    p = torch.nn.Parameter(torch.zeros(3, 2))
    old = D._EngineTrainStep.apply(eng, torch.zeros(2, 3, 64, 128), names, p)
    new = D._EngineTrainStep.apply(eng, torch.zeros(2, 3, 64, 128), names, p)
    sum(m.sum() for m in new).backward()
    assert len(eng.calls) == 1
    with pytest.raises(RuntimeError, match='not the most recent'):
        sum(m.sum() for m in old).backward()
