"""The fused clip + AdamW optimiser behind the reference's solver API (engine/monocon_engine.py:39-55, solver/cyclic_scheduler.py):
it must BE a torch.optim.Optimizer called AdamW, otherwise the unmodified CyclicScheduler refuses it.  CPU part: construction,
scheduler, checkpoint format.  GPU part: the scheduled steps against torch.optim.AdamW + clip_grad_norm_."""
import os
import sys

import pytest
import torch

from monocon_pytorch_b200 import train_ops as T

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _reference_scheduler():
    for base in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.exists(os.path.join(base, 'solver', 'cyclic_scheduler.py')):
            import importlib.util
            spec = importlib.util.spec_from_file_location('ref_cyclic_scheduler', os.path.join(base, 'solver', 'cyclic_scheduler.py'))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.CyclicScheduler
    pytest.skip('the reference (solver/cyclic_scheduler.py) is not available here')


def test_is_a_torch_optimizer_named_adamw_and_the_reference_scheduler_accepts_it():
    CyclicScheduler = _reference_scheduler()
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = T.ClipAdamW(params, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    assert isinstance(opt, torch.optim.Optimizer) and opt.__class__.__name__ == 'AdamW'
    sched = CyclicScheduler(opt, total_steps=100)                    # asserts the class name, then _LRScheduler.__init__
    ref = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in params], lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5)
    ref_sched = CyclicScheduler(ref, total_steps=100)
    for _ in range(5):          # the scheduler alone (no optimiser step: that needs the GPU); lr / momentum trajectories must agree
        sched.step()
        ref_sched.step()
        assert opt.param_groups[0]['lr'] == ref.param_groups[0]['lr']
        assert opt.param_groups[0]['betas'] == ref.param_groups[0]['betas']
    with pytest.raises(Exception):
        opt.step()              # CPU tensors: the fused kernel has no CPU fallback and must say so


def test_state_dict_interchanges_with_torch_adamw():
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), weight_decay=1e-2)
    for p in params:
        p.grad = torch.randn_like(p)
    ref.step()
    sd = ref.state_dict()
    ours = T.ClipAdamW(params, lr=5.0)
    ours.load_state_dict(sd)
    assert ours.param_groups[0]['lr'] == 1e-3 and ours.step_count == 1
    for a, p in zip(ours.exp_avg, params):
        assert torch.equal(a, ref.state[p]['exp_avg'])
    back = ours.state_dict()
    assert set(back['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    torch.optim.AdamW(params).load_state_dict(back)                  # torch accepts our checkpoint


@pytest.mark.gpu
def test_scheduled_steps_match_torch_adamw_with_clip():
    CyclicScheduler = _reference_scheduler()
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(3)
    base = [torch.randn(64, 32, 3, 3, generator=g), torch.randn(64, generator=g), torch.randn(10, 64, generator=g)]
    ours = [torch.nn.Parameter(t.clone().to(dev)) for t in base]
    theirs = [torch.nn.Parameter(t.clone().to(dev)) for t in base]
    opt = T.ClipAdamW(ours, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    ref = torch.optim.AdamW(theirs, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5)
    s1, s2 = CyclicScheduler(opt, total_steps=20), CyclicScheduler(ref, total_steps=20)
    for step in range(6):
        for a, b in zip(ours, theirs):
            gr = torch.randn(a.shape, generator=g).to(dev) * (30.0 if step == 2 else 0.1)
            a.grad, b.grad = gr.clone(), gr.clone()
        torch.nn.utils.clip_grad_norm_(theirs, max_norm=35, norm_type=2)
        ref.step(); s2.step()
        opt.step(); s1.step()
    for a, b in zip(ours, theirs):
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max()))
    sd = opt.state_dict()
    assert float(sd['state'][0]['step']) == 6.0
    opt.close()
