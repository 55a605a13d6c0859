"""Device time of the training-side kernels at BASELINE.json configs[2] geometry (B = 32, 384x1280 -> 96x320 maps),
CUDA events, next to the CPU oracle (restatement of the reference's Python / torch path) on the host cores."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                               # noqa: E402
import monocon_pytorch_b200 as M                           # noqa: E402
from monocon_pytorch_b200 import train_ops as T            # noqa: E402
from oracle import train_fixtures as TF, train_oracle as TO   # noqa: E402

dev = torch.device('cuda', 0)
B, pad_hw, feat_hw = 32, (384, 1280), (96, 320)
label = TF.make_labels(B, pad_hw, seed=21, max_objs_per_image=8)
pred_np = TF.make_pred(B, feat_hw, seed=22)
data = {'img': torch.zeros(B, 3, 4, 4, device=dev), 'img_metas': {'pad_shape': [pad_hw] * B},
        'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
pred = {k: torch.from_numpy(v).to(dev) for k, v in pred_np.items()}
gen = T.TargetGenerator()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


tgt = gen(data, (B, 64, *feat_hw))
out = {'B': B, 'objects': int(tgt['mask_target'].sum())}
out['targets_ms'] = timed(lambda: gen(data, (B, 64, *feat_hw)))
out['losses_ms'] = timed(lambda: T.get_losses(pred, tgt, with_grad=False, check_empty=False))
out['losses_with_grad_ms'] = timed(lambda: T.get_losses(pred, tgt, with_grad=True, check_empty=False))
torch.manual_seed(0)
model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
params = [torch.nn.Parameter(p.detach().clone().to(dev).contiguous()) for p in model.parameters()]
for p in params:
    p.grad = torch.randn_like(p) * 1e-2
opt = T.ClipAdamW(params)
out['clip_adamw_ms'] = timed(opt.step)
n_el = sum(p.numel() for p in params)
out['clip_adamw_GBs'] = n_el * 4 * (1 + 3 + 3) / (out['clip_adamw_ms'] * 1e-3) / 1e9       # grad read twice?  sumsq pass 1x + update 4 reads 3 writes
ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
for a, b in zip(ref_params, params):
    a.grad = b.grad.clone()
ref = torch.optim.AdamW(ref_params, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5)


def torch_step():
    torch.nn.utils.clip_grad_norm_(ref_params, max_norm=35, norm_type=2)
    ref.step()


out['torch_clip_adamw_same_gpu_ms'] = timed(torch_step)
# CPU oracle
t0 = time.perf_counter(); TO.generate_targets(label, pad_hw, feat_hw); out['cpu_oracle_targets_ms'] = (time.perf_counter() - t0) * 1e3
pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in pred_np.items()}
tt = {k: v.cpu() for k, v in tgt.items()}
t0 = time.perf_counter(); l = TO.losses(pt, tt); sum(l.values()).backward(); out['cpu_oracle_losses_with_grad_ms'] = (time.perf_counter() - t0) * 1e3
out['cpu_threads'] = torch.get_num_threads()
print(json.dumps(out))
