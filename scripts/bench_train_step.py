"""Device time of ONE device-resident training iteration at BASELINE.json configs[2] geometry (default B = 32, 384x1280):
forward_train -> TargetGenerator -> get_losses(with_grad) -> backward_train -> ResidentClipAdamW.step, CUDA events on the
current stream, with the split per phase.  EXPERIMENTAL: the backward kernels are the correctness-first fp32 set of
csrc/train_backward.cu (no shared-memory reuse), so this line is the BASELINE the tensor-core dgrad / wgrad have to beat, not a
result.  Not the repo's bench contract (bench.py keeps measuring BASELINE.json's headline metric).

    python scripts/bench_train_step.py [--batch 32] [--steps 5] [--warmup 2] [--hw 384 1280]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                               # noqa: E402
import monocon_pytorch_b200 as M                           # noqa: E402
from monocon_pytorch_b200 import engine as E               # noqa: E402
from monocon_pytorch_b200 import train_ops as T            # noqa: E402
from oracle import train_fixtures as TF                    # noqa: E402  (synthetic labels only)

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--hw', type=int, nargs=2, default=(384, 1280))
args = ap.parse_args()

dev = torch.device('cuda', 0)
B, (H, W) = args.batch, args.hw
torch.manual_seed(0)
model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
eng = E.Engine(dev, B, H, W, 'fp32')
eng.load_state_dict(model.state_dict(), training=2)
opt = T.ResidentClipAdamW(eng)
label = TF.make_labels(B, (H, W), seed=21, max_objs_per_image=8)
g = torch.Generator().manual_seed(1)
imgs = [(torch.randn(B, 3, H, W, generator=g) * 0.5).to(dev) for _ in range(2)]
data = {'img': imgs[0], 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
gen = T.TargetGenerator()
pred = eng.alloc_pred(B)
names = ('forward', 'targets+losses', 'backward', 'optimizer')


def iteration(i, ev=None):
    mark = (lambda k: ev[k].record()) if ev is not None else (lambda k: None)
    mark(0)
    eng.forward_train(imgs[i % 2], out=pred)
    mark(1)
    tgt = gen(data, (B, 64, H // 4, W // 4))
    loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True, check_empty=False)
    mark(2)
    eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES])
    mark(3)
    opt.step()
    mark(4)
    return loss


for i in range(max(args.warmup, 1)):
    loss = iteration(i)
torch.cuda.synchronize()
phase = [0.0] * 4
total = 0.0
for i in range(args.steps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    loss = iteration(i, ev)
    torch.cuda.synchronize()
    for k in range(4):
        phase[k] += ev[k].elapsed_time(ev[k + 1])
    total += ev[0].elapsed_time(ev[4])
ms = total / args.steps
print(json.dumps({'metric': 'training iteration (fwd + losses + bwd + clip/AdamW), device-resident', 'value': B / (ms * 1e-3), 'unit': 'images/sec',
                  'ms_per_step': ms, 'batch': B, 'hw': [H, W], 'dtype': 'f32', 'steps': args.steps,
                  'phases_ms': {n: p / args.steps for n, p in zip(names, phase)},
                  'total_loss_last': float(sum(loss.values())), 'workspace_GB': eng.workspace_bytes / 1e9,
                  'note': 'EXPERIMENTAL baseline: FFMA forward, sync-free fp32 backward kernels (csrc/train_backward.cu)'}))
