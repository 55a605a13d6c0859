"""GPU diagnostic: the bf16 tensor-core training step (Engine(..., 'bf16') + training=2) next to the fp32 FFMA engine on the same
batch -- maps, losses, every parameter gradient (cosine / relative L2), a few resident iterations, and a rough step time.
Usage: python scripts/check_train_tc.py [--big B]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np    # noqa: E402
import torch          # noqa: E402


LAYERS = ['backbone.base_layer', 'backbone.level0', 'backbone.level1', 'backbone.level2.tree1.conv1', 'backbone.level2.tree1.conv2',
          'backbone.level2.tree2.conv1', 'backbone.level2.tree2.conv2', 'backbone.level2.root', 'backbone.level2', 'backbone.level3.tree1.tree1.conv1',
          'backbone.level3.tree1.root', 'backbone.level3', 'backbone.level4', 'backbone.level5.tree1.conv1', 'backbone.level5.tree1.conv2',
          'backbone.level5.root', 'backbone.level5', 'neck.ida_0.proj_1', 'neck.ida_0.up_1.weight', 'neck.ida_0.node_1', 'neck.ida_1.node_2',
          'neck.ida_2.node_1', 'neck.ida_2.node_2', 'neck.feat']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--big', type=int, default=0, help='also time a 384x1280 step at this batch')
    ap.add_argument('--fixture', action='store_true', help='the calibrated (chaotic) test fixture instead of the reference init')
    args = ap.parse_args()
    import monocon_pytorch_b200 as M
    from monocon_pytorch_b200 import engine as E
    from monocon_pytorch_b200 import train_ops as T
    from oracle import fixtures as FX
    from oracle import train_fixtures as TF
    dev = torch.device('cuda', 0)
    B, H, W = 2, 128, 256
    if args.fixture:
        sd = FX.make_state_dict(0)
    else:
        torch.manual_seed(0)
        sd = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False).state_dict()
    img = FX.make_images(B, H, W, seed=31)
    label = TF.make_labels(B, (H, W), seed=32)
    data = {'img': img.to(dev), 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
    tgt = T.TargetGenerator()(data, (B, 64, H // 4, W // 4))
    out = {}
    for prec in ('fp32_simt', 'bf16'):
        eng = E.Engine(dev, B, H, W, prec)
        eng.load_state_dict(sd, training=2)
        pred = eng.forward_train(img.to(dev))
        loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True)
        eng.backward_train(pred, [grad[k].contiguous() for k in E.PRED_NAMES])
        torch.cuda.synchronize()
        grads = {}
        for k, v in sd.items():
            if not torch.is_floating_point(v) or 'running_' in k or k.startswith(('backbone.level3.project.', 'backbone.level4.project.')):
                continue
            grads[k] = eng.get_grad(k, v.shape).double().reshape(-1)
        inter = {}
        for name in LAYERS:
            try:
                inter[name] = eng.debug_tensor(name, B).cpu()
            except Exception as e:          # noqa: BLE001
                inter[name] = None
        out[prec] = {'pred': [p.cpu() for p in pred], 'loss': {k: float(v) for k, v in loss.items()}, 'grads': grads, 'inter': inter}
        print(prec, 'launches', eng.kernel_launches, 'workspace GB', eng.workspace_bytes / 1e9, flush=True)
        eng.close()
    a, b = out['fp32_simt'], out['bf16']
    for name in LAYERS:
        ta, tb = a['inter'][name], b['inter'][name]
        if ta is None or tb is None:
            print(f'layer {name:40s} unavailable')
            continue
        print(f'layer {name:40s} rel-l2 {float((ta - tb).norm() / ta.norm().clamp_min(1e-30)):.3e}  |fp32| {float(ta.norm()):.3e} |bf16| {float(tb.norm()):.3e}')
    for k, pa, pb in zip(E.PRED_NAMES, a['pred'], b['pred']):
        print(f'map {k:28s} rel-to-max {float((pa - pb).abs().max() / pa.abs().max()):.3e}  rel-l2 {float((pa - pb).norm() / pa.norm()):.3e}')
    for k in a['loss']:
        print(f'loss {k:24s} fp32 {a["loss"][k]:.6f} bf16 {b["loss"][k]:.6f}')
    rows = []
    for k, ga in a['grads'].items():
        gb = b['grads'][k]
        cos = float((ga * gb).sum() / (ga.norm() * gb.norm()).clamp_min(1e-300))
        rel = float((ga - gb).norm() / ga.norm().clamp_min(1e-300))
        rows.append((cos, rel, k, float(ga.norm()), float(gb.norm())))
    rows.sort()
    print('worst 40 gradients by cosine (cos, rel-l2, key, |fp32|, |bf16|):')
    for r in rows[:40]:
        print(f'  {r[0]: .4f} {r[1]:.3e} {r[2]:58s} {r[3]:.3e} {r[4]:.3e}')
    cs = np.array([r[0] for r in rows])
    print(f'gradients: {len(rows)} tensors, cosine min {cs.min():.4f} median {np.median(cs):.4f}; > 0.99: {(cs > 0.99).sum()}, > 0.9: {(cs > 0.9).sum()}')

    # resident iterations on one batch
    eng = E.Engine(dev, B, H, W, 'bf16')
    eng.load_state_dict(sd, training=2)
    opt = T.ResidentClipAdamW(eng, lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    totals = []
    for it in range(8):
        pred = eng.forward_train(data['img'])
        loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True)
        totals.append(float(sum(loss.values())))
        eng.backward_train(pred, [grad[k].contiguous() for k in E.PRED_NAMES])
        opt.step()
    print('bf16 resident iterations, total loss:', ' '.join(f'{t:.4f}' for t in totals))
    opt.close()
    eng.close()

    if args.big:
        B2, H2, W2 = args.big, 384, 1280
        eng = E.Engine(dev, B2, H2, W2, 'bf16')
        eng.load_state_dict(sd, training=2)
        opt = T.ResidentClipAdamW(eng)
        label2 = TF.make_labels(B2, (H2, W2), seed=21, max_objs_per_image=8)
        img2 = (torch.randn(B2, 3, H2, W2) * 0.5).to(dev)
        data2 = {'img': img2, 'img_metas': {'pad_shape': [(H2, W2)] * B2}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label2.items()}}
        gen = T.TargetGenerator()
        pred = eng.alloc_pred(B2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        for it in range(4):
            ev[0].record()
            eng.forward_train(img2, out=pred)
            ev[1].record()
            tgt2 = gen(data2, (B2, 64, H2 // 4, W2 // 4))
            loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt2, with_grad=True, check_empty=False)
            ev[2].record()
            eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES])
            ev[3].record()
            opt.step()
            ev[4].record()
            torch.cuda.synchronize()
            print(f'B={B2} 384x1280 it {it}: forward {ev[0].elapsed_time(ev[1]):.2f} ms, targets+losses {ev[1].elapsed_time(ev[2]):.2f}, '
                  f'backward {ev[2].elapsed_time(ev[3]):.2f}, optimizer {ev[3].elapsed_time(ev[4]):.2f}, total {ev[0].elapsed_time(ev[4]):.2f} ms '
                  f'= {B2 / ev[0].elapsed_time(ev[4]) * 1e3:.1f} img/s; loss {float(sum(loss.values())):.3f}; workspace {eng.workspace_bytes / 1e9:.1f} GB', flush=True)
        opt.close()
        eng.close()


if __name__ == '__main__':
    main()
