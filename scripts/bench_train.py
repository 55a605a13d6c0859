"""`bench.py --mode train`: one device-resident training iteration per step at BASELINE.json configs[2] geometry (batch 32
per GPU, 384x1280): forward_train -> TargetGenerator -> losses + dL/dpred -> backward_train -> [gradient average over NCCL at
N > 1, overlapped with the backward walk: configs[4]] -> fused clip + AdamW on the engine's packed buffers.  Reference step:
engine/monocon_engine.py:80-102.  CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.

A second metric next to the repo's headline (forward + decode); the JSON line follows the same contract.  `--train-precision bf16`
(default, what configs[2] names): the tensor-core step of csrc/train_engine_tc.cu -- bf16 activations / gradients, fp32 master
weights, statistics and parameter gradients -- with `roofline` against the measured bf16 tensor peak; `fp32_simt`: its FFMA twin
(csrc/train_backward.cu), quoted against the FP32 FFMA peak."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 384, 1280


def main(args, rank, local_rank, world, return_line=False):
    import torch
    import torch.distributed as dist
    import monocon_pytorch_b200 as M
    from monocon_pytorch_b200 import dist as mcdist
    from monocon_pytorch_b200 import engine as E
    from monocon_pytorch_b200 import train_ops as T
    from oracle import train_fixtures as TF                    # synthetic labels only (test-infrastructure generator, not a checker here)

    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback for the product path'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    prec = getattr(args, 'train_precision', 'bf16')
    tc = prec == 'bf16'
    eng = E.Engine(dev, B, H, W, prec)
    eng.load_state_dict(model.state_dict(), training=2)
    opt = T.ResidentClipAdamW(eng)
    label = TF.make_labels(B, (H, W), seed=21 + rank, max_objs_per_image=8)
    g = torch.Generator().manual_seed(1 + rank)
    n_rot = 2                                                   # 2 x 189 MB of frames; the step itself streams ~10 GB of activations
    imgs_host = [(torch.randn(B, 3, H, W, generator=g) * 0.5).pin_memory() for _ in range(n_rot)]
    imgs = [t.to(dev) for t in imgs_host]
    data = {'img': imgs[0], 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
    gen = T.TargetGenerator()
    pred = eng.alloc_pred(B)
    averager = None
    if world > 1:
        views = mcdist.engine_grad_views(eng)
        averager = mcdist.OverlappedGradientAverager(views, eng.train_tensor_stages, eng.num_backward_stages, n_segments=8 if tc else 4)
    names = ('forward', 'targets+losses', 'backward(+allreduce issue)', 'allreduce wait', 'optimizer')

    def iteration(i, ev=None):
        mark = (lambda k: ev[k].record()) if ev is not None else (lambda k: None)
        mark(0)
        eng.forward_train(imgs[i % n_rot], out=pred)
        mark(1)
        tgt = gen(data, (B, 64, H // 4, W // 4))
        loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True, check_empty=False)
        mark(2)
        if averager is not None:
            eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES], segments=averager.segments, on_segment=averager.on_segment)
            mark(3)
            averager.finish()
        else:
            eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES])
            mark(3)
        mark(4)
        opt.step()
        mark(5)
        return loss

    for i in range(Wm):
        loss = iteration(i)
    torch.cuda.synchronize()
    from bench import ClockSampler
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = [0.0] * 5
    evs = []
    ev0.record()
    for i in range(K):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        loss = iteration(i, ev)
        evs.append(ev)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for ev in evs:
        for k in range(5):
            phase[k] += ev[k].elapsed_time(ev[k + 1])
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # end to end: host frames + labels in (pinned), the step's loss value out, every step.  A two-slot loop, as a training loop with a
    # prefetching loader runs it: the H2D copy of batch i + 1 goes on a copy stream while batch i computes, and the loss of step i is
    # read (D2H, blocking) after step i + 1 has been enqueued, so the host stays one step ahead of the device.
    lab_host = {k: torch.from_numpy(v).pin_memory() for k, v in label.items()}
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{'img': torch.empty_like(imgs[0]), 'label': {k: torch.empty_like(v, device=dev) for k, v in lab_host.items()},
              'ready': torch.cuda.Event(), 'free': torch.cuda.Event()} for _ in range(2)]

    def prefetch(i):
        sl = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl['free'])                  # the step that last used this slot has finished with it
            sl['img'].copy_(imgs_host[i % n_rot], non_blocking=True)
            for k, v in lab_host.items():
                sl['label'][k].copy_(v, non_blocking=True)
            sl['ready'].record(copy_stream)

    def e2e_launch(i):
        sl = slots[i % 2]
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(sl['ready'])
        d = {'img': sl['img'], 'img_metas': data['img_metas'], 'label': sl['label']}
        eng.forward_train(sl['img'], out=pred)
        tgt = gen(d, (B, 64, H // 4, W // 4))
        ls, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True, check_empty=False)
        if averager is not None:
            eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES], segments=averager.segments, on_segment=averager.on_segment)
            averager.finish()
        else:
            eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES])
        opt.step()
        sl['free'].record(cur)
        return sum(ls.values())                               # device scalar; read one step later

    for sl in slots:
        sl['free'].record(torch.cuda.current_stream(dev))
    prefetch(0)
    float(e2e_launch(0))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    Ke = max(2, min(K, 8))
    prefetch(0)
    t0 = time.perf_counter()
    pending = None
    for i in range(Ke):
        prefetch(i + 1)
        loss_dev = e2e_launch(i)
        if pending is not None:
            last = float(pending)                             # D2H of step i - 1's result while step i runs
        pending = loss_dev
    last = float(pending)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms = ms_total / K
    value = world * B * K / (ms_total * 1e-3)
    fl_img = eng.flops_per_image
    # forward + dgrad + wgrad = 3x the forward convolution FLOPs (the stem needs no dgrad; neglected)
    train_tflops = 3 * fl_img * B / (ms * 1e-3) / 1e12
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                      # 148 SMs x 128 FP32 lanes x FMA at 1965 MHz = 74.4 TFLOP/s
    from bench import peaks
    pk = peaks()
    if tc:
        roof = {'bound': 'tensor', 'achieved': train_tflops, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': train_tflops / pk['tf_burst'],
                'traffic': None, 'peak_source': pk['src'] + ': burst bf16 dense; sustained = %.0f' % pk['tf_sustained'],
                'kernel': 'whole step (conv_tc* forward + dgrad, wgrad_tc_kernel; bandwidth kernels included in the time)',
                'note': 'ALGORITHMIC FLOPs = 3 x forward convolution FLOPs (forward + dgrad + wgrad) over the whole step time; the five stride-2 '
                        'layers execute 4x their dgrad / wgrad MMAs on zero-inserted gradients, and about half of the step is HBM-bound '
                        'BatchNorm / head / pooling traffic (per-kernel split: profiles/r02_train_tc_launches.txt)'}
    else:
        roof = {'bound': 'fp32-ffma', 'achieved': train_tflops, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': train_tflops / fp32_peak,
                'traffic': None, 'peak_source': 'nominal: 148 SMs x 128 FP32 lanes x 2 x 1965 MHz (no measured FFMA peak in MEASURED_PEAKS.json)',
                'note': 'algorithmic FLOPs = 3 x forward convolution FLOPs (forward + dgrad + wgrad); the kernels are fp32 SIMT'}
    h2d = B * 3 * H * W * 4 + sum(v.numel() * v.element_size() for v in lab_host.values())
    nbytes_grad = sum(m for _, _, _, m in eng.train_tensors()) * 4
    line = {'metric': 'images/sec training step (fwd + losses + bwd + clip/AdamW) at 384x1280', 'value': value, 'unit': 'images/s',
            'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': ('bf16 operands / activations / activation gradients on tcgen05, fp32 accumulate, fp32 master weights, statistics and '
                      'parameter gradients') if tc else 'f32 (FFMA forward and backward kernels)', 'data': 'synthetic',
            'config': {'workload': f'batch={B}/GPU training iteration 384x1280 (BASELINE.json configs[2]; configs[4] data parallel at N>1)',
                       'arch': 'DLA-34 + DLAUp + MonoCon heads, reference random init (seed 0)', 'global_batch': B * world, 'train_precision': prec,
                       'optimizer': 'clip_grad_norm_(35) + AdamW fused over the engine-resident packed parameters',
                       'parallelism': (f'dp{world}: rank-local BatchNorm statistics (the reference has no SyncBN), gradient average = '
                                       f'{len(averager.segments)} NCCL all-reduces of about {nbytes_grad / len(averager.segments) / 1e6:.0f} MB each, issued as '
                                       'the backward walk finishes each quarter of the stage list') if world > 1 else 'single GPU',
                       'l2': f'{eng.workspace_bytes / 1e9:.1f} GB workspace streamed per step: working set >> 126 MB L2'},
            'clocks': clocks,
            'phases_ms': {n: p / K for n, p in zip(names, phase)},
            'e2e': {'value': world * B * Ke / e2e_s, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'api': 'Engine.forward_train / train_ops.TargetGenerator / get_losses / Engine.backward_train / ResidentClipAdamW.step with '
                           'pinned host frames + labels copied in (copy stream, two slots: batch i + 1 while batch i computes) and the total loss of '
                           'every step read back (one step late)'},
            'gpu_launches': eng.kernel_launches * K,
            'roofline': roof,
            'total_loss_last': last, 'gradient_bytes': nbytes_grad, 'workspace_GB': eng.workspace_bytes / 1e9}
    opt.close()
    eng.close()
    if return_line:
        return line
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
