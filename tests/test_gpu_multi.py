"""2-GPU test (skipped on a single-GPU box): the peer-memory all-gather fused into the decode kernel (mc_gather_*)
returns, on every rank, exactly the rows each rank's own decode produced -- compared with the NCCL all-gather of the
same outputs -- over several generations of both buffers, with different frames per rank."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ok):
    import numpy as np
    import torch.distributed as dist
    from monocon_pytorch_b200 import dist as D
    from monocon_pytorch_b200 import engine as E
    from oracle import fixtures as FX
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        B, H, W, topk = 2, 128, 256, 30
        eng = E.Engine(dev, B, H, W, 'fp32')
        eng.load_state_dict(FX.make_state_dict(0))
        eng.set_option('use_graph', 1)
        pg = D.PeerGather(eng, topk)
        P2_np = FX.kitti_p2(B, 19)
        P2 = torch.from_numpy(P2_np).to(dev)
        invP = E.inverse_viewpad(P2_np).to(dev)
        good = True
        pending = {}
        for step in range(5):                                   # 5 steps: buffers 0,1,0,1,0 -> three generations of buffer 0
            buf = step & 1
            if buf in pending:
                good = good and _check(pg.result(buf), pending.pop(buf), world)
            img = FX.make_images(B, H, W, seed=100 + 10 * step + rank).to(dev)
            pg.infer(img, P2, invP, buf=buf, thres=0.0)
            # reference: the same frames through the plain call + NCCL all-gather
            flat, views = D.alloc_packed(B, topk, dev)
            eng.infer_device(img, P2, invP, topk=topk, thres=0.0, out=views)
            pending[buf] = D.all_gather_decoded(flat, B, topk)
        for buf, ref in pending.items():
            good = good and _check(pg.result(buf), ref, world)
        torch.cuda.synchronize()
        ok[rank] = 1 if good else 0
        eng.close()
    finally:
        dist.destroy_process_group()


def _check(got, ref, world):
    good = True
    for k in ref:
        good = good and got[k].shape == ref[k].shape and torch.equal(got[k], ref[k])
    # different frames per rank -> different rows per slot (the slots are not copies of one another)
    n = ref['box3d'].shape[0] // world
    good = good and not torch.equal(ref['box3d'][:n], ref['box3d'][n:2 * n])
    return bool(good)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_peer_gather_matches_nccl_all_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context('spawn')
    ok = ctx.Array('i', [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def _train_worker(rank, world, port, ok):
    """Data-parallel bf16 tensor-core training (BASELINE.json configs[4]): every rank steps on its own frames / labels; after the
    overlapped average every rank must hold the SAME gradients, equal to the mean of the per-rank gradients of a plain
    (non-overlapped, single pass) backward, and after the optimiser step the same weights."""
    import torch.distributed as dist
    from monocon_pytorch_b200 import dist as D
    from monocon_pytorch_b200 import engine as E
    from monocon_pytorch_b200 import train_ops as T
    from oracle import fixtures as FX
    from oracle import train_fixtures as TF
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        B, H, W = 2, 128, 256
        eng = E.Engine(dev, B, H, W, 'bf16')
        eng.load_state_dict(FX.make_state_dict(0), training=2)
        opt = T.ResidentClipAdamW(eng)
        img = FX.make_images(B, H, W, seed=300 + rank).to(dev)
        label = TF.make_labels(B, (H, W), seed=400 + rank)
        data = {'img': img, 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
        tgt = T.TargetGenerator()(data, (B, 64, H // 4, W // 4))
        pred = eng.forward_train(img)
        loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True)
        dpred = [grad[k].contiguous() for k in E.PRED_NAMES]
        views = D.engine_grad_views(eng)
        # plain pass: local gradients, then their mean over the ranks through one blocking all-reduce each
        eng.backward_train(pred, dpred)
        torch.cuda.synchronize()
        ref = [v.clone() for v in views]
        for r in ref:
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            r.div_(world)
        # overlapped pass on the same forward
        ov = D.OverlappedGradientAverager(views, eng.train_tensor_stages, eng.num_backward_stages, n_segments=5)
        eng.backward_train(pred, dpred, segments=ov.segments, on_segment=ov.on_segment)
        ov.finish()
        torch.cuda.synchronize()
        good = len(ov.segments) >= 3
        worst, worst_key = 0.0, ''
        keys = [k for k, _, _, _ in eng.train_tensors()]
        for key, v, r in zip(keys, views, ref):
            scale = float(r.abs().max())
            if scale == 0:
                continue
            err = float((v - r).abs().max()) / scale
            # The two passes are not bitwise equal: atomics sum in different orders, a last-bit difference flips bf16 roundings of the
            # activation gradients, and tensors that are plain sums over all pixels with heavy cancellation (BatchNorm biases, stem
            # biases, the attention branch) amplify that to a few per cent (measured 2.8e-2 on backbone.level1.1.bias).  A bucket that
            # was not averaged, or written back to the wrong tensor, is O(1): the ranks train on different frames.
            good = good and err < 0.1
            if err > worst:
                worst, worst_key = err, key
        # every rank holds the same averaged gradients -> the same weights after the step
        opt.step()
        torch.cuda.synchronize()
        w = eng.get_param('backbone.level3.tree1.tree1.conv1.weight', (128, 64, 3, 3)).to(dev)
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        # (the clip coefficient comes from an atomically summed norm: the ranks may differ in its last bit)
        wdiff = max(float((ws[0] - x).abs().max()) for x in ws) / float(ws[0].abs().max())
        good = good and wdiff < 1e-6
        print(f'rank {rank}: segments {len(ov.segments)}, overlapped vs plain averaged gradients {worst:.2e} ({worst_key}), weights across ranks {wdiff:.2e}', flush=True)
        ok[rank] = 1 if good else 0
        opt.close()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_data_parallel_bf16_training_step():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context('spawn')
    ok = ctx.Array('i', [0] * world)
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
