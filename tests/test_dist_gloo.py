"""CPU, world_size 2, gloo: the shard / pack / all-gather / unpack host logic of the multi-GPU inference path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monocon_pytorch_b200 import dist as D


def _fake_decode(global_idx: int, topk: int):
    rng = np.random.RandomState(100 + global_idx)
    return {'box2d': torch.from_numpy(rng.randn(topk, 5).astype(np.float32)),
            'box3d': torch.from_numpy(rng.randn(topk, 7).astype(np.float32)),
            'labels': torch.from_numpy(rng.randint(0, 3, topk).astype(np.int64)),
            'inds': torch.from_numpy(rng.randint(0, 30720, topk).astype(np.int64)),
            'valid': torch.from_numpy(rng.randint(0, 2, topk).astype(np.uint8))}


def _worker(rank: int, world: int, port: int, n_images: int, topk: int, ok):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        start, stop = D.shard_range(n_images, world, rank)
        B_local = stop - start
        flat, views = D.alloc_packed(B_local, topk, 'cpu')
        for i in range(B_local):                       # what the engine does on the GPU: write into the views
            for k, v in _fake_decode(start + i, topk).items():
                views[k][i].copy_(v)
        out = D.all_gather_decoded(flat, B_local, topk)
        # pipelined form: two batches in flight on alternating buffers, finished out of issue order
        flat2, views2 = D.alloc_packed(B_local, topk, 'cpu')
        for i in range(B_local):
            for k, v in _fake_decode(1000 + start + i, topk).items():
                views2[k][i].copy_(v)
        g1 = torch.empty(world * flat.numel(), dtype=torch.uint8)
        g2 = torch.empty(world * flat.numel(), dtype=torch.uint8)
        fin1 = D.all_gather_decoded_async(flat, B_local, topk, g1)
        fin2 = D.all_gather_decoded_async(flat2, B_local, topk, g2)
        out2, out1 = fin2(), fin1()
        good = True
        for g in range(n_images):
            ref, ref2 = _fake_decode(g, topk), _fake_decode(1000 + g, topk)
            for k in ref:
                good = good and torch.equal(out[k][g], ref[k]) and torch.equal(out1[k][g], ref[k]) and torch.equal(out2[k][g], ref2[k])
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for n in (1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_pack_unpack_roundtrip():
    flat, views = D.alloc_packed(3, 30, 'cpu')
    for i in range(3):
        for k, v in _fake_decode(i, 30).items():
            views[k][i].copy_(v)
    out = D.unpack(flat.clone(), 3, 30)
    for i in range(3):
        for k, v in _fake_decode(i, 30).items():
            assert torch.equal(out[k][i], v)


def test_all_gather_decoded_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world, n_images, topk = 2, 8, 30
    ctx = mp.get_context('spawn')
    ok = ctx.Array('i', [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, topk, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


# ---- data-parallel training: gradient averaging (configs[4]) ---------------------------------------------------------------
def _grad_worker(rank: int, world: int, port: int, ok):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shapes = {'backbone.level2.tree1.conv1.weight': (8, 4, 3, 3), 'backbone.level3.project.0.weight': (6, 3, 1, 1),
                  'backbone.level3.tree1.bn1.bias': (7,), 'neck.ida_0.up_1.weight': (5, 1, 4, 4), 'backbone.level4.project.1.bias': (4,),
                  'head.wh_head.3.weight': (2, 9, 1, 1)}
        named = [(n, torch.nn.Parameter(torch.zeros(s))) for n, s in shapes.items()]
        grads = lambda r: {n: torch.from_numpy(np.random.RandomState(1000 * r + i).randn(*s).astype(np.float32)) for i, (n, s) in enumerate(shapes.items())}
        for n, p in named:
            p.grad = None if n.startswith(D.DEAD_PARAMETER_PREFIXES) else grads(rank)[n].clone()
        red = D.GradientAllReducer(named, bucket_mb=0.0003)          # 78 floats per bucket: several buckets, oversized tensors alone
        assert len(red.buckets) >= 3 and [n for b in red.buckets for n, _ in b] == [n for n, _ in reversed(named) if not n.startswith(D.DEAD_PARAMETER_PREFIXES)]
        red.reduce()
        good = True
        for n, p in named:
            if n.startswith(D.DEAD_PARAMETER_PREFIXES):
                good = good and p.grad is None                        # untouched, as AdamW expects (grad is None -> skipped)
            else:
                mean = sum(grads(r)[n] for r in range(world)) / world
                good = good and torch.allclose(p.grad, mean, rtol=0, atol=1e-7)
        red.reduce()                                                  # buffers are reused: averaging equal gradients is the identity
        for n, p in named:
            if p.grad is not None:
                good = good and torch.allclose(p.grad, sum(grads(r)[n] for r in range(world)) / world, rtol=0, atol=1e-6)
        # the same exchange on bare tensors (what the device-resident loop hands over: views of the engine's gradient buffers)
        bare = lambda r: [torch.from_numpy(np.random.RandomState(77 * r + i).randn(n).astype(np.float32)) for i, n in enumerate((5, 300, 17, 64, 1))]
        mine = bare(rank)
        D.average_tensors_(mine, bucket_mb=0.0003)
        for i, t in enumerate(mine):
            good = good and torch.allclose(t, sum(bare(r)[i] for r in range(world)) / world, rtol=0, atol=1e-7)
        # overlapped form: the stage list cut into segments, each segment's gradients exchanged as soon as its backward is done
        n_stages = 12
        sizes, stages = (40, 7, 300, 5, 64, 9, 120, 33), (11, 11, 9, 8, 5, 5, 2, 0)
        truth = lambda r: [torch.from_numpy(np.random.RandomState(31 * r + i).randn(n).astype(np.float32)) for i, n in enumerate(sizes)]
        views = [torch.zeros(n) for n in sizes]
        ov = D.OverlappedGradientAverager(views, stages, n_stages, n_segments=3)
        good = good and ov.segments[0][1] == n_stages and ov.segments[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(ov.segments, ov.segments[1:]))
        good = good and sorted(i for m in ov.members for i in m) == list(range(len(sizes))) and len(ov.segments) >= 2
        for k, (first, last) in enumerate(ov.segments):               # what engine.backward_train(segments=..., on_segment=...) does
            for i in ov.members[k]:
                views[i].copy_(truth(rank)[i])                        # "the backward of stages [first, last) wrote these gradients"
            ov.on_segment(k)
        ov.finish()
        for i, v in enumerate(views):
            good = good and torch.allclose(v, sum(truth(r)[i] for r in range(world)) / world, rtol=0, atol=1e-7)
        named[0][1].grad = None                                       # a missing gradient is an error, not a silent shift
        try:
            red.reduce()
            good = False
        except RuntimeError:
            pass
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context('spawn')
    ok = ctx.Array('i', [0] * world)
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
