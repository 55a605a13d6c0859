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
