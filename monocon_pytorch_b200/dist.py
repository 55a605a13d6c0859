"""Multi-GPU inference: the batch dimension shards across ranks (one process per GPU); every rank runs forward +
decode on its shard and the fixed-shape decoded boxes are exchanged with ONE all-gather (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The reference has no multi-GPU path (README.MD:11,15); this is the B200-side
addition described in SURVEY.md §8(e).  Images are independent in eval mode (running-stat BN, per-sample AttnBN),
so there is no other data-path collective.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

# name, dtype, trailing shape per (image, detection)
FIELDS = (('box2d', torch.float32, (5,)), ('box3d', torch.float32, (7,)), ('labels', torch.int64, ()),
          ('inds', torch.int64, ()), ('valid', torch.uint8, ()))


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of n images for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _field_bytes(B: int, topk: int):
    out, off = [], 0
    for name, dt, tail in FIELDS:
        n = B * topk
        for t in tail:
            n *= t
        nbytes = n * torch.tensor([], dtype=dt).element_size()
        out.append((name, dt, (B, topk) + tail, off, nbytes))
        off += (nbytes + 15) // 16 * 16
    return out, off


def alloc_packed(B: int, topk: int, device) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One flat byte buffer holding all decode outputs of B images, plus typed views into it.  The engine writes
    straight into the views, so the all-gather needs no packing kernel."""
    fields, total = _field_bytes(B, topk)
    flat = torch.zeros(total, dtype=torch.uint8, device=device)
    views = {name: flat[off:off + nb].view(dt).view(shape) for name, dt, shape, off, nb in fields}
    return flat, views


def unpack(flat: torch.Tensor, B: int, topk: int) -> Dict[str, torch.Tensor]:
    fields, total = _field_bytes(B, topk)
    assert flat.numel() == total
    return {name: flat[off:off + nb].view(dt).view(shape) for name, dt, shape, off, nb in fields}


def _split_gathered(gathered: torch.Tensor, nbytes: int, world: int, B_local: int, topk: int):
    parts = [unpack(gathered[r * nbytes:(r + 1) * nbytes], B_local, topk) for r in range(world)]
    return {name: torch.cat([p[name] for p in parts], 0) for name, _, _ in FIELDS}


def all_gather_decoded(flat_local: torch.Tensor, B_local: int, topk: int, gathered: torch.Tensor = None):
    """All ranks must hold the same B_local.  Returns {field: (world * B_local, topk, ...)} in rank order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return unpack(flat_local, B_local, topk)
    if gathered is None:
        gathered = torch.empty(world * flat_local.numel(), dtype=torch.uint8, device=flat_local.device)
    dist.all_gather_into_tensor(gathered, flat_local)
    return _split_gathered(gathered, flat_local.numel(), world, B_local, topk)


def all_gather_decoded_async(flat_local: torch.Tensor, B_local: int, topk: int, gathered: torch.Tensor):
    """Pipelined form: issues the all-gather asynchronously (on NCCL's stream, ordered after the work already enqueued
    on the current stream) and returns ``finish``; calling ``finish()`` waits for it and returns the gathered fields.
    The caller keeps `flat_local` / `gathered` untouched until then (alternate two buffer pairs to overlap the
    collective of batch i with the forward of batch i + 1, as bench.py does)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return lambda: unpack(flat_local, B_local, topk)
    work = dist.all_gather_into_tensor(gathered, flat_local, async_op=True)

    def finish():
        work.wait()
        return _split_gathered(gathered, flat_local.numel(), world, B_local, topk)
    return finish


class PeerGather:
    """All-gather of the decode outputs over peer memory, fused into the decode kernel (include/monocon_b200.h,
    mc_gather_*): no NCCL kernel on the data path, the rows travel as plain NVLink stores issued by the decode CTAs.

        pg = PeerGather(engine, topk)          # collective: exchanges the CUDA IPC handles through torch.distributed
        pg.infer(img, P2, invP, buf=i & 1)     # forward + decode + scatter of batch i into buffer i & 1 of every rank
        fields = pg.result(buf)                # waits (on the current stream) and returns {field: (world * B, topk, ...)}

    Alternate the two buffers and ask for the result one batch late to overlap the exchange with the next forward."""

    def __init__(self, engine, topk: int = 30):
        self.engine, self.topk = engine, topk
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.B = engine.max_batch
        mine = engine.gather_create(self.world, self.rank, topk)
        if self.world > 1:
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(engine.device)
            allh = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(allh, t)
            handles = b''.join(bytes(h.cpu().numpy().tobytes()) for h in allh)
        else:
            handles = mine
        # connect, then agree on the outcome with ONE collective that every rank reaches whatever happened locally (a rank
        # that raised early would otherwise leave the others inside a different collective)
        err = None
        try:
            engine.gather_connect(handles)
            self.slot_bytes = engine.gather_slot_bytes()
            fields, total = _field_bytes(self.B, topk)
            assert total == self.slot_bytes, (total, self.slot_bytes)
            self._views = [_wrap_device_bytes(engine.gather_buffer_ptr(buf), self.world * self.slot_bytes, engine.device)
                           for buf in range(2)]
        except Exception as e:                  # noqa: BLE001
            err = e
        if self.world > 1:
            ok = torch.tensor([0 if err is not None else 1], device=engine.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # also the barrier: every rank has mapped every block
            if int(ok.item()) == 0 and err is None:
                err = RuntimeError('another rank could not map the peer gather blocks')
        if err is not None:
            raise err

    def infer(self, img, P2, invP, buf: int, thres: float = 0.4) -> None:
        self.engine.infer_device_gather(img, P2, invP, buf, thres)

    def wait(self, buf: int) -> None:
        self.engine.gather_wait(buf)

    def result(self, buf: int) -> Dict[str, torch.Tensor]:
        self.wait(buf)
        return _split_gathered(self._views[buf], self.slot_bytes, self.world, self.B, self.topk)

    def gathered_bytes(self, buf: int) -> torch.Tensor:
        """The raw gather block of buffer `buf` (world slots, rank order) -- for checks against an NCCL all-gather."""
        return self._views[buf]

    def local_packed(self, buf: int) -> torch.Tensor:
        """A copy of this rank's own slot of buffer `buf` (what an NCCL all-gather of the same outputs would send)."""
        return self._views[buf][self.rank * self.slot_bytes:(self.rank + 1) * self.slot_bytes].clone()


def _wrap_device_bytes(ptr: int, nbytes: int, device) -> torch.Tensor:
    """A uint8 tensor view of engine-owned device memory (no copy, no ownership)."""
    class _Mem:
        __cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}
    return torch.as_tensor(_Mem(), device=device)


# ------------------------------------------------------------------------------------------------------------------------
# Data-parallel training (BASELINE.json configs[4]): the batch shards across ranks, BatchNorm statistics stay rank-local (the
# reference has no SyncBN, model/norm/attentive_norm.py:52,91), and the one exchange of the step is the average of the
# parameter gradients -- what torch's DistributedDataParallel would do around the reference's model, written out because the
# engine produces all gradients at the end of its own backward walk rather than through autograd hooks.
# ------------------------------------------------------------------------------------------------------------------------
DEAD_PARAMETER_PREFIXES = ('backbone.level3.project.', 'backbone.level4.project.')      # no gradient in the reference


class GradientAllReducer:
    """Bucketed all-reduce (average) of ``param.grad`` over a process group; NCCL on the GPU box, gloo in the CPU tests.

    * The parameter list is fixed at construction and must be the same on every rank (same names, same order); the six dead
      ``project`` tensors are excluded statically -- the DDP "unused parameter" hazard never arises.
    * Buckets are filled in REVERSE registration order (the order the backward walk finishes gradients) and are ``bucket_mb``
      large, so that a future hook-driven backward can launch bucket k while bucket k+1 is still being produced; today
      ``reduce()`` is called once after ``loss.backward()`` and issues every bucket asynchronously before waiting for the first.
    * ``clip_grad_norm_`` / ``ClipAdamW`` then see the averaged gradients, i.e. the global norm -- no extra scalar exchange.
    """

    def __init__(self, named_parameters, group=None, bucket_mb: float = 25.0):
        self.group = group
        self.named = [(n, p) for n, p in named_parameters if p.requires_grad and not n.startswith(DEAD_PARAMETER_PREFIXES)]
        assert self.named, 'no parameters to reduce'
        limit = max(1, int(bucket_mb * (1 << 20)) // 4)
        self.buckets, cur, size = [], [], 0
        for n, p in reversed(self.named):
            if cur and size + p.numel() > limit:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append((n, p))
            size += p.numel()
        self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)

    def _buffer(self, i: int) -> torch.Tensor:
        if self._flat[i] is None:
            p0 = self.buckets[i][0][1]
            self._flat[i] = torch.empty(sum(p.numel() for _, p in self.buckets[i]), dtype=torch.float32, device=p0.device)
        return self._flat[i]

    def reduce(self) -> None:
        world = dist.get_world_size(self.group)
        work = []
        for i, bucket in enumerate(self.buckets):
            flat, off = self._buffer(i), 0
            for n, p in bucket:
                if p.grad is None:              # a rank that skipped a tensor would shift every later gradient of the bucket
                    raise RuntimeError(f'GradientAllReducer: {n} has no gradient on this rank')
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
            work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, bucket in enumerate(self.buckets):
            work[i].wait()
            flat, off = self._flat[i], 0
            flat.div_(world)
            for _, p in bucket:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()


def average_tensors_(tensors, group=None, bucket_mb: float = 25.0) -> None:
    """In-place bucketed average of a fixed list of float32 tensors over the process group (same list, same order on every
    rank).  The device-resident training loop uses it on views of the engine's own gradient buffers (``engine_grad_views``), so
    the exchange needs no ``param.grad`` at all."""
    world = dist.get_world_size(group)
    limit = max(1, int(bucket_mb * (1 << 20)) // 4)
    buckets, cur, size = [], [], 0
    for t in tensors:
        assert t.dtype == torch.float32 and t.is_contiguous()
        if cur and size + t.numel() > limit:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(t)
        size += t.numel()
    if cur:
        buckets.append(cur)
    flats, work = [], []
    for b in buckets:
        flat = torch.cat([t.reshape(-1) for t in b]) if len(b) > 1 else b[0].reshape(-1)      # a single tensor is reduced in place
        flats.append(flat)
        work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for b, flat, w in zip(buckets, flats, work):
        w.wait()
        flat.div_(world)
        if len(b) > 1:
            off = 0
            for t in b:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()


def engine_grad_views(engine):
    """float32 tensor views (no copy) of the gradient buffers of an engine loaded with ``training=2``, in ``train_tensors()``
    order -- what ``average_tensors_`` averages between ``backward_train`` and ``ResidentClipAdamW.step`` under data parallelism."""
    return [_wrap_device_bytes(g, m * 4, engine.device).view(torch.float32) for _, _, g, m in engine.train_tensors()]


class OverlappedGradientAverager:
    """Data-parallel gradient averaging overlapped with the backward walk (BASELINE.json configs[4]).

    ``views`` are float32 tensors over the gradient buffers (``engine_grad_views``) and ``stages`` the stage of the stage list whose
    backward finishes each of them (``engine.train_tensor_stages``).  The stage list is cut into ``n_segments`` contiguous pieces of
    about equal gradient volume; ``engine.backward_train(pred, dpred, segments=self.segments, on_segment=self.on_segment)`` walks
    them from the end, and after each piece the gradients that became final are packed into one flat bucket and all-reduced
    asynchronously -- the collective queues behind the kernels already launched and runs next to those of the following pieces.
    ``finish()`` waits, divides by the world size and writes the averages back.  World size 1: no-ops."""

    def __init__(self, views, stages, n_stages: int, n_segments: int = 4, group=None):
        assert len(views) == len(stages) and n_stages >= 1
        self.group, self.views, self.stages = group, list(views), list(stages)
        total = sum(v.numel() for v in self.views)
        per_stage = [0] * n_stages
        for v, st in zip(self.views, self.stages):
            assert 0 <= st < n_stages
            per_stage[st] += v.numel()
        cuts, acc, want = [n_stages], 0, total / max(1, n_segments)
        for st in range(n_stages - 1, 0, -1):                     # from the end of the stage list, as the backward walks it
            acc += per_stage[st]
            if acc >= want and len(cuts) < n_segments:
                cuts.append(st)
                acc = 0
        cuts.append(0)
        self.segments = [(cuts[i + 1], cuts[i]) for i in range(len(cuts) - 1) if cuts[i + 1] < cuts[i]]
        self.members = [[i for i, st in enumerate(self.stages) if first <= st < last] for first, last in self.segments]
        self._flat = [None] * len(self.segments)
        self._work = [None] * len(self.segments)

    def on_segment(self, k: int) -> None:
        idx = self.members[k]
        if not idx or dist.get_world_size(self.group) == 1:
            return
        self._flat[k] = torch.cat([self.views[i].reshape(-1) for i in idx])
        # NCCL averages inside the collective; gloo (the CPU tests) sums and finish() divides
        self._avg = dist.get_backend(self.group) == 'nccl'
        self._work[k] = dist.all_reduce(self._flat[k], op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        world = dist.get_world_size(self.group)
        for k, idx in enumerate(self.members):
            if self._work[k] is None:
                continue
            self._work[k].wait()
            flat = self._flat[k]
            if not getattr(self, '_avg', False):
                flat.div_(world)
            dst = [self.views[i].reshape(-1) for i in idx]
            # one multi-tensor copy per bucket instead of one kernel per parameter (about 370 tensors in all)
            torch._foreach_copy_(dst, list(flat.split([d.numel() for d in dst])))
            self._work[k] = None
            self._flat[k] = None
