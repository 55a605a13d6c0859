#!/usr/bin/env python
"""bench.py -- images/sec of the MonoCon forward + decode hot path at 384x1280 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batch of 16 synthetic 384x1280 frames per GPU through forward + decode
(BASELINE.json configs[1]); at N > 1 every rank runs its own shard and the decoded boxes are
all-gathered with one NCCL collective inside the timed region (configs[3]).  Prints ONE JSON line.

* value     device-resident inputs, CUDA-graph replay, CUDA events, max over ranks
* e2e       the same metric through the host-buffer C-ABI call (mc_infer_host): pinned host frames in,
            decoded boxes on the host out, H2D/D2H inside the timed region
* roofline  the dominant kernel family (the tcgen05 implicit-GEMM convolution): algorithmic conv FLOPs /
            summed per-launch durations measured live with CUDA events (mc_profile_stages)
* cpu_baseline  the CPU oracle (a restatement of the reference's PyTorch path) on the host cores
* --impl reference  times that CPU implementation on the box's host cores with all threads
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 384, 1280
METRIC = 'images/sec fwd+decode at 384x1280'
UNIT = 'images/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-gather', action='store_true', help='diagnostic only: skip the all-gather at N > 1 (not a valid bench line)')
    ap.add_argument('--gather', default='p2p', choices=['p2p', 'nccl'],
                    help='N > 1: p2p = all-gather fused into the decode kernel over peer memory (mc_gather_*), nccl = torch.distributed')
    ap.add_argument('--stage-table', default='', help='write the per-stage timing table to this file')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


def synthetic_state_dict():
    """Random-init weights of the reference architecture (its own init distributions), seed 0."""
    import torch
    import monocon_pytorch_b200 as M
    torch.manual_seed(0)
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    return {k: v.clone() for k, v in model.state_dict().items()}


def synthetic_frames(batch, seed):
    """randn * 0.01: the tie-free recipe of SURVEY.md §8(d) for the reference's random init."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, H, W, generator=g) * 0.01


def kitti_p2(batch):
    import numpy as np
    base = np.array([[721.5377, 0., 609.5593, 44.85728], [0., 721.5377, 172.854, 0.2163791], [0., 0., 1., 0.002745884]],
                    dtype=np.float32)
    return np.repeat(base[None], batch, 0)


class ClockSampler:
    """Polls NVML (SM clock, power, clock-event reasons) from a thread every ~5 ms during the timed region."""

    def __init__(self, index):
        import threading
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = False
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:                       # pragma: no cover
            self.err = repr(e)
            self.th = None

    def _run(self):
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:                   # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.005)

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples: ' + str(self.err)]}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz, 'power_w_max': max(self.power),
                'samples': len(self.samples), 'reasons': sorted(self.reasons), 'source': 'NVML polled every 5 ms in the timed region'}


def cpu_oracle_rate(sd, seconds_budget=15.0, warmup=1, max_iters=40):
    """The CPU oracle (reference restatement) on the host cores: images/s for B=1 forward + decode."""
    import torch
    from oracle import monocon_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img = synthetic_frames(1, 100)
    P2 = kitti_p2(1)
    sd = {k: v.float() for k, v in sd.items()}
    for _ in range(warmup):
        O.forward_and_decode(sd, img, P2)
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_start) < seconds_budget:
        t0 = time.perf_counter()
        O.forward_and_decode(sd, img, P2)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {'value': 1.0 / med, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{len(times)} x (1 frame 384x1280 fp32 forward+decode), median {med * 1e3:.1f} ms, torch CPU ops, '
                      f'{cores} threads'}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the reference is
    Python and cannot travel to the GPU box) on the host cores, one frame per step."""
    if rank != 0:
        return
    import torch
    from oracle import monocon_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_state_dict()
    img = synthetic_frames(1, 100)
    P2 = kitti_p2(1)
    for _ in range(min(args.warmup, 3)):
        O.forward_and_decode(sd, img, P2)
    steps = min(args.steps, 30)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward_and_decode(sd, img, P2)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': min(args.warmup, 3), 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'forward+decode 384x1280, 1 frame per step on the host cores (bounded sample of the '
                                   f'batch={args.batch}/GPU workload), random-init DLA-34 + MonoCon heads'},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} steps x 1 frame, oracle port of the reference PyTorch CPU path'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from monocon_pytorch_b200 import engine as E

    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback for the product path'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    if world > 1:          # leave SMs to the NCCL kernel that overlaps the next batch (csrc/common.cuh: reserved_sms)
        os.environ.setdefault('MC_RESERVE_SMS', '0')   # measured: no effect at N = 2 (profiles/README.md)
    sd = synthetic_state_dict()
    eng = E.Engine(dev, B, H, W, args.precision)
    eng.load_state_dict(sd)
    eng.set_option('use_graph', 0 if args.no_graph else 1)

    n_rot = 4                                               # rotate input batches: 4 x 94 MB > L2 (126 MB)
    imgs_host = [synthetic_frames(B, 1000 * rank + i).pin_memory() for i in range(n_rot)]
    imgs = [t.to(dev) for t in imgs_host]
    P2_np = kitti_p2(B)
    P2_h = torch.from_numpy(P2_np)
    invP_h = E.inverse_viewpad(P2_np)
    P2, invP = P2_h.to(dev), invP_h.to(dev)

    # decode outputs live in one flat buffer so that N > 1 needs a single all-gather per batch.  Two output buffers
    # alternate: the all-gather of batch i is issued asynchronously (it runs on NCCL's stream behind the decode of batch
    # i) and is only waited for before batch i + 2 reuses the buffer, so the collective overlaps the next batch's
    # forward instead of serialising the ranks after every step; every batch is still gathered inside the timed region.
    topk = 30
    n = B * topk
    from monocon_pytorch_b200 import dist as mcdist
    packs = [mcdist.alloc_packed(B, topk, dev) for _ in range(2)]
    flat, out = packs[0]
    total = flat.numel()
    gathered = [torch.zeros(world * total, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    works = [None, None]

    # N > 1, default: the all-gather is fused into the decode kernel (peer-memory stores over NVLink, mc_gather_*); if the
    # CUDA IPC mapping is not available on this box the NCCL path (also on the GPU) is used and named in `config`
    pg = None
    gather_impl = 'none'
    if os.environ.get('BENCH_FORCE_PG') == '1' and world == 1:      # diagnostic: the gather path's launches without peers
        pg = mcdist.PeerGather(eng, topk)
        gather_impl = 'p2p(world=1)'
    if world > 1 and not args.no_gather:
        gather_impl = 'nccl'
        if args.gather == 'p2p':
            try:                                          # PeerGather agrees on success / failure across ranks itself
                pg = mcdist.PeerGather(eng, topk)
                gather_impl = 'p2p'
            except Exception as e:                       # noqa: BLE001
                print(f'[bench] peer-memory gather unavailable ({e}); using the NCCL all-gather', file=sys.stderr, flush=True)
                pg = None
    waiting = [False, False]

    def step(i):
        j = i & 1
        if pg is not None:
            if waiting[j]:
                pg.wait(j)                                # the gathered batch i - 2 is complete on this rank
            pg.infer(imgs[i % n_rot], P2, invP, buf=j, thres=0.4)
            waiting[j] = True
            return
        if works[j] is not None:
            works[j].wait()
            works[j] = None
        eng.infer_device(imgs[i % n_rot], P2, invP, topk=topk, thres=0.4, out=packs[j][1])
        if world > 1 and not args.no_gather:
            works[j] = dist.all_gather_into_tensor(gathered[j], packs[j][0], async_op=True)

    def drain():
        for j in range(2):
            if pg is not None and waiting[j]:
                pg.wait(j)
                waiting[j] = False
            if works[j] is not None:
                works[j].wait()
                works[j] = None

    for i in range(Wm):
        step(i)
    drain()
    torch.cuda.synchronize()
    # the NVML sampler starts BEFORE the barrier: initialising it takes ~5 ms on rank 0, and a rank that enters the timed
    # region late makes every other rank's last all-gather wait for it (measured: 0.16 ms per step of apparent overhead)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(K):
        step(i)
    drain()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1 and os.environ.get('BENCH_VERBOSE'):
        print(f'[bench] rank {rank}: {ms_total / K:.4f} ms/step on its own device clock', file=sys.stderr, flush=True)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches = eng.kernel_launches

    # ---- end to end through the host-buffer C-ABI calls ------------------------------------------
    # (a) synchronous call per batch: H2D -> forward -> decode -> D2H, nothing overlapped
    host_out = None
    for i in range(3):
        host_out = eng.infer_host(imgs_host[i % n_rot], P2_h, invP_h, topk=topk, thres=0.4, out=host_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        host_out = eng.infer_host(imgs_host[i % n_rot], P2_h, invP_h, topk=topk, thres=0.4, out=host_out)
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    # (b) the streaming form of the same API: two slots, batch i+1 is submitted (H2D on the copy engine) before the
    #     results of batch i are waited for; every batch still pays its full H2D and D2H inside the timed region,
    #     and at N > 1 the all-gather of every batch's boxes
    outs = [E.Engine.alloc_host_out(B, topk), E.Engine.alloc_host_out(B, topk)]
    gath_h = [torch.empty(world * total, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None

    def e2e_loop(nsteps):
        eng.infer_host_submit(0, imgs_host[0], P2_h, invP_h, outs[0], topk=topk, thres=0.4)
        for i in range(nsteps):
            if i + 1 < nsteps:
                eng.infer_host_submit((i + 1) & 1, imgs_host[(i + 1) % n_rot], P2_h, invP_h, outs[(i + 1) & 1], topk=topk, thres=0.4)
            eng.infer_host_wait(i & 1)
            if world > 1:                      # host results of this batch -> device -> all ranks (same bytes as the device path)
                j = i & 1
                if works[j] is not None:
                    works[j].wait()
                for k in ('box2d', 'box3d', 'labels', 'inds', 'valid'):
                    packs[j][1][k].copy_(outs[j][k], non_blocking=True)
                works[j] = dist.all_gather_into_tensor(gath_h[j], packs[j][0], async_op=True)
        drain()

    e2e_loop(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, e2e_sync_s = float(te[0].item()), float(te[1].item())
    h2d = B * 3 * H * W * 4 + B * 12 * 4 + B * 16 * 4
    d2h = n * (5 * 4 + 7 * 4 + 8 + 8 + 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (rank 0, eager launches, CUDA events per launch) ---
    pk = peaks()
    stages = eng.profile_stages(imgs[0], P2, invP, iters=3)
    conv = [s for s in stages if s['flops'] > 0]
    tc = [s for s in conv if s['tensor_core']]
    dom = tc if tc else conv
    dom_ms = sum(s['ms'] for s in dom)
    dom_flops = sum(s['flops'] for s in dom)
    all_ms = sum(s['ms'] for s in stages)
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r01_traffic.json')       # mean DRAM bytes per convolution launch, from the committed ncu pass
    if os.path.exists(tp) and B == 16 and args.precision == 'bf16':
        tj = json.load(open(tp))
        traffic, traffic_src = tj['dram_bytes_per_launch_avg'], tj['source']
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tf_sustained'], 'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu, read + write)',
                'traffic_source': traffic_src, 'algorithmic_bytes_per_launch_avg': sum(s['bytes'] for s in dom) / max(1, len(dom)),
                'kernel': 'conv_tc (tcgen05 implicit GEMM)' if tc else 'conv_simt (fp32 FFMA implicit GEMM)',
                'launches_per_step': len(dom), 'share_of_step': dom_ms / all_ms if all_ms else None,
                'flops_per_launch_avg': dom_flops / max(1, len(dom)), 'peak_source': pk['src'] + ' sustained bf16',
                'hbm': {'algorithmic_bytes_per_step': eng.bytes_per_image * B,
                        'achieved_gbs': eng.bytes_per_image * B / (ms_total / K * 1e-3) / 1e9, 'peak_gbs': pk['hbm_gbs']}}
    if args.stage_table:
        with open(args.stage_table, 'w') as f:
            f.write(f'# per-stage device time, batch {B}, {args.precision}, CUDA events, eager launches\n')
            f.write('stage,ms,GFLOP,TFLOP/s,MB_algorithmic,GB/s,tensor_core\n')
            for s in stages:
                tf = s['flops'] / (s['ms'] * 1e-3) / 1e12 if s['ms'] > 0 else 0
                gb = s['bytes'] / (s['ms'] * 1e-3) / 1e9 if s['ms'] > 0 else 0
                f.write(f"{s['name']},{s['ms']:.4f},{s['flops'] / 1e9:.3f},{tf:.1f},{s['bytes'] / 1e6:.2f},{gb:.0f},{s['impl']}\n")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_oracle_rate(sd)

    value = world * B * K / (ms_total * 1e-3)
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm,
            'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'batch={B}/GPU forward+decode 384x1280 (BASELINE.json configs[1]; configs[3] sharding at N>1)',
                       'arch': 'DLA-34 + DLAUp + MonoCon heads, reference random init (seed 0)', 'frames': 'randn*0.01 (tie-free recipe)',
                       'global_batch': B * world, 'topk': topk, 'cuda_graph': not args.no_graph,
                       'l2': f'{n_rot} rotating input batches ({n_rot * B * 3 * H * W * 4 / 1e6:.0f} MB) and '
                             f'{eng.workspace_bytes / 1e9:.1f} GB of activations per step: working set >> 126 MB L2',
                       'parallelism': (f'dp{world}: batch sharded, decoded boxes all-gathered every batch, ' + ('fused into the decode kernel over peer memory (NVLink stores), waited for two batches later' if gather_impl == 'p2p' else 'NCCL all-gather, asynchronous, waited for two batches later')) if world > 1 else 'single GPU',
                       'gather': gather_impl},
            'clocks': clocks,
            'e2e': {'value': world * B * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'mc_infer_host_submit / mc_infer_host_wait (pinned host frames in, decoded boxes on the host out; two slots, '
                           'H2D of batch i+1 overlaps the compute of batch i)',
                    'sync_call_value': world * B * K / e2e_sync_s,
                    'sync_call_api': 'mc_infer_host (one blocking call per batch, nothing overlapped)'},
            'gpu_launches': launches * K,
            'roofline': roofline,
            'cpu_baseline': cpu,
            'flops_per_image': eng.flops_per_image,
            'model_tflops': eng.flops_per_image * value / 1e12}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
