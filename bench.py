#!/usr/bin/env python
"""bench.py -- images/sec of the MonoCon forward + decode hot path at 384x1280 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16] [--mode infer|train]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batch of 16 synthetic 384x1280 frames per GPU through forward + decode (BASELINE.json configs[1]); at
N > 1 every rank runs its own shard and the decoded boxes are all-gathered inside the timed region (configs[3]).
Prints ONE JSON line.

Headline mode = `--precision fp32`: the reference's fp32 results (TF32 off, test.py:30-33) on the tensor cores
(MC_PREC_FP32_TC: fp16 hi + lo planes, three tcgen05 MMAs per K-block).  BEFORE anything is timed, the exact CUDA graph that
is timed runs one batch of 16 at 384x1280 and all ten maps + the top-k are compared with the CPU oracle (`parity` in the
line; the run aborts if the 1e-3 map tolerance is missed).  The bf16 throughput mode is measured in the same process and
reported beside it (`bf16_mode`, with its own -- much larger -- distance to the oracle).

* value     device-resident inputs, CUDA-graph replay, CUDA events, max over ranks
* e2e       the same metric through the host-buffer C-ABI calls: pinned host uint8 frames in (what the reference's loader
            holds before its transforms), decoded boxes on the host out, H2D/D2H inside the timed region
* roofline  the dominant kernel family (tcgen05 implicit-GEMM convolutions): ALGORITHMIC conv FLOPs / summed per-launch
            durations measured live with CUDA events, against the measured burst bf16 peak (MEASURED_PEAKS.json)
* cpu_baseline / --impl reference  the UNMODIFIED reference (baseline/_ref, batch of 16, all host threads); the oracle port if
            baseline/_ref is absent
* --mode train  BASELINE.json configs[2] / [4]: one training iteration (forward, targets, losses, backward, clip + AdamW) per
            step, gradient all-reduce over NCCL at N > 1 (a second metric; the driver's default run is --mode infer)
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 384, 1280
METRIC = 'images/sec fwd+decode at 384x1280'
UNIT = 'images/s'
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')
MAP_TOL = 1e-3            # BASELINE.json north_star: 1e-3 relative on the ten maps in the fp32-accurate mode
NEAR_TIE = 2e-4           # top-k: reference scores closer than this may swap on the tensor cores (measured map error ~2e-4)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'])
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per step (default 16; 32 in --mode train)')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16', 'fp32_simt'],
                    help="fp32 = the reference's fp32 results on the tensor cores (headline, parity-gated); bf16 = throughput mode")
    ap.add_argument('--train-precision', default='bf16', choices=['bf16', 'fp32_simt'],
                    help='--mode train: bf16 = the tensor-core training step (BASELINE.json configs[2] names bf16), fp32_simt = its FFMA twin')
    ap.add_argument('--no-secondary', action='store_true', help='skip the second-mode line (bf16_mode)')
    ap.add_argument('--no-train-line', action='store_true', help='skip the training-step sub-line (train_step) of the default run at N = 1')
    ap.add_argument('--no-dcn-line', action='store_true', help='skip the DCN-neck-variant sub-line (dcn_variant) of the default run at N = 1')
    ap.add_argument('--no-parity', action='store_true', help='diagnostic only: skip the oracle comparison (not a valid bench line)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-gather', action='store_true', help='diagnostic only: skip the all-gather at N > 1 (not a valid bench line)')
    ap.add_argument('--gather', default='p2p', choices=['p2p', 'nccl'],
                    help='N > 1: p2p = all-gather fused into the decode kernel over peer memory (mc_gather_*), nccl = torch.distributed')
    ap.add_argument('--min-seconds', type=float, default=0.0, help='opt-in: repeat the timed loop until it has run this long (sustained clocks)')
    ap.add_argument('--stage-table', default='', help='write the per-stage timing table to this file')
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 32 if a.mode == 'train' else 16
    return a


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'src': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback (B200_PROFILING.md)'}


def synthetic_state_dict():
    """Random-init weights of the reference architecture (its own init distributions), seed 0."""
    import torch
    import monocon_pytorch_b200 as M
    torch.manual_seed(0)
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    return {k: v.clone() for k, v in model.state_dict().items()}


def synthetic_frames(batch, seed):
    """randn * 0.01: the tie-free recipe of SURVEY.md section 8(d) for the reference's random init."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, H, W, generator=g) * 0.01


def synthetic_frames_u8(batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, H, W, 3), generator=g, dtype=torch.uint8)


U8_MEAN, U8_STD = (127.5, 127.5, 127.5), (4250.0, 4250.0, 4250.0)     # uint8 -> about +-0.03: the scale of the fp32 frames above


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref) or, where that copy is absent, the oracle port
# ----------------------------------------------------------------------------------------------------------------------
class _Calib:                                          # the reference's decode reads only .P2 (monocon_heads.py:501,543)
    def __init__(self, p2):
        self.P2 = p2


def reference_runner(batch):
    """Returns (kind, callable): one call = forward + decode of `batch` frames on the host cores with all threads."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img = synthetic_frames(batch, 100)
    P2 = kitti_p2(batch)
    if os.path.exists(os.path.join(REF_DIR, 'model', '__init__.py')):
        sys.path.insert(0, REF_DIR)
        try:
            from model import MonoConDetector as RefDetector           # the UNMODIFIED reference (model/detector/monocon_detector.py)
            torch.manual_seed(0)
            model = RefDetector(num_dla_layers=34, pretrained_backbone=False).eval()
            torch.backends.cuda.matmul.allow_tf32 = False              # test.py:30-33 (no effect on the CPU; kept for fidelity)
            torch.backends.cudnn.allow_tf32 = False
            data = {'img': img, 'img_metas': {'pad_shape': [(H, W)] * batch}, 'calib': [_Calib(p) for p in P2]}

            def run():
                with torch.no_grad():
                    pred = model(data)                                   # monocon_detector.py:53-65
                    return model.head._get_bboxes(data, pred)            # monocon_heads.py:313-329
            return 'reference', run
        finally:
            sys.path.remove(REF_DIR)
    from oracle import monocon_oracle as O
    sd = {k: v.float() for k, v in synthetic_state_dict().items()}

    def run_port():
        return O.forward_and_decode(sd, img, P2)
    return 'port', run_port


def cpu_baseline(batch, seconds_budget=20.0):
    kind, run = reference_runner(batch)
    run()
    times = []
    t_start = time.perf_counter()
    while len(times) < 10 and (not times or time.perf_counter() - t_start < seconds_budget):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    cores = os.cpu_count() or 1
    what = 'the unmodified reference (baseline/_ref: MonoConDetector.forward + head._get_bboxes)' if kind == 'reference' else \
        'oracle port of the reference PyTorch CPU path (baseline/_ref absent)'
    return {'value': batch / med, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': f'{len(times)} x (batch of {batch} frames 384x1280 fp32 forward+decode), median {med * 1e3:.0f} ms per batch, '
                      f'{what}, torch CPU ops, {cores} threads'}


def run_reference(args, rank):
    if rank != 0:
        return
    kind, run = reference_runner(args.batch)
    cores = os.cpu_count() or 1
    Wm = max(1, min(args.warmup, 3))
    for _ in range(Wm):
        run()
    steps = max(1, min(args.steps, 30))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    val = steps * args.batch / dt
    what = 'the unmodified reference from baseline/_ref (MonoConDetector.forward + head._get_bboxes)' if kind == 'reference' else \
        'oracle port of the reference PyTorch CPU path (baseline/_ref absent)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': Wm, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'batch={args.batch} forward+decode 384x1280 per step on the host cores (BASELINE.json configs[1] '
                                   'geometry), random-init DLA-34 + MonoCon heads, randn*0.01 frames', 'implementation': what},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': kind,
                             'sample': f'{steps} steps x batch of {args.batch}, {what}, {cores} threads'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# parity gate: the timed graph against the CPU oracle, at the bench shape
# ----------------------------------------------------------------------------------------------------------------------
def parity_check(eng, sd, img_host, P2_np, topk, precision):
    """One batch through eng.infer_device (CUDA-graph replay when enabled -- the path that is timed) vs the oracle."""
    import numpy as np
    import torch
    from monocon_pytorch_b200 import engine as E
    from oracle import compare as CMP
    from oracle import monocon_oracle as O
    dev = eng.device
    B = img_host.shape[0]
    P2 = torch.from_numpy(P2_np).to(dev)
    invP = E.inverse_viewpad(P2_np).to(dev)
    img = img_host.to(dev)
    out = None
    for _ in range(2):                                    # capture, then one replay: the replay is what gets compared
        out = eng.infer_device(img, P2, invP, topk=topk, thres=0.4)
        torch.cuda.synchronize()
    maps = [t.cpu().numpy() for t in eng.pred_views(B)]
    t0 = time.perf_counter()
    torch.set_num_threads(os.cpu_count() or 1)
    ref_pred = O.forward({k: v.float() for k, v in sd.items()}, img_host)
    ref_np = {k: v.numpy() for k, v in ref_pred.items()}
    ref_dec = O.decode(ref_np, P2_np, (H, W), topk=topk + 1, thres=0.4)
    oracle_s = time.perf_counter() - t0
    errs = {k: CMP.rel_to_max(m, ref_np[k]) for k, m in zip(E.PRED_NAMES, maps)}
    l2 = {k: CMP.rel_l2(m, ref_np[k]) for k, m in zip(E.PRED_NAMES, maps)}
    n_cells = (H // 4) * (W // 4)
    gi, gl = out['inds'].cpu().numpy(), out['labels'].cpu().numpy()
    strict = CMP.topk_matches(gi, gl, ref_dec['inds'], ref_dec['labels'], ref_dec['scores_raw'], n_cells, 0.0)
    tolerant = CMP.topk_matches(gi, gl, ref_dec['inds'], ref_dec['labels'], ref_dec['scores_raw'], n_cells, NEAR_TIE)
    pos, sym = CMP.count_topk_differences(gi, gl, ref_dec['inds'], ref_dec['labels'], n_cells)
    gaps = -np.diff(ref_dec['scores_raw'].astype(np.float64), axis=1)
    return {'checked': f'batch of {B} at {H}x{W} through the timed path (mc_infer_device, CUDA graph replay) vs the CPU oracle '
                       '(oracle/monocon_oracle.py, pinned to the unmodified reference by tests/golden)',
            'precision': precision, 'map_tolerance': MAP_TOL, 'max_map_error_rel_to_max': max(errs.values()),
            'map_errors_rel_to_max': errs, 'max_map_error_rel_l2': max(l2.values()),
            'maps_within_tolerance': bool(max(errs.values()) < MAP_TOL),
            'topk_identical': bool(strict), 'topk_identical_up_to_near_ties': bool(tolerant), 'near_tie': NEAR_TIE,
            'topk_positions_differing': pos, 'topk_set_difference': sym, 'topk_entries': int(gi.size),
            'reference_min_score_gap': float(gaps.min()), 'oracle_seconds': oracle_s}


# ----------------------------------------------------------------------------------------------------------------------
# one precision mode on this rank: engine, parity, timed device loop, e2e loop
# ----------------------------------------------------------------------------------------------------------------------
def run_mode(args, precision, sd, rank, local_rank, world, steps, warmup, with_e2e=True, with_parity=True, with_module=False):
    import torch
    import torch.distributed as dist
    from monocon_pytorch_b200 import dist as mcdist
    from monocon_pytorch_b200 import engine as E
    dev = torch.device('cuda', local_rank)
    B, K, Wm = args.batch, steps, max(warmup, 3)
    topk = 30
    eng = E.Engine(dev, B, H, W, precision)
    eng.load_state_dict(sd)
    n_rot = 4                                               # rotate input batches: 4 x 94 MB > L2 (126 MB)
    imgs_host = [synthetic_frames(B, 1000 * rank + i).pin_memory() for i in range(n_rot)]
    imgs = [t.to(dev) for t in imgs_host]
    if eng.tensor_core_fp32:
        eng.calibrate_scales(imgs[0])                       # per-tensor scales of the fp16 planes, fitted to the first batch
    eng.set_option('use_graph', 0 if args.no_graph else 1)
    P2_np = kitti_p2(B)
    P2_h = torch.from_numpy(P2_np)
    invP_h = E.inverse_viewpad(P2_np)
    P2, invP = P2_h.to(dev), invP_h.to(dev)

    parity = None
    if with_parity and rank == 0 and not args.no_parity:
        parity = parity_check(eng, sd, imgs_host[0], P2_np, topk, precision)
        if precision != 'bf16' and not parity['maps_within_tolerance']:
            print(json.dumps({'error': 'parity gate failed: prediction maps differ from the oracle by more than 1e-3', 'parity': parity}), flush=True)
            raise SystemExit(3)
        if precision != 'bf16' and not parity['topk_identical_up_to_near_ties']:
            print(json.dumps({'error': 'parity gate failed: top-k differs from the oracle beyond near-ties', 'parity': parity}), flush=True)
            raise SystemExit(3)

    # decode outputs live in one flat buffer so that N > 1 needs a single all-gather per batch.  Two output buffers
    # alternate: the all-gather of batch i is asynchronous and is only waited for before batch i + 2 reuses the buffer, so
    # the exchange overlaps the next batch's forward; every batch is still gathered inside the timed region.
    n = B * topk
    packs = [mcdist.alloc_packed(B, topk, dev) for _ in range(2)]
    total = packs[0][0].numel()
    gathered = [torch.zeros(world * total, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    works = [None, None]
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
    # region late makes every other rank's last all-gather wait for it
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 0
    ev0.record()
    t_wall = time.perf_counter()
    while True:
        for i in range(K):
            step(i)
        reps += 1
        if args.min_seconds <= 0 or time.perf_counter() - t_wall >= args.min_seconds or world > 1:
            break
    drain()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    K_timed = K * reps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches = eng.kernel_launches

    # multi-GPU correctness inside the bench: the gathered block of the last batch of every buffer must equal an NCCL
    # all-gather of the same decode outputs
    gather_check = None
    if world > 1 and not args.no_gather:
        ok = True
        for j in range(2):
            if pg is not None:
                pg.infer(imgs[j % n_rot], P2, invP, buf=j, thres=0.4)
                pg.wait(j)
                torch.cuda.synchronize()
                mine = pg.local_packed(j) if hasattr(pg, 'local_packed') else None
                got = pg.gathered_bytes(j) if hasattr(pg, 'gathered_bytes') else None
                if mine is None or got is None:
                    gather_check = 'unavailable (PeerGather lacks the byte views)'
                    ok = None
                    break
                ref = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(ref, mine)
                ok = ok and bool(torch.equal(ref, got))
            else:
                eng.infer_device(imgs[j % n_rot], P2, invP, topk=topk, thres=0.4, out=packs[j][1])
                w = dist.all_gather_into_tensor(gathered[j], packs[j][0], async_op=True)
                w.wait()
                ref = [torch.empty_like(packs[j][0]) for _ in range(world)]
                dist.all_gather(ref, packs[j][0])
                ok = ok and bool(torch.equal(torch.cat(ref), gathered[j]))
        if ok is not None:
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            gather_check = 'ok' if int(flag.item()) == 1 else 'MISMATCH'

    res = {'precision': precision, 'ms_total': ms_total, 'K': K_timed, 'launches': launches, 'clocks': clocks, 'parity': parity,
           'gather_impl': gather_impl, 'gather_check': gather_check, 'workspace_bytes': eng.workspace_bytes,
           'flops_per_image': eng.flops_per_image, 'bytes_per_image': eng.bytes_per_image, 'n_rot': n_rot}

    # ---- end to end through the host-buffer C-ABI calls ------------------------------------------
    if with_e2e:
        # uint8 HWC frames in pinned host memory (what the reference's loader holds before Normalize / Pad / ToTensor)
        eng.set_normalization(U8_MEAN, U8_STD)
        u8_host = [synthetic_frames_u8(B, 2000 * rank + i).pin_memory() for i in range(n_rot)]
        hw_h = torch.tensor([[H, W]] * B, dtype=torch.int32).pin_memory()
        outs = [E.Engine.alloc_host_out(B, topk), E.Engine.alloc_host_out(B, topk)]
        gath_h = [torch.empty(world * total, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None

        def e2e_loop(nsteps, u8):
            def submit(i):
                if u8:
                    eng.infer_host_u8_submit(i & 1, u8_host[i % n_rot], hw_h, P2_h, invP_h, outs[i & 1], topk=topk, thres=0.4)
                else:
                    eng.infer_host_submit(i & 1, imgs_host[i % n_rot], P2_h, invP_h, outs[i & 1], topk=topk, thres=0.4)
            submit(0)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    submit(i + 1)
                eng.infer_host_wait(i & 1)
                if world > 1 and not args.no_gather:   # host results of this batch -> device -> all ranks (same bytes as the device path)
                    j = i & 1
                    if works[j] is not None:
                        works[j].wait()
                    for k in ('box2d', 'box3d', 'labels', 'inds', 'valid'):
                        packs[j][1][k].copy_(outs[j][k], non_blocking=True)
                    works[j] = dist.all_gather_into_tensor(gath_h[j], packs[j][0], async_op=True)
            drain()

        def timed(u8):
            e2e_loop(3, u8)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_loop(K, u8)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        e2e_u8_s = timed(True)
        e2e_f32_s = timed(False)
        # synchronous single call per batch (nothing overlapped)
        host_out = None
        for i in range(2):
            host_out = eng.infer_host(imgs_host[i % n_rot], P2_h, invP_h, topk=topk, thres=0.4, out=host_out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            host_out = eng.infer_host(imgs_host[i % n_rot], P2_h, invP_h, topk=topk, thres=0.4, out=host_out)
        torch.cuda.synchronize()
        e2e_sync_s = time.perf_counter() - t0
        te = torch.tensor([e2e_u8_s, e2e_f32_s, e2e_sync_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        res['e2e'] = {'u8_s': float(te[0].item()), 'f32_s': float(te[1].item()), 'sync_s': float(te[2].item()), 'K': K,
                      'h2d_u8': B * H * W * 3 + B * 2 * 4 + B * 12 * 4 + B * 16 * 4, 'h2d_f32': B * 3 * H * W * 4 + B * 12 * 4 + B * 16 * 4,
                      'd2h': n * (5 * 4 + 7 * 4 + 8 + 8 + 1)}
        if eng.tensor_core_fp32:
            res['scale_status'] = dict(zip(('max_fraction_of_fp16_range', 'saturated_tensors'), eng.scale_status()))
        # the reference's own call surface: MonoConDetector.batch_eval(data_dict) -> KITTI annotation dicts on the host
        # (engine/monocon_engine.py:134-139), one blocking call per batch: pinned fp32 frames in, H2D inside the call
        if with_module:
            import monocon_pytorch_b200 as M
            model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False, precision=precision, max_batch=B)
            model.load_state_dict(sd)
            model = model.to(dev).eval()
            calibs = [_Calib(p) for p in P2_np]
            metas = {'pad_shape': [(H, W)] * B, 'ori_shape': [(H, W)] * B, 'sample_idx': list(range(B))}

            def module_step(i):
                data = {'img': imgs_host[i % n_rot].to(dev, non_blocking=True), 'img_metas': metas, 'calib': calibs}
                return model.batch_eval(data, get_vis_format=False)
            for i in range(3):
                module_step(i)
            model.freeze_engine(True)                 # serving: weights are fixed, skip the per-call change / range checks
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(K):
                out_mod = module_step(i)
            torch.cuda.synchronize()
            tm = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            res['e2e']['module_s'] = float(tm.item())
            res['e2e']['module_boxes_last'] = int(sum(len(a['score']) for a in out_mod['img_bbox']))
            # the same call with a prefetching loader in front of it (what DataLoader(pin_memory=True) + a CUDA prefetcher give the
            # reference's loop): the frames of batch i + 1 are copied on a copy stream while model.batch_eval(batch i) runs
            copy_stream = torch.cuda.Stream(device=dev)
            bufs = [torch.empty_like(imgs[0]) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            free = [torch.cuda.Event() for _ in range(2)]
            for ev in free:
                ev.record(torch.cuda.current_stream(dev))

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[i % 2])
                    bufs[i % 2].copy_(imgs_host[i % n_rot], non_blocking=True)
                    ready[i % 2].record(copy_stream)

            def module_step_pf(i):
                prefetch(i + 1)
                torch.cuda.current_stream(dev).wait_event(ready[i % 2])
                out = model.batch_eval({'img': bufs[i % 2], 'img_metas': metas, 'calib': calibs}, get_vis_format=False)
                free[i % 2].record(torch.cuda.current_stream(dev))
                return out
            prefetch(0)
            for i in range(2):
                module_step_pf(i)
            torch.cuda.synchronize()
            prefetch(0)
            t0 = time.perf_counter()
            for i in range(K):
                out_mod = module_step_pf(i)
            torch.cuda.synchronize()
            tm = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            res['e2e']['module_pf_s'] = float(tm.item())
            for e_ in list(model._engines.values()):
                e_.close()
            model._engines.clear()
    res['eng'] = eng
    res['imgs'] = imgs
    res['P2'], res['invP'] = P2, invP
    return res


def roofline_of(args, res, precision):
    """Dominant kernel family: the tcgen05 convolutions, per-launch CUDA events (eager launches on the bench's stream)."""
    pk = peaks()
    eng, B = res['eng'], args.batch
    stages = eng.profile_stages(res['imgs'][0], res['P2'], res['invP'], iters=3)
    conv = [s for s in stages if s['flops'] > 0]
    tc = [s for s in conv if s['tensor_core']]
    dom = tc if tc else conv
    dom_ms = sum(s['ms'] for s in dom)
    dom_flops = sum(s['flops'] for s in dom)
    all_ms = sum(s['ms'] for s in stages)
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    mma_factor = 3 if precision == 'fp32' else 1
    ms_step = res['ms_total'] / res['K']
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', f'r02_traffic_{precision}.json')   # mean DRAM bytes per convolution launch, committed ncu --set full pass
    if os.path.exists(tp) and B == 16:
        tj = json.load(open(tp))
        traffic, traffic_src = tj['dram_bytes_per_launch_avg'], tj['source']
    bytes_elem = 2 if precision == 'bf16' else 4
    alg_bytes = [s['bytes'] * bytes_elem / 2 for s in dom]            # mc_stage_info counts 2-byte elements
    rl = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tf_burst'],
          'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu --set full, read + write)', 'traffic_source': traffic_src,
          'algorithmic_bytes_per_launch_avg': sum(alg_bytes) / max(1, len(dom)),
          'kernel': 'conv_tc / conv_tc2 / conv_tc3 (tcgen05 implicit GEMM)' if tc else 'conv_simt (fp32 FFMA implicit GEMM)',
          'launches_per_step': len(dom), 'share_of_step': dom_ms / all_ms if all_ms else None,
          'flops_per_launch_avg': dom_flops / max(1, len(dom)),
          'peak_source': pk['src'] + ': burst bf16/fp16 dense (kernels timed one by one); sustained = ' + f"{pk['tf_sustained']:.0f}",
          'frac_of_sustained_peak': achieved / pk['tf_sustained'],
          'note': ('ALGORITHMIC conv FLOPs (2 * Ho * Wo * Cout * k * k * Cin per image).  The fp32-accurate mode issues three fp16 MMAs per '
                   'algorithmic K-block, so the tensor pipe executes mma_flops_executed = 3x that') if mma_factor == 3 else
                  'ALGORITHMIC conv FLOPs (2 * Ho * Wo * Cout * k * k * Cin per image); one bf16 MMA per K-block',
          'mma_flops_executed_tflops': achieved * mma_factor, 'tensor_pipe_frac_executed': achieved * mma_factor / pk['tf_burst'],
          'whole_step': {'tflops_algorithmic': res['flops_per_image'] * B / (ms_step * 1e-3) / 1e12,
                         'frac_of_burst': res['flops_per_image'] * B / (ms_step * 1e-3) / 1e12 / pk['tf_burst']},
          'hbm': {'algorithmic_bytes_per_step': res['bytes_per_image'] * B * bytes_elem / 2,
                  'achieved_gbs': res['bytes_per_image'] * B * bytes_elem / 2 / (ms_step * 1e-3) / 1e9, 'peak_gbs': pk['hbm_gbs']}}
    if args.stage_table:
        path = args.stage_table if precision == args.precision else args.stage_table + '.' + precision
        with open(path, 'w') as f:
            f.write(f'# per-stage device time, batch {B}, {precision}, CUDA events, eager launches\n')
            f.write('stage,ms,GFLOP,TFLOP/s,MB_algorithmic,GB/s,tensor_core\n')
            for s in stages:
                tf = s['flops'] / (s['ms'] * 1e-3) / 1e12 if s['ms'] > 0 else 0
                mb = s['bytes'] * bytes_elem / 2
                gb = mb / (s['ms'] * 1e-3) / 1e9 if s['ms'] > 0 else 0
                f.write(f"{s['name']},{s['ms']:.4f},{s['flops'] / 1e9:.3f},{tf:.1f},{mb / 1e6:.2f},{gb:.0f},{s['impl']}\n")
    return rl


DTYPE_NAME = {'fp32': 'f32 results; fp16 hi+lo operand planes, 3 tcgen05 MMAs per K-block, fp32 accumulate',
              'bf16': 'bf16', 'fp32_simt': 'f32 (FFMA)'}


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.mode == 'train':
        from scripts import bench_train
        bench_train.main(args, rank, local_rank, world)
        return
    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback for the product path'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
        os.environ.setdefault('MC_RESERVE_SMS', '0')   # measured: no effect at N = 2 (profiles/README.md)
    B = args.batch
    sd = synthetic_state_dict()

    res = run_mode(args, args.precision, sd, rank, local_rank, world, args.steps, args.warmup, with_module=True)
    rl = roofline_of(args, res, args.precision) if rank == 0 else None
    res.pop('eng').close()
    res.pop('imgs')
    torch.cuda.empty_cache()

    second = None
    other = None if args.no_secondary else ('bf16' if args.precision != 'bf16' else 'fp32')
    if other is not None:
        r2 = run_mode(args, other, sd, rank, local_rank, world, max(5, args.steps // 2), 3, with_e2e=True)
        rl2 = roofline_of(args, r2, other) if rank == 0 else None
        r2.pop('eng').close()
        r2.pop('imgs')
        if rank == 0:
            v2 = world * B * r2['K'] / (r2['ms_total'] * 1e-3)
            second = {'precision': other, 'dtype': DTYPE_NAME[other], 'value': v2, 'unit': UNIT, 'ms_per_step': r2['ms_total'] / r2['K'],
                      'steps': r2['K'], 'e2e_value': world * B * r2['e2e']['K'] / r2['e2e']['u8_s'], 'parity': r2['parity'],
                      'roofline': {k: rl2[k] for k in ('achieved', 'peak', 'frac', 'unit', 'share_of_step', 'tensor_pipe_frac_executed')},
                      'gpu_launches_per_step': r2['launches'],
                      'note': ('throughput mode: bf16 storage and operands; its distance to the fp32 reference is reported, not gated'
                               if other == 'bf16' else 'the fp32-accurate tensor-core mode (parity-gated)')}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(B)

    # second metric of the repo (BASELINE.json configs[2]): one device-resident bf16 training iteration per step at batch 32, measured
    # in the same process at N = 1 (`--mode train` is the full line, and the data-parallel form at N > 1)
    train_line = None
    if world == 1 and not args.no_train_line and not args.no_secondary:
        try:
            import argparse as _ap
            from scripts import bench_train
            torch.cuda.empty_cache()
            targs = _ap.Namespace(batch=32, steps=max(5, min(args.steps, 10)), warmup=3, train_precision='bf16')
            tl = bench_train.main(targs, 0, local_rank, 1, return_line=True)
            train_line = {k: tl[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'steps', 'warmup', 'dtype', 'phases_ms', 'clocks')}
            train_line.update({'batch': 32, 'e2e_value': tl['e2e']['value'], 'gpu_launches_per_step': tl['gpu_launches'] // tl['steps'],
                               'roofline': {k: tl['roofline'][k] for k in ('bound', 'achieved', 'peak', 'frac', 'unit')},
                               'parity': 'tests/test_gpu_train_tc.py: the pass replayed op by op on the device tensors against float64 formulas '
                                         '(weight gradients 3e-6, activation gradients <= 6e-3 = bf16 storage rounding)'})
        except Exception as ex:                     # noqa: BLE001  (the headline line must not depend on the second metric)
            train_line = {'error': repr(ex)[:300]}

    # the DCN variant of the neck that north_star names, same batch and geometry: its own oracle parity (B = 2) and
    # throughput in both tensor-core modes -- reported beside the headline, never part of it
    dcn_line = None
    if world == 1 and not args.no_secondary and not args.no_dcn_line:
        try:
            import argparse as _ap
            from scripts import bench_dcn
            torch.cuda.empty_cache()
            dl = bench_dcn.main(_ap.Namespace(steps=max(5, min(args.steps, 10)), warmup=3, batch=B), return_lines=True,
                                dev=torch.device('cuda', local_rank))
            dcn_line = {'config': 'MC_NECK_DCN: every IDAUp 3x3 convolution a DCNv2 pack (operator = torchvision.ops.deform_conv2d); fused '
                                  'tcgen05 deformable convolution (csrc/dcn_tc.cu); fixture with conv_offset gain 0.1',
                        'gflop_per_image': dl[0]['gflop_per_image']}
            for d in dl:
                dcn_line[d['precision_mode']] = {'value': d['value'], 'unit': d['unit'], 'ms_per_step': d['ms_per_step'], 'steps': d['steps'],
                                                 'max_map_error_rel_to_max': d['parity']['max_map_error_rel_to_max'],
                                                 'parity_checked': d['parity']['checked'], 'stage_ms': d['stage_ms'],
                                                 'kernel_launches': d['kernel_launches']}
        except Exception as ex:                     # noqa: BLE001  (the headline line must not depend on the variant)
            dcn_line = {'error': repr(ex)[:300]}

    K = res['K']
    ms_total = res['ms_total']
    value = world * B * K / (ms_total * 1e-3)
    e = res['e2e']
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPE_NAME[args.precision], 'data': 'synthetic',
            'config': {'workload': f'batch={B}/GPU forward+decode 384x1280 (BASELINE.json configs[1]; configs[3] sharding at N>1)',
                       'precision_mode': args.precision,
                       'arch': 'DLA-34 + DLAUp + MonoCon heads, reference random init (seed 0)', 'frames': 'randn*0.01 (tie-free recipe)',
                       'global_batch': B * world, 'topk': 30, 'cuda_graph': not args.no_graph,
                       'l2': f"{res['n_rot']} rotating input batches ({res['n_rot'] * B * 3 * H * W * 4 / 1e6:.0f} MB) and "
                             f"{res['workspace_bytes'] / 1e9:.1f} GB of activations per step: working set >> 126 MB L2",
                       'parallelism': (f'dp{world}: batch sharded, decoded boxes all-gathered every batch, ' + ('fused into the decode kernel over peer memory (NVLink stores), waited for two batches later' if res['gather_impl'] == 'p2p' else 'NCCL all-gather, asynchronous, waited for two batches later')) if world > 1 else 'single GPU',
                       'gather': res['gather_impl']},
            'clocks': res['clocks'],
            'parity': res['parity'],
            'gather_check': res['gather_check'],
            'e2e': {'value': world * B * e['K'] / e['u8_s'], 'unit': UNIT, 'h2d_bytes_per_step': e['h2d_u8'], 'd2h_bytes_per_step': e['d2h'],
                    'api': 'mc_infer_host_u8_submit / mc_infer_host_wait: pinned host uint8 HWC frames in (the reference loader\'s format before '
                           'Normalize / Pad / ToTensor, which run inside the input-packing kernel), decoded boxes on the host out; two slots, '
                           'H2D of batch i+1 overlaps the compute of batch i',
                    'fp32_frames_value': world * B * e['K'] / e['f32_s'], 'fp32_frames_h2d_bytes_per_step': e['h2d_f32'],
                    'fp32_frames_api': 'mc_infer_host_submit (pinned host fp32 NCHW frames, already normalised)',
                    'sync_call_value': world * B * e['K'] / e['sync_s'],
                    'sync_call_api': 'mc_infer_host (one blocking call per batch, fp32 frames, nothing overlapped)',
                    'module_value': (world * B * e['K'] / e['module_s']) if e.get('module_s') else None,
                    'module_prefetch_value': (world * B * e['K'] / e['module_pf_s']) if e.get('module_pf_s') else None,
                    'module_prefetch_api': 'the same batch_eval call behind a prefetching loader: the pinned fp32 frames of batch i + 1 are copied '
                                           'on a copy stream while batch i runs (H2D still inside the timed region)',
                    'module_api': 'MonoConDetector.batch_eval(data_dict, get_vis_format=False): the reference\'s call surface -- pinned fp32 frames '
                                  'copied in, forward + decode + KITTI conversion on the device, one read-back, annotation dicts built on the host; '
                                  'one blocking call per batch'},
            'gpu_launches': res['launches'] * K,
            'roofline': rl,
            'cpu_baseline': cpu,
            'bf16_mode' if (second and second['precision'] == 'bf16') else 'second_mode': second,
            'train_step': train_line,
            'dcn_variant': dcn_line,
            'scale_status': res.get('scale_status'),
            'flops_per_image': res['flops_per_image'],
            'model_tflops': res['flops_per_image'] * value / 1e12}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
