"""DCN neck variant (MC_NECK_DCN) at the bench shape: parity against the oracle at B = 2, throughput of forward + decode at B = 16
(CUDA-graph replay, CUDA events on the launching stream, 4 rotating input batches > L2) and the per-stage split of one eager pass.

    python scripts/bench_dcn.py [--steps 20] [--warmup 5] > gpurun_out/dcn_bench.json

One JSON line per precision mode.  `columns` is the roofline entry of the new kernel (dcn_columns_kernel, HBM-bound): algorithmic
bytes = (Cin + 32 + 9 Cin) x bytes per element per pixel, summed over the 12 deformable blocks, over the summed per-launch time
(only the unfused plan, MC_DCN_FUSE=0, has that kernel; the default plan runs the fused dcn_tc_kernel).  bench.py calls main() for
its `dcn_variant` sub-line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocon_pytorch_b200 import engine as E      # noqa: E402
from oracle import compare as CMP                 # noqa: E402  (checker only)
from oracle import fixtures as FX                 # noqa: E402
from oracle import monocon_oracle as O            # noqa: E402


def main(a=None, return_lines=False, dev=None):
    if a is None:
        ap = argparse.ArgumentParser()
        ap.add_argument('--steps', type=int, default=20)
        ap.add_argument('--warmup', type=int, default=5)
        ap.add_argument('--batch', type=int, default=16)
        a = ap.parse_args()
    dev = dev if dev is not None else torch.device('cuda', 0)
    lines = []
    H, W, B = 384, 1280, a.batch
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    sd = FX.make_state_dict(0, use_dcn=True)
    small = FX.make_images(2, H, W, seed=3)
    ref = O.forward(sd, small)
    emu = O.forward(sd, small, emulate_bf16=True)
    imgs = [FX.make_images(B, H, W, seed=100 + i).to(dev) for i in range(4)]
    P2h = FX.kitti_p2(B, 7)
    P2, invP = torch.from_numpy(P2h).to(dev), E.inverse_viewpad(P2h).to(dev)
    for precision in ('fp32', 'bf16'):
        eng = E.Engine(dev, B, H, W, precision, use_dcn=True)
        eng.load_state_dict(sd)
        if eng.tensor_core_fp32:
            eng.calibrate_scales(imgs[0])
        out = eng.forward(small.to(dev))
        torch.cuda.synchronize()
        against = ref if precision == 'fp32' else emu
        errs = {k: CMP.rel_to_max(t.cpu().numpy(), against[k].numpy()) for k, t in zip(E.PRED_NAMES, out)}
        eng.set_option('use_graph', 1)
        outs = [eng.alloc_decode(B, 30) for _ in range(4)]
        for i in range(max(a.warmup, 4)):
            eng.infer_device(imgs[i % 4], P2, invP, out=outs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            eng.infer_device(imgs[i % 4], P2, invP, out=outs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        stages = eng.profile_stages(imgs[0], P2, invP, iters=3)
        grp = {'conv_offset': 0.0, 'columns': 0.0, 'dcn_gemm': 0.0, 'other': 0.0}
        col_bytes = 0.0
        eb = 2 if precision == 'bf16' else 4            # fp16 hi + lo planes = 4 bytes per logical element
        shapes = {s['name']: s for s in stages}
        for s in stages:
            n = s['name']
            if n.endswith('.conv_offset'):
                grp['conv_offset'] += s['ms']
            elif n.endswith('.columns'):
                grp['columns'] += s['ms']
            elif n.startswith('neck.') and ('.proj_' in n or '.node_' in n) and s['flops'] > 0:
                grp['dcn_gemm'] += s['ms']
                # flops of the 1x1 layer = 2 * pixels * Cout * 9 Cin  ->  9 Cin = flops / (2 * pixels * Cout); bytes via stage info
            else:
                grp['other'] += s['ms']
        # algorithmic column traffic: per block pixels * (Cin + 32 + 9 Cin) * eb
        blocks = [(512, 12, 40), (256, 24, 80), (256, 24, 80), (128, 48, 160), (128, 48, 160), (128, 48, 160),          # proj_j inputs
                  (512, 24, 80), (256, 48, 160), (256, 48, 160), (128, 96, 320), (128, 96, 320), (128, 96, 320)]       # node_j inputs (2 x Cout)
        for cin, h, w in blocks:
            col_bytes += B * h * w * (cin + 32 + 9 * cin) * eb
        hbm = peaks.get('hbm_copy_gbs') or peaks.get('hbm_gbs') or 6536.0
        line = {'metric': 'images/sec fwd+decode at 384x1280, DCN neck variant', 'value': B / ms * 1e3, 'unit': 'images/s', 'ms_per_step': ms,
                'precision_mode': precision, 'batch': B, 'steps': a.steps, 'warmup': max(a.warmup, 4), 'cuda_graph': True,
                'parity': {'checked': 'B = 2 at 384x1280 vs oracle/monocon_oracle.py with the DCNv2 neck (oracle/dcn_oracle.py)' +
                           ('' if precision == 'fp32' else ', bf16-emulating'), 'max_map_error_rel_to_max': max(errs.values()), 'maps': errs},
                'stage_ms': grp, 'neck_stage_ms': {s['name']: round(s['ms'], 4) for s in stages if s['name'].startswith('neck.')}, 'kernel_launches': eng.kernel_launches,
                'columns': {'bound': 'hbm', 'algorithmic_bytes': col_bytes, 'ms': grp['columns'],
                            'achieved': col_bytes / (grp['columns'] * 1e-3) / 1e9 if grp['columns'] > 0 else None, 'peak': hbm, 'unit': 'GB/s'},
                'gflop_per_image': eng.flops_per_image / 1e9}
        if line['columns']['achieved']:
            line['columns']['frac'] = line['columns']['achieved'] / hbm
        lines.append(line)
        if not return_lines:
            print(json.dumps(line), flush=True)
        eng.close()
    return lines


if __name__ == '__main__':
    main()
