"""GPU check of the fp32-accurate tensor-core mode (MC_PREC_FP32_TC): prints, never asserts.
1. every convolution geometry of tests/test_gpu_parity.py through mc_conv2d in 'fp32' (tensor cores, fp16 hi + lo planes)
   and 'fp32_simt' (FFMA) against float64 conv2d;
2. the full network on the golden fixtures: per-map error, per-level intermediates, top-k agreement, with and without
   scale calibration.
Usage (GPU box): python scripts/check_fp32_tc.py [--skip-conv] [--skip-e2e]"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from monocon_pytorch_b200 import engine as E   # noqa: E402
from oracle import fixtures as FX              # noqa: E402
from oracle import monocon_oracle as O         # noqa: E402

DEV = torch.device('cuda', 0)


def rel_to_max(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


def conv_cases():
    import test_gpu_parity as TP
    for case in TP.CONV_CASES:
        B, Cin, H, W, Cout, k, stride, pad, use_res, relu, split = case
        g = torch.Generator().manual_seed(hash(case) & 0xffff)
        x = torch.randn(B, Cin, H, W, generator=g)
        w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
        scale = 0.5 + torch.rand(Cout, generator=g)
        shift = 0.2 * torch.randn(Cout, generator=g)
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        res = torch.randn(B, Cout, Ho, Wo, generator=g) if use_res else None
        ref = F.conv2d(x.double(), w.double(), None, stride=stride, padding=pad)
        ref = ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
        if res is not None:
            ref = ref + res.double()
        if relu:
            ref = ref.relu()
        line = f'{str(case):62s}'
        for precision in ('fp32', 'fp32_simt'):
            try:
                y = E.conv2d(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), stride=stride, pad=pad,
                             residual=None if res is None else res.to(DEV), relu=relu, split=split, precision=precision).cpu()
                line += f' | {precision}: max {rel_to_max(y.numpy(), ref.numpy()):.2e} l2 {rel_l2(y.numpy(), ref.numpy()):.2e}'
            except Exception as e:   # noqa: BLE001
                line += f' | {precision}: ERROR {str(e)[:120]}'
        print(line, flush=True)


def e2e(size):
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', f'{size}.npz')))
    sd = FX.make_state_dict(0)
    h, w = [int(v) for v in g['hw']]
    img = FX.make_images(2, h, w, seed=int(g['img_seed']))
    ref_maps, inter = None, None
    if size == 'small':
        ref_maps, inter = O.forward(sd, img, return_intermediates=True)
    for precision, calibrate in (('fp32_simt', False), ('fp32', False), ('fp32', True)):
        eng = E.Engine(DEV, 2, h, w, precision)
        eng.load_state_dict(sd)
        if calibrate:
            eng.calibrate_scales(img.to(DEV))
        t0 = time.time()
        out = eng.forward(img.to(DEV))
        torch.cuda.synchronize()
        tag = f'[{size} {precision}{" calibrated" if calibrate else ""}]'
        errs = []
        for k, t in zip(E.PRED_NAMES, out):
            if size == 'small':
                errs.append(rel_to_max(t.cpu().numpy(), g['pred/' + k]))
            else:
                vals = t.cpu().numpy().reshape(-1)[g['pos/' + k]]
                errs.append(float(np.abs(vals - g['val/' + k]).max() / g['mom/' + k][2]))
        print(tag, 'map errors (rel to max):', ' '.join(f'{e:.1e}' for e in errs), flush=True)
        if inter is not None:
            lv = []
            for lvl in range(2, 6):
                lv.append(rel_to_max(eng.debug_tensor(f'backbone.level{lvl}', 2).cpu().numpy(), inter['backbone'][lvl].numpy()))
            lv.append(rel_to_max(eng.debug_tensor('neck.feat', 2).cpu().numpy(), inter['feat'].numpy()))
            print(tag, 'levels 2..5, neck.feat:', ' '.join(f'{e:.1e}' for e in lv), flush=True)
        P2 = torch.from_numpy(np.asarray(g['P2'], dtype=np.float32)).to(DEV)
        invP = E.inverse_viewpad(g['P2']).to(DEV)
        dec = {k: v.cpu().numpy() for k, v in eng.decode(out, P2, invP, (h, w), topk=30, thres=0.4).items()}
        same = bool(np.array_equal(dec['inds'], g['topk/inds'][:, :30]) and np.array_equal(dec['labels'], g['topk/clses'][:, :30]))
        print(tag, 'top-k identical:', same, '| launches', eng.kernel_launches, flush=True)
        if eng.tensor_core_fp32:
            print(tag, 'scale status (max fraction of fp16 range, saturated tensors):', eng.scale_status(), flush=True)
        eng.close()


def layers(size='small'):
    """Tensor-by-tensor distance between the tensor-core fp32 engine and the FFMA fp32 engine, in execution order."""
    import ctypes
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', f'{size}.npz')))
    sd = FX.make_state_dict(0)
    h, w = [int(v) for v in g['hw']]
    img = FX.make_images(2, h, w, seed=int(g['img_seed'])).to(DEV)
    engs = {}
    for precision in ('fp32_simt', 'fp32'):
        eng = E.Engine(DEV, 2, h, w, precision)
        eng.load_state_dict(sd)
        if precision == 'fp32' and '--calibrate' in sys.argv:
            eng.calibrate_scales(img)
        eng.forward(img)
        engs[precision] = eng
    a, b = engs['fp32_simt'], engs['fp32']
    n = int(a.lib.mc_num_stages(a._h))
    for i in range(1, n - 2):
        name = ctypes.create_string_buffer(128)
        fl, by, tc = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        a.lib.mc_stage_info(a._h, i, name, 128, ctypes.byref(fl), ctypes.byref(by), ctypes.byref(tc))
        nm = name.value.decode()
        b.lib.mc_stage_info(b._h, i, name, 128, ctypes.byref(fl), ctypes.byref(by), ctypes.byref(tc))
        try:
            x, y = a.debug_tensor(nm, 2).cpu().numpy(), b.debug_tensor(nm, 2).cpu().numpy()
        except Exception as e:   # noqa: BLE001
            print(f'{nm:40s} {e}')
            continue
        print(f'{nm:40s} impl {tc.value} shape {tuple(x.shape)} | rel-to-max {rel_to_max(y, x):.2e}  rel-L2 {rel_l2(y, x):.2e}  |x|max {np.abs(x).max():.3g} mean {x.mean():.3g}', flush=True)


if __name__ == '__main__':
    if '--layers' in sys.argv:
        layers()
        sys.exit(0)
    if '--skip-conv' not in sys.argv:
        conv_cases()
    if '--skip-e2e' not in sys.argv:
        e2e('small')
        e2e('full')
