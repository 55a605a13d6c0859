"""GPU tests of the bf16 tensor-core training kernels (run with ``-m gpu`` on a B200), through the C ABI.

Kernel level: the weight gradient (csrc/wgrad_tc.cu) against float64 autograd of torch.nn.functional.conv2d on identical
bf16-rounded operands (the kernel multiplies bf16 x bf16 exactly and accumulates in fp32, so the only differences are the
accumulation order and the tensor core's accumulate rounding).
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from monocon_pytorch_b200 import engine as E          # noqa: E402

DEV = torch.device('cuda', 0)

WGRAD_CASES = [
    # B, Cin, H, W, Cout, k, split
    (2, 64, 16, 16, 64, 3, 1),          # one 64-channel chunk, one Cout tile with 64 real rows (SWIZZLE_128B both operands)
    (2, 64, 24, 40, 128, 3, 1),         # two dy boxes per step (M = 128), H = 24 -> 12-row tiles
    (3, 128, 12, 40, 128, 3, 2),        # two sources (IDAUp node, dla_neck.py:104), odd batch
    (2, 256, 12, 24, 512, 3, 1),        # level5 conv1 shape class: 4 Cout tiles x 4 chunks x 2 tap groups
    (1, 64, 24, 80, 576, 3, 1),         # nine head stems as one convolution: last Cout tile has 64 rows
    (2, 32, 32, 64, 64, 3, 1),          # 32 input channels: SWIZZLE_64B halo tile, all nine taps in one group
    (2, 16, 32, 64, 16, 3, 1),          # level0: 16 -> 16, SWIZZLE_32B on both sides
    (2, 16, 32, 64, 32, 3, 1),          # level1 (as the stride-1 problem on the zero-inserted gradient)
    (2, 32, 18, 40, 64, 3, 1),          # H = 18: 16-row tiles with a zero-filled remainder
    (2, 512, 12, 24, 128, 1, 4),        # Root 1x1 over four children (dla.py:126)
    (2, 128, 24, 40, 64, 1, 2),         # level2 Root
    (2, 32, 24, 40, 64, 1, 1),          # project 1x1 (dla.py:181-185)
    (5, 128, 48, 160, 128, 3, 1),       # more pixel tiles than SMs x stages: the ring wraps, split-K over many CTAs
    (2, 3, 32, 64, 16, 7, 1),           # stem 7x7 over the padded 8-channel image: one N = 64 MMA per filter row (dla.py:231-234)
    (3, 3, 48, 40, 16, 7, 1),
    (2, 64, 5, 12, 64, 3, 1),           # odd height, width not a multiple of 8: tiles stick out of the image (zero fill)
    (2, 256, 2, 4, 512, 3, 1),          # level5 of a 64 x 128 frame
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_wgrad_tc_kernel_parity(case):
    B, Cin, H, W, Cout, k, split = case
    g = torch.Generator().manual_seed(hash(case) & 0xffff)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    dy = torch.randn(B, Cout, H, W, generator=g).bfloat16().float()
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, stride=1, padding=(k - 1) // 2).backward(dy.double())
    ref = w.grad
    got = E.conv2d_wgrad_tc(x.to(DEV), dy.to(DEV), k, split=split).cpu().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    # per-tap diagnostics make a wrong descriptor interpretation readable in the log
    if err >= 1e-4:
        per_tap = [(float((got[:, :, i // k, i % k] - ref[:, :, i // k, i % k]).abs().max() / ref.abs().max())) for i in range(k * k)]
        per_co = (got - ref).abs().amax(dim=(1, 2, 3)) / ref.abs().max()
        per_ci = (got - ref).abs().amax(dim=(0, 2, 3)) / ref.abs().max()
        print(f'{case}: per-tap {per_tap}\n bad co {torch.nonzero(per_co > 1e-4).flatten().tolist()[:40]}\n bad ci {torch.nonzero(per_ci > 1e-4).flatten().tolist()[:40]}')
    assert err < 1e-4, f'{case}: rel-to-max error {err:.3e}'


# ------------------------------------------------------------------------------------------------------------------------
# The whole bf16 tensor-core training step, replayed op by op on the DEVICE's own stored tensors
# ------------------------------------------------------------------------------------------------------------------------
def _replay_setup(B, H, W, sd, head_fast=True):
    import ctypes as C
    import numpy as np
    import backward_cases as BC
    from monocon_pytorch_b200 import dist as mcdist
    from monocon_pytorch_b200 import train_ops as T
    from oracle import fixtures as FX
    from oracle import train_fixtures as TF
    img = FX.make_images(B, H, W, seed=41)
    label = TF.make_labels(B, (H, W), seed=42)
    eng = E.Engine(DEV, B, H, W, 'bf16')
    eng.load_state_dict(sd, training=2)
    eng.set_option('train_debug', 1)                   # keep the fp32 gradient of the head stems for the dumps
    eng.set_option('head_backward', 1 if head_fast else 0)
    pred = eng.forward_train(img.to(DEV))
    data = {'img': img.to(DEV), 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(DEV) for k, v in label.items()}}
    tgt = T.TargetGenerator()(data, (B, 64, H // 4, W // 4))
    loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True)
    dpred = [grad[k].contiguous() for k in E.PRED_NAMES]
    eng.backward_train(pred, dpred)
    torch.cuda.synchronize()
    lib = eng.lib
    tp, op_p, nt, nops = C.POINTER(BC.Tensor)(), C.POINTER(BC.Op)(), C.c_int(), C.c_int()
    lib.mc_debug_bw_graph.argtypes = [C.c_void_p, C.POINTER(C.POINTER(BC.Tensor)), C.POINTER(C.c_int), C.POINTER(C.POINTER(BC.Op)), C.POINTER(C.c_int)]
    assert lib.mc_debug_bw_graph(eng._h, C.byref(tp), C.byref(nt), C.byref(op_p), C.byref(nops)) == 0
    lib.mc_debug_train_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def dump(kind, index, shape):
        out = torch.empty((B,) + tuple(shape), dtype=torch.float32, device=DEV)
        rc = lib.mc_debug_train_dump(eng._h, kind, index, B, out.data_ptr(), None)
        assert rc == 0, (kind, index, lib.mc_last_error(eng._h))
        torch.cuda.synchronize()
        return out.double()

    def dev_f32(ptr, n):
        addr = C.cast(ptr, C.c_void_p).value
        return mcdist._wrap_device_bytes(addr, int(n) * 4, DEV).view(torch.float32).clone().double()

    return eng, tp, nt.value, op_p, nops.value, dump, dev_f32, pred, dpred, BC


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_bf16_training_backward_replayed_on_the_device_tensors(fixture_sd):
    """Parity gate of the bf16 tensor-core training step (csrc/train_engine_tc.cu).  A network's gradients amplify any difference
    of the forward (tests/test_gpu_zz_train_backward.py), so the pass is checked op by op on the tensors the DEVICE itself stored:
    for every stage, the inputs of its backward (gradient of the output, stored activations, raw convolution output, batch
    statistics, master weights rounded to bf16 as the plans hold them) are read back and the formulas of oracle/backward_oracle.py
    are evaluated with torch in float64; the engine's outputs -- gradient of the raw output (dense or zero-inserted), the weight
    gradient in the master layout, BatchNorm weight / bias gradients, and, accumulated over all consumers, the gradient of every
    activation tensor -- must agree to bf16 storage rounding (2^-9) / fp32 accumulation."""
    B, H, W = 2, 128, 256
    eng, tp, nt, op_p, nops, dump, dev_f32, pred, dpred, BC = _replay_setup(B, H, W, fixture_sd)
    T_ = [tp[i] for i in range(nt)]
    shape = lambda t: (t.C, t.H, t.W)
    gexp = {}                                                   # tensor index -> expected gradient (sum over consumers), float64
    worst = {}

    def note(what, err, tol):
        worst[what.split(' ')[0]] = max(worst.get(what.split(' ')[0], 0.0), err)
        assert err <= tol, (what, err)

    def add(t, val):
        gexp[t] = val if t not in gexp else gexp[t] + val

    n_conv = 0
    for i in range(nops - 1, -1, -1):
        op = op_p[i]
        if op.type == BC.HEADS:
            t = op.src[0]
            gexp[t] = dump(1, t, shape(T_[t]))                   # the head backward is the fp32 kernel set validated elsewhere
            continue
        if op.type == BC.POOL:
            s, d = op.src[0], op.dst
            x = dump(0, s, shape(T_[s])).requires_grad_(True)
            gd = dump(1, d, shape(T_[d]))
            note(f'pool-in g[{d}]', _rel(gd, gexp[d]), 2e-2)
            F.max_pool2d(x, 2, 2).backward(gd)
            add(s, x.grad)
            continue
        if op.type == BC.UP:
            s, d = op.src[0], op.dst
            Cs = T_[s].C
            x = dump(0, s, shape(T_[s])).requires_grad_(True)
            w = dev_f32(op.w, Cs * 16).reshape(Cs, 1, 4, 4).requires_grad_(True)
            gd = dump(1, d, shape(T_[d]))
            note(f'up-in g[{d}]', _rel(gd, gexp[d]), 2e-2)
            F.conv_transpose2d(x, w, None, stride=2, padding=1, groups=Cs).backward(gd)
            add(s, x.grad)
            note(f'up-dw stage {i}', _rel(dev_f32(op.dw, Cs * 16).reshape(Cs, 1, 4, 4), w.grad), 2e-3)
            continue
        n_conv += 1
        d = op.dst
        srcs = [op.src[s] for s in range(op.nsrc)]
        cin = sum(T_[s].C for s in srcs)
        k, st, pad, cout = op.k, op.stride, op.pad, op.cout
        gy = dump(1, d, shape(T_[d]))
        note(f'conv-in g[{d}] stage {i}', _rel(gy, gexp[d]), 2e-2)     # the accumulated gradient of this stage's output
        draw_dev = dump(3, i, (cout, T_[d].H * st, T_[d].W * st))
        if op.has_bn:
            y = dump(0, d, shape(T_[d]))
            raw = dump(2, i, shape(T_[d]))
            mean, inv = dev_f32(op.mean, cout), dev_f32(op.inv, cout)
            gamma = dev_f32(op.gamma, cout)
            dz = torch.where(y > 0, gy, torch.zeros_like(gy)) if op.relu else gy
            xh = (raw - mean[None, :, None, None]) * inv[None, :, None, None]
            P = B * T_[d].H * T_[d].W
            s0, s1 = dz.sum((0, 2, 3)), (dz * xh).sum((0, 2, 3))
            draw = (gamma * inv)[None, :, None, None] * (dz - (s0[None, :, None, None] + xh * s1[None, :, None, None]) / P)
            note(f'bn-dgamma stage {i}', _rel(dev_f32(op.dgamma, cout), s1), 2e-3)
            note(f'bn-dbeta stage {i}', _rel(dev_f32(op.dbeta, cout), s0), 2e-3)
            if op.residual >= 0:
                add(op.residual, dz)
        else:
            draw = gy
            note(f'bias stage {i}', _rel(dev_f32(op.dbias, cout), gy.sum((0, 2, 3))), 2e-3)
        if st == 2:                                             # zero-inserted at input resolution
            dense = draw_dev[:, :, ::2, ::2]
            odd = draw_dev.clone()
            odd[:, :, ::2, ::2] = 0
            assert float(odd.abs().max()) == 0.0, f'stage {i}: odd rows / columns of the zero-inserted gradient'
        else:
            dense = draw_dev
        note(f'draw stage {i}', _rel(dense, draw), 1e-2)
        # from here on the DEVICE's bf16 gradient of the raw output is the operand, as in the kernels
        w_m = dev_f32(op.w, k * k * cin * cout).reshape(k * k, cin, cout)
        w_oihw = w_m.permute(2, 1, 0).reshape(cout, cin, k, k)
        wq = w_oihw.float().bfloat16().double().requires_grad_(True)
        is_stem = (k == 7)
        if is_stem:                                             # the padded image is not dumpable: 3 colour channels from the frames
            from oracle import fixtures as FX
            x_in = FX.make_images(B, H, W, seed=41).to(DEV).bfloat16().double()
            xcat = torch.cat([x_in, torch.zeros(B, cin - 3, H, W, dtype=torch.float64, device=DEV)], 1).requires_grad_(True)
        else:
            xcat = torch.cat([dump(0, s, shape(T_[s])) for s in srcs], 1).requires_grad_(True)
        F.conv2d(xcat, wq, None, stride=st, padding=pad).backward(dense)
        dw_ref = wq.grad.reshape(cout, cin, k * k).permute(2, 1, 0)
        note(f'wgrad stage {i}', _rel(dev_f32(op.dw, k * k * cin * cout).reshape(k * k, cin, cout), dw_ref), 2e-3)
        if not is_stem:
            c0 = 0
            for s in srcs:
                add(s, xcat.grad[:, c0:c0 + T_[s].C])
                c0 += T_[s].C
    assert n_conv == 50
    print('bf16 training replay, worst relative errors:', {k: f'{v:.2e}' for k, v in sorted(worst.items())})
    eng.close()


def test_bf16_heads_backward_restructured_equals_fp32_twin(fixture_sd):
    """csrc/train_tc_head.cu (thread = 4 channels x pixel lane, one CTA per stem for the mixture algebra, bf16 stem gradient, stem-bias
    gradient assembled from the sums) against the fp32 kernel set of csrc/train_backward.cu -- itself pinned to the reference's
    gradients (tests/test_backward_oracle.py, test_gpu_zz_train_backward.py) -- on the same forward: every head parameter gradient,
    the gradient of the pre-norm stems and a weight gradient downstream of it."""
    B, H, W = 2, 128, 256
    out = []
    for fast in (False, True):
        eng, tp, nt, op_p, nops, dump, dev_f32, pred, dpred, BC = _replay_setup(B, H, W, fixture_sd, head_fast=fast)
        stems_t = [op_p[i].src[0] for i in range(nops) if op_p[i].type == BC.HEADS][0]
        grads = {k: eng.get_grad(k, v.shape).double() for k, v in fixture_sd.items()
                 if k.startswith('head.') and torch.is_floating_point(v) and 'running_' not in k}
        grads['neck.ida_2.node_3.conv.weight'] = eng.get_grad('neck.ida_2.node_3.conv.weight', fixture_sd['neck.ida_2.node_3.conv.weight'].shape).double()
        out.append((grads, dump(1, stems_t, (tp[stems_t].C, tp[stems_t].H, tp[stems_t].W)).cpu()))
        eng.close()
    (g0, d0), (g1, d1) = out
    assert len(g0) > 80
    assert _rel(d1, d0) < 1e-4                                            # gradient of the stems (fp32 copy of the debug mode)
    worst = ('', 0.0)
    for k, a in g0.items():
        b = g1[k]
        err = float((a - b).norm() / a.norm().clamp_min(1e-30))
        if err > worst[1]:
            worst = (k, err)
        # stem biases / attention 1x1: what is left after the normalisation cancelled everything else (rounding-noise dominated in
        # fp32 on one device already, tests/test_backward_oracle.py); the new path sums them analytically instead of over bf16 values
        cancel = k.endswith(('.0.bias', 'attention.0.weight', 'attention.1.weight', 'attention.1.bias'))
        assert err < (5e-2 if cancel else 2e-3), (k, err)
    print('heads backward, restructured vs fp32 twin: worst', worst)


def test_module_training_iterations_on_the_bf16_engine(fixture_sd):
    """Drop-in surface (engine/monocon_engine.py:80-100) on the tensor-core step: ``model.train_precision = 'bf16'``,
    ``pred, loss = model(data); sum(loss.values()).backward(); optimizer.step()`` for a few iterations -- param.grad as the
    reference leaves it (None on the six dead tensors), the engine refreshed IN PLACE from the module's stepped parameters
    (mc_refresh_params: same buffers, the next forward re-derives every bf16 plan from the new fp32 masters), loss going down."""
    import monocon_pytorch_b200 as M
    from monocon_pytorch_b200 import train_ops as T
    from oracle import fixtures as FX
    from oracle import train_fixtures as TF
    B, H, W = 2, 128, 256
    img = FX.make_images(B, H, W, seed=31)
    label = TF.make_labels(B, (H, W), seed=32)
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    model.load_state_dict(fixture_sd, strict=True)
    model = model.to(DEV).train()
    model.experimental_backward = True
    model.train_precision = 'bf16'
    data = {'img': img.to(DEV), 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(DEV) for k, v in label.items()}}
    opt = T.ClipAdamW([p for p in model.parameters()], lr=2.25e-4, betas=(0.95, 0.99), weight_decay=1e-5, max_norm=35.0)
    totals = []
    for it in range(4):
        pred, loss = model(data)
        total = sum(loss.values())
        totals.append(float(total.detach()))
        total.backward()
        if it == 0:
            n_grad = 0
            for name, p in model.named_parameters():
                if name.startswith(('backbone.level3.project.', 'backbone.level4.project.')):
                    assert p.grad is None, name
                else:
                    assert p.grad is not None and torch.isfinite(p.grad).all(), name
                    n_grad += 1
            assert n_grad == 236
        opt.step()
        opt.zero_grad()
    eng = [e for k, e in model._engines.items() if k[-1] == 'train'][0]
    assert eng.precision == 'bf16'
    assert all(t == t for t in totals) and totals[-1] < 0.8 * totals[0], totals         # the first AdamW steps move every weight by ~lr
    opt.close()
