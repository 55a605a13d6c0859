"""GPU parity of the DCN neck variant (BASELINE.json north_star: "the DCN variant"; include/monocon_b200.h MC_NECK_DCN).

Operator level: mc_deform_conv2d -- the fused tcgen05 deformable convolution (csrc/dcn_tc.cu) and, with MC_DCN_FUSE=0 or in the FFMA
twin, deformable columns (csrc/dcn.cu) + the 1x1 convolution kernels -- against oracle/dcn_oracle.py in float64 on the same seeded
inputs, and against torchvision's own stored output (tests/golden/dcn.npz).  Model level: Engine(use_dcn=True) against
tests/golden/dcn_model.npz (the unmodified reference detector with DCNv2 packs on torchvision.ops.deform_conv2d in its IDAUp
blocks, tests/golden/gen_dcn_golden.py), every stage of the 12 deformable blocks on the engine's own inputs, and the module
surface.  Tolerances: fp32-accurate modes 1e-5 per operator (3e-5 for the fused kernel at K >= 2304, see the test) and the
north-star 1e-3 end to end with identical top-k; bf16 mode against the bf16-emulating oracle, bounded."""
import os

import numpy as np
import pytest
import torch

from oracle import compare as CMP
from oracle import dcn_oracle as D
from oracle import fixtures as FX
from oracle import monocon_oracle as O

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import engine as E          # noqa: E402

DEV = torch.device('cuda', 0)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
OP_TOL = {'fp32': 1e-5, 'fp32_simt': 1e-5, 'bf16': 6e-3}      # bf16: the output itself is stored as bf16 (2^-9 relative)


class _fuse:
    """MC_DCN_FUSE for the duration of a block: '1' = the fused tcgen05 kernel (csrc/dcn_tc.cu, default where it applies),
    '0' = deformable columns + 1x1 layer (csrc/dcn.cu).  The library reads the variable when a plan is built."""

    def __init__(self, on):
        self.val = '1' if on else '0'

    def __enter__(self):
        self.old = os.environ.get('MC_DCN_FUSE')
        os.environ['MC_DCN_FUSE'] = self.val

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop('MC_DCN_FUSE', None)
        else:
            os.environ['MC_DCN_FUSE'] = self.old


def _run(case, precision, split=1, offset_gain=1.0, fuse=True):
    x, offset, mask, w, b = D.make_case(*case)
    offset = offset * offset_gain
    with _fuse(fuse):
        y = E.deform_conv2d(x.to(DEV), offset.to(DEV), mask.to(DEV), w.to(DEV), b.to(DEV), split=split, precision=precision).cpu()
    if precision == 'bf16':      # the throughput mode stores x, the offset / mask field, the columns and the weights as bf16: the
        q = lambda t: t.float().bfloat16().double()        # checker rounds the same tensors (a bf16 offset of 3 px is 0.008 px off)
        ref = D.deform_conv2d(q(x), q(offset), q(mask), q(w), b.double(), col_round=q)
    else:
        ref = D.deform_conv2d(x.double(), offset.double(), mask.double(), w.double(), b.double())
    return y.numpy(), ref.numpy()


MODES = [('fp32', True), ('fp32', False), ('fp32_simt', False), ('bf16', True), ('bf16', False)]      # (precision, fused kernel)


@pytest.mark.parametrize('precision,fuse', MODES)
@pytest.mark.parametrize('case', D.GPU_CASES)
def test_deform_conv2d_vs_oracle(case, precision, fuse):
    y, ref = _run(case, precision, fuse=fuse)
    err = CMP.rel_to_max(y, ref)
    tol = OP_TOL[precision]
    if precision == 'fp32' and fuse and case[1] >= 256:
        # the fused kernel gathers each tile once and issues hi x w_lo, lo x w_hi, hi x w_hi back to back, so the cross terms meet a
        # full-size accumulator: three truncations per K-step instead of one (measured 2.0e-5 at K = 4608, 7e-6 unfused)
        tol = 3e-5
    assert err < tol, f'{case} {precision} fused={fuse}: {err:.3e}'


@pytest.mark.parametrize('precision,fuse', [m for m in MODES if m[0] != 'fp32_simt'])
def test_deform_conv2d_two_sources(precision, fuse):
    """The node blocks sample torch.cat([layers[i - 1], up]) (dla_neck.py:104) without materialising it: two channel groups."""
    y, ref = _run((2, 128, 16, 24, 64, 6), precision, split=2, fuse=fuse)
    err = CMP.rel_to_max(y, ref)
    assert err < OP_TOL[precision], f'{err:.3e}'


@pytest.mark.parametrize('precision,fuse', [m for m in MODES if m[0] != 'bf16'])
def test_deform_conv2d_border_heavy_offsets(precision, fuse):
    """Offsets of ~6 pixels on a 9 x 13 map: most samples touch or leave the image (the zero rule of bilinear_interpolate)."""
    y, ref = _run((1, 64, 9, 13, 64, 2), precision, offset_gain=3.0, fuse=fuse)
    err = CMP.rel_to_max(y, ref)
    assert err < OP_TOL[precision], f'{err:.3e}'


def test_deform_conv2d_ragged_tiles_fused():
    """Pixel counts that are no multiple of the 128-pixel tile, more tiles than SMs: the fused kernel's flattened-pixel tiling."""
    for case in ((3, 64, 37, 53, 64, 21), (1, 64, 150, 131, 32, 22)):
        y, ref = _run(case, 'fp32', fuse=True)
        assert CMP.rel_to_max(y, ref) < OP_TOL['fp32'], case


def test_deform_conv2d_vs_torchvision_golden():
    """Case 1 of tests/golden/dcn.npz (64 channels) is torchvision's own float64 output."""
    g = np.load(os.path.join(GOLDEN, 'dcn.npz'))
    case = D.GOLDEN_CASES[1]
    y, _ = _run(case, 'fp32')
    assert CMP.rel_to_max(y, g['y1']) < 1e-5


def test_small_channel_counts_run_on_the_ffma_twin():
    """9 C not a multiple of 64: no tensor-core layer takes the columns; the strict fp32 twin does (golden cases 0 and 2)."""
    g = np.load(os.path.join(GOLDEN, 'dcn.npz'))
    for n in (0, 2):
        y, _ = _run(D.GOLDEN_CASES[n], 'fp32_simt')
        assert CMP.rel_to_max(y, g[f'y{n}']) < 1e-5, n


# ---- the detector with the DCN neck ---------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def dcn_sd():
    return FX.make_state_dict(0, use_dcn=True)


@pytest.fixture(scope='module')
def dcn_golden():
    return np.load(os.path.join(GOLDEN, 'dcn_model.npz'))


@pytest.mark.parametrize('precision,fuse', [('fp32', True), ('fp32', False), ('fp32_simt', False)])
def test_dcn_detector_vs_reference_golden(dcn_sd, dcn_golden, precision, fuse):
    g = dcn_golden
    h, w = (int(v) for v in g['hw'])
    with _fuse(fuse):
        eng = E.Engine(DEV, 2, h, w, precision, use_dcn=True)
    eng.load_state_dict(dcn_sd)
    img = FX.make_images(2, h, w, seed=int(g['img_seed'])).to(DEV)
    if eng.tensor_core_fp32:
        eng.calibrate_scales(img)
    out = eng.forward(img)
    for k, t in zip(E.PRED_NAMES, out):
        err = CMP.rel_to_max(t.cpu().numpy(), g['pred/' + k])
        assert err < 1e-3, f'{k}: {err:.3e}'
    P2 = torch.from_numpy(FX.kitti_p2(2, 1)).to(DEV)
    invP = E.inverse_viewpad(FX.kitti_p2(2, 1)).to(DEV)
    dec = eng.decode(out, P2, invP, (h, w), topk=30, thres=0.4)
    assert np.array_equal(dec['inds'].cpu().numpy(), g['topk/inds'][:, :30])        # smallest score gap of the golden: 1.4e-4
    stages = eng.profile_stages(img, P2, invP, iters=1)
    names = [s['name'] for s in stages]
    assert sum(n.endswith('.conv_offset') for n in names) == 12
    assert sum(n.endswith('.columns') for n in names) == (0 if fuse else 12) and sum(s['impl'] == 4 for s in stages) == (12 if fuse else 0)
    if precision == 'fp32':
        assert all(s['impl'] > 0 for s in stages if s['flops'] > 0), 'a convolution of the DCN plan fell back to the FFMA kernel'
    eng.close()


def test_dcn_detector_bf16_vs_emulating_oracle(dcn_sd, dcn_golden):
    g = dcn_golden
    h, w = (int(v) for v in g['hw'])
    eng = E.Engine(DEV, 2, h, w, 'bf16', use_dcn=True)
    eng.load_state_dict(dcn_sd)
    img = FX.make_images(2, h, w, seed=int(g['img_seed']))
    out = eng.forward(img.to(DEV))
    emu = O.forward(dcn_sd, img, emulate_bf16=True)
    worst = max(CMP.rel_l2(t.cpu().numpy(), emu[k].numpy()) for k, t in zip(E.PRED_NAMES, out))
    print(f'bf16 DCN detector: max rel-L2 vs the bf16-emulating oracle {worst:.3e}')
    assert worst < 0.2
    eng.close()


@pytest.mark.parametrize('precision,fuse', MODES)
def test_dcn_blocks_stagewise_pixel_scale_offsets(precision, fuse):
    """Every stage of every deformable block -- offset convolution (27 of 32 channels, the rest zero), columns, 1x1 layer + BatchNorm
    + ReLU -- against torch / the DCN oracle evaluated on the ENGINE's own input tensors, on the gain-1 fixture (offsets of several
    pixels).  End to end that fixture is ill-conditioned (oracle/fixtures.py: make_state_dict), stage by stage it is the strongest
    check of the kernels inside the real plan.  Measured: 6e-6 / 2e-7 / 8e-6 on the tensor cores, 2e-6 / 2e-7 / 3e-6 with FFMA."""
    import torch.nn.functional as F
    sd = FX.make_state_dict(0, use_dcn=True, dcn_offset_gain=1.0)
    B, H, W = 2, 128, 256
    img = FX.make_images(B, H, W, seed=1).to(DEV)
    with _fuse(fuse):
        eng = E.Engine(DEV, B, H, W, precision, use_dcn=True)
    eng.load_state_dict(sd)
    if eng.tensor_core_fp32:
        eng.calibrate_scales(img)
    eng.forward(img)
    tol = (4e-5 if (precision == 'fp32' and fuse) else 2e-5) if precision != 'bf16' else 8e-3     # fused fp32: see test_deform_conv2d_vs_oracle
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    get = lambda n: eng.debug_tensor(n, B).cpu()
    layers = ['backbone.level2', 'backbone.level3', 'backbone.level4', 'backbone.level5']
    checked = 0
    for i in range(3):                                        # DLAUp.forward / IDAUp.forward, dla_neck.py:94-106,136-143
        start = len(layers) - i - 2
        pre = f'neck.ida_{i}'
        for j in range(1, len(layers) - start):
            for name, srcs in ((f'{pre}.proj_{j}', [layers[start + j]]), (f'{pre}.node_{j}', [layers[start + j - 1], f'{pre}.up_{j}.weight'])):
                x = torch.cat([get(s) for s in srcs], 1)
                off = get(name + '.conv_offset')
                off_ref = F.conv2d(x, sd[name + '.conv.conv_offset.weight'], sd[name + '.conv.conv_offset.bias'], padding=1)
                assert rel(off[:, :27], off_ref) < tol and float(off[:, 27:].abs().max()) == 0.0, name
                col = D.deform_columns(x, off[:, :18], torch.sigmoid(off[:, 18:27]))
                if not fuse:
                    got_col = get(name + '.columns')
                    assert rel(got_col, col) < tol, name
                    col = got_col                                  # the 1x1 layer is checked on the engine's own columns
                elif precision == 'bf16':
                    col = col.bfloat16().float()                   # the fused kernel rounds the sampled tile to bf16 in shared memory
                w3 = sd[name + '.conv.weight']
                wk = w3.permute(0, 2, 3, 1).reshape(w3.shape[0], -1, 1, 1)
                bn = name + '.bn1'
                y = F.relu(F.batch_norm(F.conv2d(col, wk), sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'], False, 0.0, 1e-5))
                assert rel(get(name), y) < tol, name
                checked += 1
            layers[start + j] = f'{pre}.node_{j}'
    assert checked == 12
    eng.close()


def test_dcn_plan_refuses_training(dcn_sd):
    eng = E.Engine(DEV, 2, 128, 256, 'bf16', use_dcn=True)
    with pytest.raises(E.EngineError):
        eng.load_state_dict(dcn_sd, training=True)
    eng.close()


class _Calib:
    def __init__(self, p2):
        self.P2 = p2


def test_dcn_module_drop_in(dcn_sd, dcn_golden):
    """MonoConDetector(use_dcn=True): the reference's call surface on the DCN plan (model(data_dict), batch_eval)."""
    import monocon_pytorch_b200 as M
    g = dcn_golden
    h, w = (int(v) for v in g['hw'])
    model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False, precision='fp32', max_batch=2, use_dcn=True)
    model.load_state_dict(dcn_sd, strict=True)
    model = model.to(DEV).eval()
    data = {'img': FX.make_images(2, h, w, seed=int(g['img_seed'])).to(DEV), 'img_metas': {'pad_shape': [(h, w)] * 2},
            'calib': [_Calib(p) for p in FX.kitti_p2(2, 1)]}
    pred = model(data, return_loss=False)
    for k in E.PRED_NAMES:
        assert CMP.rel_to_max(pred[k].cpu().numpy(), g['pred/' + k]) < 1e-3, k
    res = model.batch_eval(data, get_vis_format=True)
    assert len(res) == 2 and all('img_bbox' in r for r in res)
    with pytest.raises(NotImplementedError):
        model.train()
        model(data, return_loss=True)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_dcn_engine_partial_batch_and_refresh(dcn_sd, dcn_golden, precision):
    """An engine planned for 3 images runs 2 (plane offsets of the fp16 storage follow max_batch, tiles follow B), and
    mc_refresh_params repacks new weights -- the deformable layers' included -- into the same buffers."""
    g = dcn_golden
    h, w = (int(v) for v in g['hw'])
    img = FX.make_images(2, h, w, seed=int(g['img_seed']))
    eng = E.Engine(DEV, 3, h, w, precision, use_dcn=True)
    eng.load_state_dict(dcn_sd)
    if eng.tensor_core_fp32:
        eng.calibrate_scales(img.to(DEV))
    out = [t.clone() for t in eng.forward(img.to(DEV))]
    ref = O.forward(dcn_sd, img, emulate_bf16=(precision == 'bf16'))
    if precision == 'fp32':
        for k, t in zip(E.PRED_NAMES, out):
            assert CMP.rel_to_max(t.cpu().numpy(), g['pred/' + k]) < 1e-3, k
    else:
        assert max(CMP.rel_l2(t.cpu().numpy(), ref[k].numpy()) for k, t in zip(E.PRED_NAMES, out)) < 0.2
    # new values for the deformable layers only, then back: the outputs must move and return
    sd2 = dict(dcn_sd)
    for k in dcn_sd:
        if '.conv_offset.bias' in k or ('neck.' in k and k.endswith('.conv.weight')):
            sd2[k] = dcn_sd[k] * 1.25
    eng.refresh_state_dict(sd2)
    moved = [t.clone() for t in eng.forward(img.to(DEV))]
    assert CMP.rel_to_max(moved[2].cpu().numpy(), out[2].cpu().numpy()) > 1e-2
    ref2 = O.forward(sd2, img, emulate_bf16=(precision == 'bf16'))
    if precision == 'fp32':
        for k, t in zip(E.PRED_NAMES, moved):
            assert CMP.rel_to_max(t.cpu().numpy(), ref2[k].numpy()) < 1e-3, k
    eng.refresh_state_dict(dcn_sd)
    back = eng.forward(img.to(DEV))
    for a, b in zip(back, out):
        assert torch.equal(a, b)
    eng.close()
