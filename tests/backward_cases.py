"""Shared bodies of the backward-kernel checks: the same cases run (a) on the CPU against the host-shim build of
csrc/train_backward.cu (tests/test_backward_kernels_host.py) and (b) on a GPU against the product library
(tests/test_gpu_zz_train_backward.py).  A backend `bk` supplies the library and moves arrays:
    bk.lib                      ctypes library exporting mc_bw_*
    bk.dev(np_array) -> handle  copy to where the kernels run (None stays None)
    bk.ptr(handle, ctype)       pointer argument
    bk.host(handle) -> np array copy back (after a synchronise)
The references are the pinned formulas of oracle/backward_oracle.py evaluated in float64."""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from oracle import backward_oracle as BO
from oracle import monocon_oracle as O

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)


class HostBackend:
    def __init__(self, lib):
        self.lib = lib

    def dev(self, a):
        return a

    def ptr(self, h, t=fp):
        return None if h is None else h.ctypes.data_as(t)

    def host(self, h):
        return h


class CudaBackend:
    def __init__(self, lib):
        self.lib = lib
        self.device = torch.device('cuda', 0)

    def dev(self, a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def ptr(self, h, t=fp):
        return None if h is None else C.cast(C.c_void_p(h.data_ptr()), t)

    def host(self, h):
        torch.cuda.synchronize()
        return h.cpu().numpy()


def ok(bk, rc):
    assert rc == 0, bk.lib.mc_bw_last_error().decode()


def nhwc(t):
    return np.ascontiguousarray(t.permute(0, 2, 3, 1).numpy().astype(np.float32))


def close(got, ref, tol=2e-5, what=''):
    ref = np.asarray(ref, np.float64)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = float(np.abs(np.asarray(got, np.float64) - ref).max()) / scale
    assert err <= tol, (what, err)


R = lambda g, *s: torch.randn(*s, generator=g, dtype=torch.float64)

# ctypes mirrors of mc_bw_tensor / mc_bw_heads_args / mc_bw_op (include/monocon_b200.h)
class Tensor(C.Structure):
    _fields_ = [('x', fp), ('g', fp), ('C', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Wp', C.c_int), ('xoff', C.c_int)]


class HeadsArgs(C.Structure):
    _fields_ = [('pred', fp * 10), ('dpred', fp * 10), ('sums', dp)] + \
               [(n, fp) for n in ('coefA', 'coefB', 'att_w', 'att_gamma', 'att_beta', 'bank_w', 'bank_b', 'w')] + [('scratch', C.c_void_p)] + \
               [(n, fp) for n in ('dw', 'dbias', 'datt_w', 'datt_gamma', 'datt_beta', 'dbank_w', 'dbank_b')]


class Op(C.Structure):
    _fields_ = [('type', C.c_int), ('nsrc', C.c_int), ('src', C.c_int * 4), ('dst', C.c_int), ('residual', C.c_int), ('relu', C.c_int),
                ('k', C.c_int), ('stride', C.c_int), ('pad', C.c_int), ('cout', C.c_int), ('w', fp), ('dw', fp), ('wT', fp), ('dbias', fp), ('has_bn', C.c_int),
                ('raw', fp), ('mean', fp), ('inv', fp), ('gamma', fp), ('dgamma', fp), ('dbeta', fp), ('draw', fp), ('sums', dp),
                ('heads', C.POINTER(HeadsArgs))]


CONV, POOL, UP, HEADS = 0, 1, 2, 3

CONV_CASES = [
    ((4,), 16, 7, 1, 3, 10, 14, (22, 5)),          # stem: 3 colours + zero pad channel, padded row pitch, no input gradient
    ((8, 16), 12, 3, 1, 1, 6, 9, None),            # IDAUp node: two concatenated sources
    ((16,), 24, 3, 2, 1, 8, 10, None),             # stride 2 (level1 / tree1.conv1)
    ((16,), 8, 3, 2, 1, 7, 9, None),               # stride 2 on odd sizes (last row / column never read by the forward)
    ((8, 8, 4, 4), 20, 1, 1, 0, 5, 6, None),       # Root: 1x1 over four sources
    ((6, 3), 10, 3, 1, 1, 5, 7, None),             # channel counts that are not multiples of 4: the scalar kernels
    ((5,), 7, 3, 2, 1, 6, 6, None),
]
BN_CASES = [(16, 1, 1, 1), (64, 1, 0, 1), (24, 0, 0, 1), (576, 0, 0, 0), (8, 1, 1, 1)]
HEAD_CASES = [(3, 6, 10), (2, 4, 4)]


def conv_case(bk, srcC, cout, k, s, p, h, w, pitch, B=2):
    g = torch.Generator().manual_seed(sum(srcC) * 7 + cout)
    cin = sum(srcC)
    x = R(g, B, cin, h, w)
    stem = pitch is not None
    if stem:
        x[:, 3] = 0
    wt = R(g, cout, cin, k, k)
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = R(g, B, cout, oh, ow)
    ref_dw = BO.conv2d_wgrad(x, dy, k, s, p)
    ref_dx = BO.conv2d_dgrad(dy, wt, (h, w), s, p)
    for use_wT in (False, True):                                   # strided and transposed (coalesced, 4 channels per thread) weight reads
        _conv_run(bk, x, wt, dy, ref_dw, ref_dx, srcC, cout, k, s, p, h, w, pitch, B, use_wT)


def _conv_run(bk, x, wt, dy, ref_dw, ref_dx, srcC, cout, k, s, p, h, w, pitch, B, use_wT):
    stem = pitch is not None
    cin = sum(srcC)
    oh, ow = dy.shape[2:]
    srcs, dsrcs, dsrc0, o = [], [], [], 0
    for cs in srcC:
        a = nhwc(x[:, o:o + cs].float())
        if stem:
            Wp, xo = pitch
            phys = np.full((B, h, Wp, cs), 7.0, np.float32)        # garbage in the padding columns must never be read
            phys[:, :, xo:xo + w] = a
            a = phys
        srcs.append(bk.dev(a))
        d0 = np.random.RandomState(o).randn(B, h, w, cs).astype(np.float32)
        dsrc0.append(d0.copy())
        dsrcs.append(None if stem else bk.dev(d0))
        o += cs
    w_simt = bk.dev(np.ascontiguousarray(wt.permute(2, 3, 1, 0).reshape(k * k, cin, cout).numpy().astype(np.float32)))
    dw0 = np.random.RandomState(1).randn(k * k, cin, cout).astype(np.float32)
    dw = bk.dev(dw0.copy())
    dyd = bk.dev(nhwc(dy.float()))
    n = len(srcC)
    arr = lambda xs: (fp * n)(*[bk.ptr(a) for a in xs])
    ia = lambda xs: (C.c_int * n)(*xs)
    wT = bk.dev(np.full((k * k, cout, cin), 7.0, np.float32)) if use_wT else None
    ok(bk, bk.lib.mc_bw_conv(n, arr(srcs), arr(dsrcs), ia(srcC), ia([pitch[0]] * n) if stem else None, ia([pitch[1]] * n) if stem else None,
                             B, h, w, oh, ow, cout, k, s, p, bk.ptr(w_simt), bk.ptr(dyd), bk.ptr(dw), bk.ptr(wT), None))
    close(bk.host(dw) - dw0, ref_dw.permute(2, 3, 1, 0).reshape(k * k, cin, cout).numpy(), what='dw (+=)')
    o = 0
    for cs, d, d0 in zip(srcC, dsrcs, dsrc0):
        if d is not None:
            close(bk.host(d) - d0, ref_dx[:, o:o + cs].permute(0, 2, 3, 1).numpy(), what='dsrc (+=)')
        o += cs


def bn_case(bk, C_, relu, res, affine, B=2, h=5, w=7):
    g = torch.Generator().manual_seed(C_)
    raw = R(g, B, C_, h, w) * 2 + 0.5
    gamma = R(g, C_) * 0.5 + 1 if affine else None
    beta = R(g, C_) if affine else None
    resid = R(g, B, C_, h, w) if res else None
    eps = 1e-5 if affine else 1e-3
    var, mean = torch.var_mean(raw, dim=(0, 2, 3), unbiased=False)
    inv = (var + eps).rsqrt()
    z = F.batch_norm(raw, None, None, gamma, beta, True, 0.1, eps) + (resid if res else 0)
    y = z.clamp_min(0) if relu else z
    dy = R(g, *y.shape)
    dz = dy * (y > 0) if relu else dy
    ref_dx, ref_dg, ref_db = BO.batchnorm_train_backward(raw, dz, gamma, eps)
    f32 = lambda t: None if t is None else np.ascontiguousarray(t.numpy().astype(np.float32))
    n = B * h * w
    draw = bk.dev(np.full((n, C_), 9.0, np.float32))
    dres0 = np.random.RandomState(2).randn(n, C_).astype(np.float32) if res else None
    dres = bk.dev(dres0.copy()) if res else None
    dgam, dbet = bk.dev(np.zeros(C_, np.float32)), bk.dev(np.zeros(C_, np.float32))
    sums = bk.dev(np.zeros(2 * C_, np.float64))
    keep = [bk.dev(nhwc(dy.float())), bk.dev(nhwc(y.float())) if relu else None, bk.dev(nhwc(raw.float())), bk.dev(f32(mean)), bk.dev(f32(inv)),
            bk.dev(f32(gamma))]
    ok(bk, bk.lib.mc_bw_batchnorm(*[bk.ptr(a) for a in keep], C.c_longlong(n), C_, relu, bk.ptr(sums, dp), bk.ptr(draw), bk.ptr(dres),
                                  bk.ptr(dgam) if affine else None, bk.ptr(dbet) if affine else None, None))
    close(bk.host(draw).reshape(B, h, w, C_), ref_dx.permute(0, 2, 3, 1).numpy(), what='draw')
    if res:
        close(bk.host(dres) - dres0, dz.permute(0, 2, 3, 1).reshape(n, C_).numpy(), what='dres (+=)')
    if affine:
        close(bk.host(dgam), ref_dg.numpy(), what='dgamma')
        close(bk.host(dbet), ref_db.numpy(), what='dbeta')


def colsum_case(bk, P_=1000, C_=65):
    x = np.random.RandomState(3).randn(P_, C_).astype(np.float32)
    out, sums, xd = bk.dev(np.zeros(C_, np.float32)), bk.dev(np.zeros(C_, np.float64)), bk.dev(x)
    ok(bk, bk.lib.mc_bw_colsum(bk.ptr(xd), C.c_longlong(P_), C_, bk.ptr(sums, dp), bk.ptr(out), None))
    close(bk.host(out), x.astype(np.float64).sum(0), tol=1e-6)


def maxpool_case(bk, B=2, c=6, h=8, w=12):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, c, h, w, generator=g).clamp_min(0.)       # post-ReLU: plenty of all-zero windows
    x[0, 0, 0:2, 0:2] = 1.5                                      # and a four-way non-zero tie
    dy = torch.randn(B, c, h // 2, w // 2, generator=g)
    dy = torch.sign(dy) * (0.5 + dy.abs())                       # bounded away from zero: the support test below is exact
    ref = BO.maxpool_backward(x, dy, 2)
    dx0 = np.random.RandomState(5).randn(B, h, w, c).astype(np.float32)
    dx, xd, dyd = bk.dev(dx0.copy()), bk.dev(nhwc(x)), bk.dev(nhwc(dy))
    ok(bk, bk.lib.mc_bw_maxpool2(bk.ptr(xd), bk.ptr(dyd), bk.ptr(dx), B, c, h, w, None))
    got = bk.host(dx) - dx0
    assert np.array_equal(np.abs(got) > 0.25, nhwc(ref) != 0)     # the same winner in every window, ties included
    close(got, nhwc(ref), tol=1e-6)


def upsample_case(bk, B=2, c=12, h=5, w=7):
    g = torch.Generator().manual_seed(6)
    x, wt, dy = R(g, B, c, h, w), R(g, c, 1, 4, 4), R(g, B, c, 2 * h, 2 * w)
    ref_dx, ref_dw = BO.upsample2_backward(x, wt, dy)
    dx0 = np.random.RandomState(7).randn(B, h, w, c).astype(np.float32)
    dx, dw = bk.dev(dx0.copy()), bk.dev(np.zeros((c, 4, 4), np.float32))
    keep = [bk.dev(nhwc(x.float())), bk.dev(np.ascontiguousarray(wt.numpy().astype(np.float32).reshape(c, 16))), bk.dev(nhwc(dy.float()))]
    ok(bk, bk.lib.mc_bw_upsample2(*[bk.ptr(a) for a in keep], bk.ptr(dx), bk.ptr(dw), B, c, h, w, None))
    close(bk.host(dx) - dx0, ref_dx.permute(0, 2, 3, 1).numpy(), what='dx (+=)')
    close(bk.host(dw), ref_dw[:, 0].numpy(), what='dw')


def heads_case(bk, B, h, w):
    """dL/dpred -> output transforms -> ten 1x1 convolutions -> ReLU -> nine AttnBatchNorm2d, against the oracle's formulas
    stem by stem (float64), on float32 inputs laid out as the engine holds them."""
    g = torch.Generator().manual_seed(B * 100 + h)
    HW, K, Cs = h * w, 10, 64
    stems = (R(g, B, 576, h, w) * 1.5 + 0.2).float().double()
    att_w = (R(g, 9, K, Cs) * 0.4).float().double()
    att_g, att_b = (R(g, 9, K) * 0.5 + 1).float().double(), (R(g, 9, K) * 2).float().double()      # a1 on both sides of the knees
    bank_w, bank_b = (R(g, 9, K, Cs) * 0.2 + 1).float().double(), (R(g, 9, K, Cs) * 0.2).float().double()
    W = (R(g, 65, Cs) * 0.2).float().double()
    bias = R(g, 65).float().double()
    bias[39] += 1.0
    o0 = [0, 12, 14, 18, 3, 16, 36, 39, 41]
    o1 = [3, 14, 16, 36, 12, 18, 39, 41, 65]
    pred_ch = [3, 9, 2, 2, 2, 18, 3, 2, 12, 12]
    pred_o0 = np.cumsum([0] + pred_ch[:-1]).tolist()
    # ---- forward (oracle formulas) ---------------------------------------------------------------------------------------
    saved, post, raw = [], [], torch.zeros(B, 65, h, w, dtype=torch.float64)
    coefA, coefB = torch.zeros(B, 576, dtype=torch.float64), torch.zeros(B, 576, dtype=torch.float64)
    for s in range(9):
        x = stems[:, s * 64:(s + 1) * 64]
        out, sv = BO.attn_batchnorm_forward(x, att_w[s], att_g[s], att_b[s], bank_w[s], bank_b[s])
        saved.append(sv)
        ps = out.clamp_min(0)
        post.append(ps)
        raw[:, o0[s]:o1[s]] = torch.einsum('oc,bchw->bohw', W[o0[s]:o1[s]], ps) + bias[o0[s]:o1[s], None, None]
        var, mean = torch.var_mean(x, dim=(0, 2, 3), unbiased=False)
        A = sv['wt'] * (var + 1e-3).rsqrt()
        coefA[:, s * 64:(s + 1) * 64] = A
        coefB[:, s * 64:(s + 1) * 64] = (sv['a'] @ bank_b[s]) - A * mean
    predv = raw.clone()
    predv[:, :12] = torch.clamp(torch.sigmoid(raw[:, :12]), 1e-4, 1 - 1e-4)
    predv[:, 39] = 1. / (torch.sigmoid(raw[:, 39]) + O.EPS) - 1.
    names = O.PRED_NAMES
    pred = {n: predv[:, pred_o0[i]:pred_o0[i] + pred_ch[i]].contiguous() for i, n in enumerate(names)}
    rawd = {n: raw[:, pred_o0[i]:pred_o0[i] + pred_ch[i]].contiguous() for i, n in enumerate(names)}
    dpred = {n: R(g, *pred[n].shape) for n in names}
    # ---- backward (oracle formulas) --------------------------------------------------------------------------------------
    draw = torch.cat([BO.pred_grad_to_raw(pred, rawd, dpred)[n] for n in names], 1)
    ref = dict(dstems=torch.zeros_like(stems), dw=torch.zeros(65, Cs, dtype=torch.float64), dbias=draw.sum((0, 2, 3)),
               datt_w=torch.zeros(9, K, Cs, dtype=torch.float64), datt_g=torch.zeros(9, K, dtype=torch.float64),
               datt_b=torch.zeros(9, K, dtype=torch.float64), dbank_w=torch.zeros(9, K, Cs, dtype=torch.float64),
               dbank_b=torch.zeros(9, K, Cs, dtype=torch.float64))
    for s in range(9):
        d = draw[:, o0[s]:o1[s]]
        ref['dw'][o0[s]:o1[s]] = torch.einsum('bohw,bchw->oc', d, post[s])
        dout = torch.einsum('bohw,oc->bchw', d, W[o0[s]:o1[s]]) * (post[s] > 0)
        dx, gr = BO.attn_batchnorm_backward(stems[:, s * 64:(s + 1) * 64], dout, att_w[s], att_g[s], bank_w[s], bank_b[s], saved[s])
        ref['dstems'][:, s * 64:(s + 1) * 64] = dx
        ref['datt_w'][s] = gr['attn.0.weight'].view(K, Cs); ref['datt_g'][s] = gr['attn.1.weight']; ref['datt_b'][s] = gr['attn.1.bias']
        ref['dbank_w'][s] = gr['weight_']; ref['dbank_b'][s] = gr['bias_']
    # ---- the kernels -----------------------------------------------------------------------------------------------------
    f32 = lambda t: np.ascontiguousarray(t.numpy().astype(np.float32))
    st = nhwc(stems.float()).reshape(B, HW, 576)
    sums = np.stack([st.astype(np.float64).sum(1), (st.astype(np.float64) ** 2).sum(1)], -1)           # [B][576][2]
    preds = [bk.dev(f32(pred[n])) for n in names]
    dpreds = [bk.dev(f32(dpred[n])) for n in names]
    out = dict(dstems=np.full((B, HW, 576), 5.0, np.float32), dw=np.full((65, Cs), 5.0, np.float32), dbias=np.zeros(65, np.float32),
               datt_w=np.zeros((9, K, Cs), np.float32), datt_g=np.zeros((9, K), np.float32), datt_b=np.zeros((9, K), np.float32),
               dbank_w=np.zeros((9, K, Cs), np.float32), dbank_b=np.zeros((9, K, Cs), np.float32))
    out = {k_: bk.dev(v) for k_, v in out.items()}
    nbytes = int(bk.lib.mc_bw_heads_scratch_bytes(B, HW))
    scratch = bk.dev(np.zeros(nbytes // 8 + 64, np.float64))       # 8-byte typed: alignment of the doubles carved from it
    base = (bk.ptr(scratch, C.c_void_p).value + 255) // 256 * 256
    keep = [bk.dev(a) for a in (f32(coefA), f32(coefB), f32(att_w), f32(att_g), f32(att_b), f32(bank_w), f32(bank_b), f32(W))]
    std, sumsd = bk.dev(st), bk.dev(sums)
    ok(bk, bk.lib.mc_bw_heads((fp * 10)(*[bk.ptr(a) for a in preds]), (fp * 10)(*[bk.ptr(a) for a in dpreds]), bk.ptr(std), bk.ptr(sumsd, dp),
                              *[bk.ptr(a) for a in keep], B, HW, C.c_void_p(base), bk.ptr(out['dstems']), bk.ptr(out['dw']),
                              bk.ptr(out['dbias']), bk.ptr(out['datt_w']), bk.ptr(out['datt_g']), bk.ptr(out['datt_b']),
                              bk.ptr(out['dbank_w']), bk.ptr(out['dbank_b']), None))
    got = {k_: bk.host(v) for k_, v in out.items()}
    close(got['dbias'], ref['dbias'].numpy(), 1e-4, 'dbias')
    close(got['dw'], ref['dw'].numpy(), 1e-4, 'dw')
    close(got['dbank_w'], ref['dbank_w'].numpy(), 1e-4, 'dbank_w')
    close(got['dbank_b'], ref['dbank_b'].numpy(), 1e-4, 'dbank_b')
    close(got['datt_g'], ref['datt_g'].numpy(), 5e-4, 'datt_gamma')
    close(got['datt_b'], ref['datt_b'].numpy(), 5e-4, 'datt_beta')
    close(got['datt_w'], ref['datt_w'].numpy(), 5e-4, 'datt_w')
    close(got['dstems'].reshape(B, h, w, 576), ref['dstems'].permute(0, 2, 3, 1).numpy(), 1e-4, 'dstems')


def replay_graph(host_lib, tp, nt, op_p, nops, B, read):
    """Copy a backward graph (arrays of mc_bw_tensor / mc_bw_op whose pointers live wherever `read(ptr, n, dtype)` can read them --
    device memory in the GPU test) into host buffers, run the identical pass with the host-shim library, and return
    [(what, value found behind the original pointer, value the replay produced)] for every output of the pass."""
    keep, compare = [], []
    Pp = lambda a, t=fp: None if a is None else a.ctypes.data_as(t)

    def out_buf(what, ptr, n):
        if not ptr:
            return None
        h = np.full(int(n), 3.25, np.float32)
        keep.append(h)
        compare.append((what, read(ptr, n, np.float32), h))
        return Pp(h)

    def in_buf(ptr, n, dtype=np.float32):
        if not ptr:
            return None
        a = read(ptr, n, dtype)
        keep.append(a)
        return Pp(a, dp if dtype == np.float64 else fp)

    used = set()
    for i in range(nops):
        o = op_p[i]
        used.update(o.src[s] for s in range(o.nsrc))
        if o.type != HEADS:
            used.add(o.dst)
        if o.type == CONV and o.residual >= 0:
            used.add(o.residual)
    tensors = (Tensor * nt)()
    for i in range(nt):
        t = tp[i]
        tensors[i].C, tensors[i].H, tensors[i].W, tensors[i].Wp, tensors[i].xoff = t.C, t.H, t.W, t.Wp, t.xoff
        if i in used:
            tensors[i].x = in_buf(t.x, B * t.H * t.Wp * t.C)
            tensors[i].g = out_buf(f'tensor {i} gradient', t.g, B * t.H * t.W * t.C)
    ops = (Op * nops)()
    hargs = HeadsArgs()
    chs = [3, 9, 2, 2, 2, 18, 3, 2, 12, 12]
    for i in range(nops):
        s, d = op_p[i], ops[i]
        for f in ('type', 'nsrc', 'dst', 'residual', 'relu', 'k', 'stride', 'pad', 'cout', 'has_bn'):
            setattr(d, f, getattr(s, f))
        for j in range(4):
            d.src[j] = s.src[j]
        if s.type == CONV:
            cin = sum(tp[s.src[j]].C for j in range(s.nsrc))
            nw = s.k * s.k * cin * s.cout
            P_ = B * tp[s.dst].H * tp[s.dst].W
            d.w = in_buf(s.w, nw)
            d.dw = out_buf(f'op {i} conv dw', s.dw, nw)
            wT = np.zeros(nw, np.float32); keep.append(wT); d.wT = Pp(wT)
            d.dbias = out_buf(f'op {i} dbias', s.dbias, s.cout)
            sums = np.zeros(2 * s.cout, np.float64); keep.append(sums); d.sums = Pp(sums, dp)
            if s.has_bn:
                d.raw, d.mean, d.inv, d.gamma = in_buf(s.raw, P_ * s.cout), in_buf(s.mean, s.cout), in_buf(s.inv, s.cout), in_buf(s.gamma, s.cout)
                d.dgamma, d.dbeta = out_buf(f'op {i} dgamma', s.dgamma, s.cout), out_buf(f'op {i} dbeta', s.dbeta, s.cout)
                draw = np.zeros(P_ * s.cout, np.float32); keep.append(draw); d.draw = Pp(draw)
        elif s.type == UP:
            c = tp[s.src[0]].C
            d.w, d.dw = in_buf(s.w, c * 16), out_buf(f'op {i} upsample dw', s.dw, c * 16)
        elif s.type == HEADS:
            a, HW = s.heads.contents, tp[s.src[0]].H * tp[s.src[0]].W
            for k in range(10):
                hargs.pred[k], hargs.dpred[k] = in_buf(a.pred[k], B * chs[k] * HW), in_buf(a.dpred[k], B * chs[k] * HW)
            hargs.sums = in_buf(a.sums, B * 576 * 2, np.float64)
            for f, n in (('coefA', B * 576), ('coefB', B * 576), ('att_w', 5760), ('att_gamma', 90), ('att_beta', 90), ('bank_w', 5760),
                         ('bank_b', 5760), ('w', 65 * 64)):
                setattr(hargs, f, in_buf(getattr(a, f), n))
            for f, n in (('dw', 65 * 64), ('dbias', 65), ('datt_w', 5760), ('datt_gamma', 90), ('datt_beta', 90), ('dbank_w', 5760), ('dbank_b', 5760)):
                setattr(hargs, f, out_buf(f'heads {f}', getattr(a, f), n))
            scratch = np.zeros(int(host_lib.mc_bw_heads_scratch_bytes(B, HW)) // 8 + 64, np.float64); keep.append(scratch)
            hargs.scratch = (scratch.ctypes.data + 255) // 256 * 256
            d.heads = C.pointer(hargs)
    rc = host_lib.mc_bw_run_graph(tensors, nt, ops, nops, B, None)
    assert rc == 0, host_lib.mc_bw_last_error().decode()
    return compare
