"""CPU: the DCN neck plan of the engine (MC_NECK_DCN: offset convolution run as a 32-channel layer, deformable columns, the
(Cout, Cin, 3, 3) weight regrouped for the 1x1 layer over 9 Cin column channels, BatchNorm fold) executed for real through
mc_create_ex -> mc_set_param -> mc_finalize_params -> mc_forward -> mc_debug_tensor on the host stand-in build of api.cu +
engine.cu (tests/host_shim/build_engine.sh; plain-loop launchers), against the oracle's neck output.  The CUDA kernels themselves
are covered by tests/test_gpu_dcn.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import monocon_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, 'host_shim')
LIB = os.path.join(SHIM, '_build', 'libmonocon_host_engine.so')
SRC = os.path.join(HERE, '..', 'monocon_pytorch_b200', 'csrc')
vp = C.c_void_p


@pytest.fixture(scope='module')
def lib():
    deps = [os.path.join(SRC, f) for f in ('api.cu', 'engine.cu', 'engine.h', 'common.cuh', 'train_backward.cu', 'train_backward.h')]
    deps += [os.path.join(SHIM, f) for f in ('host_shim.h', 'host_engine_stubs.cpp', 'build_engine.sh')]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(['sh', os.path.join(SHIM, 'build_engine.sh')], check=True)
    L = C.CDLL(LIB)
    L.mc_last_error.restype = C.c_char_p
    L.mc_last_error.argtypes = [vp]
    L.mc_create_ex.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mc_set_param.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), C.c_int]
    L.mc_finalize_params.argtypes = [vp, C.c_int]
    L.mc_forward.argtypes = [vp, vp, C.c_int, C.POINTER(vp), vp]
    L.mc_debug_tensor_shape.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mc_debug_tensor.argtypes = [vp, C.c_char_p, C.c_int, vp, vp]
    L.mc_destroy.argtypes = [vp]
    return L


def test_dcn_neck_plan_on_the_host(lib):
    torch.set_num_threads(os.cpu_count())
    B, H, W = 1, 64, 128
    sd = FX.make_state_dict(0, use_dcn=True)
    img = FX.make_images(B, H, W, seed=5)

    def ok(rc, h=None):
        assert rc == 0, lib.mc_last_error(h).decode()
    h = vp()
    ok(lib.mc_create_ex(C.byref(h), 0, B, H, W, 1, 1))            # MC_PREC_FP32, MC_NECK_DCN
    keep = []
    for key, val in sd.items():
        if not torch.is_floating_point(val):
            continue
        a = np.ascontiguousarray(val.detach().numpy().astype(np.float32))
        keep.append(a)
        shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
        ok(lib.mc_set_param(h, key.encode(), a.ctypes.data, shape, a.ndim), h)
    # the deformable plan is inference-only
    assert lib.mc_finalize_params(h, 1) != 0 and b'MC_NECK_DCN' in lib.mc_last_error(h)
    ok(lib.mc_finalize_params(h, 0), h)
    x = np.ascontiguousarray(img.numpy().astype(np.float32))
    pred = [np.zeros((B, c, H // 4, W // 4), np.float32) for c in (3, 9, 2, 2, 2, 18, 3, 2, 12, 12)]
    parr = (vp * 10)(*[p.ctypes.data for p in pred])
    ok(lib.mc_forward(h, x.ctypes.data, B, parr, None), h)
    _, inter = O.forward(sd, img, return_intermediates=True)
    ref = inter['feat'].numpy()
    c, hh, ww = C.c_int(), C.c_int(), C.c_int()
    ok(lib.mc_debug_tensor_shape(h, b'neck.feat', C.byref(c), C.byref(hh), C.byref(ww)), h)
    assert (c.value, hh.value, ww.value) == ref.shape[1:]
    got = np.zeros(ref.shape, np.float32)
    ok(lib.mc_debug_tensor(h, b'neck.feat', B, got.ctypes.data, None), h)
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert err <= 1e-4, err
    # one block in isolation: the 27 offset / mask channels (+5 zero) and the 9 * Cin column tensor exist with the planned shapes
    ok(lib.mc_debug_tensor_shape(h, b'neck.ida_0.proj_1.conv_offset', C.byref(c), C.byref(hh), C.byref(ww)), h)
    assert (c.value, hh.value, ww.value) == (32, H // 32, W // 32)
    ok(lib.mc_debug_tensor_shape(h, b'neck.ida_0.proj_1.columns', C.byref(c), C.byref(hh), C.byref(ww)), h)
    assert (c.value, hh.value, ww.value) == (9 * 512, H // 32, W // 32)
    lib.mc_destroy(h)
