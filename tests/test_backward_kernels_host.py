"""CPU: the backward kernels of monocon_pytorch_b200/csrc/train_backward.cu, compiled as host code by tests/host_shim (each
launch runs the kernel body once per (block, thread), sequentially) and called through the same C entry points the product
library exports (mc_bw_*), against the pinned formulas of oracle/backward_oracle.py evaluated in float64 (tests/backward_cases.py).

What this proves: the arithmetic, the NHWC / concatenated-source / padded-pitch indexing, the "+=" vs "=" contracts and the
launch-geometry arithmetic of every kernel.  What it cannot prove: anything about concurrency on a device (the kernels only meet
in atomicAdd) -- tests/test_gpu_zz_train_backward.py runs the same cases on the product library for that."""
import ctypes as C
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import backward_cases as BC   # noqa: E402

SHIM = os.path.join(HERE, 'host_shim')
LIB = os.path.join(SHIM, '_build', 'libtrain_backward_host.so')
SRC = os.path.join(HERE, '..', 'monocon_pytorch_b200', 'csrc')


@pytest.fixture(scope='module')
def bk():
    deps = [os.path.join(SRC, 'train_backward.cu'), os.path.join(SRC, 'train_backward.h'), os.path.join(SHIM, 'host_shim.h')]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(['sh', os.path.join(SHIM, 'build.sh')], check=True)
    L = C.CDLL(LIB)
    L.mc_bw_last_error.restype = C.c_char_p
    L.mc_bw_heads_scratch_bytes.restype = C.c_longlong
    return BC.HostBackend(L)


@pytest.mark.parametrize('srcC,cout,k,s,p,h,w,pitch', BC.CONV_CASES)
def test_conv_wgrad_dgrad(bk, srcC, cout, k, s, p, h, w, pitch):
    BC.conv_case(bk, srcC, cout, k, s, p, h, w, pitch)


@pytest.mark.parametrize('C_,relu,res,affine', BC.BN_CASES)
def test_batchnorm_relu_residual_backward(bk, C_, relu, res, affine):
    BC.bn_case(bk, C_, relu, res, affine)


def test_colsum(bk):
    BC.colsum_case(bk)


def test_maxpool_backward_with_ties(bk):
    BC.maxpool_case(bk)


def test_upsample_backward(bk):
    BC.upsample_case(bk)


@pytest.mark.parametrize('B,h,w', BC.HEAD_CASES)
def test_heads_backward(bk, B, h, w):
    BC.heads_case(bk, B, h, w)
