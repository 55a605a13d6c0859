"""GPU: the backward kernels (csrc/train_backward.cu) through the product library's mc_bw_* entry points on device memory --
the same cases, references (oracle/backward_oracle.py, float64) and tolerances as the CPU host-shim run
(tests/test_backward_kernels_host.py), plus larger shapes where thousands of threads meet in the atomics.
Named zz so that it runs after every test of the inference path and of the already-validated training pieces: these kernels
were written in a round whose GPU budget was already spent, and this file is their first execution on a device."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import backward_cases as BC   # noqa: E402


@pytest.fixture(scope='module')
def bk():
    import ctypes as C
    from monocon_pytorch_b200 import engine as E
    L = E.load_library()
    L.mc_bw_last_error.restype = C.c_char_p
    L.mc_bw_heads_scratch_bytes.restype = C.c_longlong
    return BC.CudaBackend(L)


@pytest.mark.parametrize('srcC,cout,k,s,p,h,w,pitch', BC.CONV_CASES + [((64, 64), 64, 3, 1, 1, 24, 80, None), ((32,), 64, 3, 2, 1, 48, 160, None)])
def test_conv_wgrad_dgrad(bk, srcC, cout, k, s, p, h, w, pitch):
    BC.conv_case(bk, srcC, cout, k, s, p, h, w, pitch)


@pytest.mark.parametrize('C_,relu,res,affine', BC.BN_CASES)
def test_batchnorm_relu_residual_backward(bk, C_, relu, res, affine):
    BC.bn_case(bk, C_, relu, res, affine)


def test_batchnorm_backward_large(bk):
    BC.bn_case(bk, 64, 1, 1, 1, B=4, h=48, w=160)


def test_colsum(bk):
    BC.colsum_case(bk)
    BC.colsum_case(bk, P_=123457, C_=576)


def test_maxpool_backward_with_ties(bk):
    BC.maxpool_case(bk)
    BC.maxpool_case(bk, B=4, c=64, h=48, w=160)


def test_upsample_backward(bk):
    BC.upsample_case(bk)
    BC.upsample_case(bk, B=4, c=64, h=24, w=80)


@pytest.mark.parametrize('B,h,w', BC.HEAD_CASES + [(4, 24, 80)])
def test_heads_backward(bk, B, h, w):
    BC.heads_case(bk, B, h, w)
