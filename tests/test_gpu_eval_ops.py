"""GPU parity of the evaluation overlap kernels (csrc/kernels_eval.cu) through the C ABI: against the golden matrices of
the unmodified reference kernels (tests/golden/iou.npz) and against the oracle on a larger seeded set."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import eval_ops as EV       # noqa: E402
from oracle import iou_oracle as IO                    # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'iou.npz')


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


@pytest.mark.parametrize('criterion', [-1, 0, 1, 2])
def test_rotate_iou_matches_reference_golden(gold, criterion):
    got = EV.rotate_iou_gpu_eval(gold['boxes'], gold['qboxes'], criterion)
    ref = gold[f'riou{criterion}']
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5 * max(1.0, float(np.abs(ref).max())))    # sinf / cosf vs the CPU's
    assert np.array_equal(got > 0, ref > 0)                                                         # same set of overlapping pairs


@pytest.mark.parametrize('criterion', [-1, 0, 1])
def test_d3_box_overlap_matches_reference_golden(gold, criterion):
    got = EV.d3_box_overlap(gold['boxes3d'], gold['qboxes3d'], criterion)
    np.testing.assert_allclose(got, gold[f'd3_{criterion}'], rtol=0, atol=1e-5)


def test_rotate_iou_matches_oracle_on_a_larger_set():
    rng = np.random.RandomState(9)
    def bev(n):
        return np.stack([rng.uniform(-10, 10, n), rng.uniform(0, 40, n), rng.uniform(0.5, 5, n), rng.uniform(0.5, 5, n),
                         rng.uniform(-3.2, 3.2, n)], 1).astype(np.float32)
    b, q = bev(70), bev(65)                         # not multiples of the reference's 64-box blocks
    ref = IO.rotate_iou(b, q, -1)
    got = EV.rotate_iou_gpu_eval(b, q, -1)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5)
    assert EV.bev_box_overlap(b[:3], q[:2]).shape == (3, 2)
    assert EV.rotate_iou_gpu_eval(b[:0], q).shape == (0, 65) and EV.rotate_iou_gpu_eval(b, q[:0]).shape == (70, 0)
