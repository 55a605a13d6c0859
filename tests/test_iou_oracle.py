"""CPU: the rotated-IoU oracle (oracle/iou_oracle.py) against the golden matrices produced by the UNMODIFIED reference
kernels under numba's CUDA simulator (tests/golden/gen_iou_golden.py)."""
import os

import numpy as np
import pytest

from oracle import iou_oracle as IO

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'iou.npz')


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


@pytest.mark.parametrize('criterion', [-1, 0, 1, 2])
def test_rotate_iou_matches_reference(gold, criterion):
    got = IO.rotate_iou(gold['boxes'], gold['qboxes'], criterion)
    ref = gold[f'riou{criterion}']
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6 * max(1.0, float(np.abs(ref).max())))
    assert (ref > 0).sum() > 20 and (ref == 0).sum() > 20          # the fixture has overlapping and disjoint pairs


@pytest.mark.parametrize('criterion', [-1, 0, 1])
def test_d3_box_overlap_matches_reference(gold, criterion):
    got = IO.d3_box_overlap(gold['boxes3d'], gold['qboxes3d'], criterion)
    np.testing.assert_allclose(got, gold[f'd3_{criterion}'], rtol=0, atol=2e-6)


def test_reference_quirk_identical_boxes(gold):
    """Identical boxes do NOT give IoU 1 in the reference (duplicate vertices in the fan triangulation); parity keeps that."""
    assert abs(float(gold['riou-1'][0, 0]) - 1 / 3) < 1e-5
    assert abs(float(IO.rotate_iou(gold['boxes'][:1], gold['qboxes'][:1])[0, 0]) - 1 / 3) < 1e-5
