"""Pins the NumPy ROIAlign restatement (oracle/roi_align.py) against outputs of
the reference's own code (tests/golden/*.npz, made by tests/golden/make_golden.py)
and, when /root/reference is mounted, against a live run of that code."""
import os

import numpy as np
import pytest

from oracle import ref_loader
from oracle import roi_align as ora

TOL = 1e-5


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize('ratio', [0, 1, 2])
def test_unit_fixture(golden_dir, ratio):
    g = _load(golden_dir, 'roi_align_unit.npz')
    oh, ow, sc = int(g['outh']), int(g['outw']), float(g['spatial_scale'])
    y = ora.roi_align_forward(g['x'], g['rois'], oh, ow, sc, ratio)
    assert y.dtype == np.float32 and y.shape == g['gy'].shape
    np.testing.assert_allclose(y, g['y_r%d' % ratio], atol=TOL, rtol=0)
    gx = ora.roi_align_backward(g['x'].shape, g['rois'], g['gy'], oh, ow, sc, ratio)
    np.testing.assert_allclose(gx, g['gx_r%d' % ratio], atol=TOL, rtol=0)


def test_check_fixture_known_answers(golden_dir):
    g = _load(golden_dir, 'roi_align_check.npz')
    y = ora.roi_align_forward(g['x'], g['rois'], 2, 2, 1.0, 0)
    np.testing.assert_allclose(y, g['y'], atol=TOL, rtol=0)
    # values printed by the reference's check script for its three RoIs
    np.testing.assert_allclose(y[0, 0], [[0.49, 0.40], [0.39, 0.525]], atol=1e-6)
    np.testing.assert_allclose(y[1, 0], [[0.4675, 0.2996875], [0.398125, 0.5815625]], atol=1e-6)
    gx = ora.roi_align_backward(g['x'].shape, g['rois'], g['gy'], 2, 2, 1.0, 0)
    np.testing.assert_allclose(gx, g['gx'], atol=TOL, rtol=0)
    # every output bin spreads a unit gradient: total mass = number of bins
    assert abs(gx.sum() - g['gy'].size) < 1e-4


@pytest.mark.parametrize('oh', [7, 14])
@pytest.mark.parametrize('ratio', [0, 2])
def test_random_fixture(golden_dir, oh, ratio):
    g = _load(golden_dir, 'roi_align_random.npz')
    y = ora.roi_align_forward(g['x'], g['rois'], oh, oh, 1. / 16, ratio)
    np.testing.assert_allclose(y, g['y_%d_r%d' % (oh, ratio)], atol=TOL, rtol=0)
    gx = ora.roi_align_backward(g['x'].shape, g['rois'], g['gy_%d' % oh], oh, oh, 1. / 16, ratio)
    np.testing.assert_allclose(gx, g['gx_%d_r%d' % (oh, ratio)], atol=2e-5, rtol=0)


def test_argument_errors():
    x = np.zeros((1, 1, 4, 4), np.float32)
    r = np.zeros((1, 5), np.float32)
    with pytest.raises(TypeError):
        ora.roi_align_forward(x, r, 2.0, 2, 1.0)
    with pytest.raises(TypeError):
        ora.roi_align_forward(x, r, 2, 2, '1')
    with pytest.raises(ValueError):
        ora.roi_align_2d(x, r, 2, 2, 1.0, axes='zz')


def test_axes_yx_is_column_swap():
    rs = np.random.RandomState(3)
    x = rs.standard_normal((1, 2, 9, 11)).astype(np.float32)
    rois_xy = np.array([[0, 1.5, 2.0, 9.0, 7.5]], np.float32)
    rois_yx = rois_xy[:, [0, 2, 1, 4, 3]]
    a = ora.roi_align_2d(x, rois_xy, 3, 4, 1.0, axes='xy')
    b = ora.roi_align_2d(x, rois_yx, 3, 4, 1.0, axes='yx')
    np.testing.assert_array_equal(a, b)


@pytest.mark.skipif(not ref_loader.reference_available(), reason='reference tree not mounted')
def test_live_reference_agrees():
    rs = np.random.RandomState(11)
    x = rs.standard_normal((2, 2, 10, 12)).astype(np.float32)
    rois = np.array([[0, 8, 8, 150, 120], [1, 0, 0, 192, 160], [1, 40, 30, 60, 44]], np.float32)
    for ratio in (0, 2):
        ref = ref_loader.ref_roi_align_forward(x, rois, 7, 7, 1. / 16, ratio)
        got = ora.roi_align_forward(x, rois, 7, 7, 1. / 16, ratio)
        np.testing.assert_allclose(got, ref, atol=TOL, rtol=0)
