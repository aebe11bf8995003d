"""oracle/prepare.py (MaskRCNN.prepare, mask_rcnn.py:152-176) pinned against cv2.resize --
the call the reference makes -- and against golden vectors of the reference's own method
run verbatim (tests/golden/make_golden.py: prepare_fixture)."""
import os

import cv2
import numpy as np
import pytest

from oracle import prepare as op

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture
def no_ipp():
    had = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)      # OpenCV's own float32 code, not the closed-source IPP path
    yield
    cv2.ipp.setUseIPP(had)


def test_resize_fxfy_bit_exact_against_cv2(no_ipp):
    rs = np.random.RandomState(0)
    for t in range(90):
        H, W = rs.randint(20, 300, 2)
        f = [0.5, 1.0, 2.0, rs.uniform(0.3, 3.0), 800. / min(H, W), 0.25][t % 6]
        img = rs.uniform(0, 255, (3, H, W)).astype(np.float32)
        want = cv2.resize(img.transpose(1, 2, 0), None, fx=f, fy=f).transpose(2, 0, 1)
        got = op.resize_fxfy(img, f, f)
        assert got.shape == want.shape, (H, W, f)
        np.testing.assert_array_equal(got, want, err_msg=str((H, W, f)))


def test_resize_fxfy_anisotropic(no_ipp):
    rs = np.random.RandomState(1)
    img = rs.uniform(0, 255, (3, 57, 91)).astype(np.float32)
    for fx, fy in ((0.5, 0.5), (0.5, 1.5), (2.0, 0.5), (1.3, 0.7)):
        want = cv2.resize(img.transpose(1, 2, 0), None, fx=fx, fy=fy).transpose(2, 0, 1)
        np.testing.assert_array_equal(op.resize_fxfy(img, fx, fy), want)


def test_size_rounds_half_to_even():
    assert op.out_size(80, 103, 0.5, 0.5) == (40, 52)       # 51.5 -> 52
    assert op.out_size(80, 101, 0.5, 0.5) == (40, 50)       # 50.5 -> 50
    assert op.out_size(500, 833, 1.6, 1.6) == (800, 1333)


def test_scale_rule():
    assert op.prepare_scale(500, 833, 800, 1333) == 800 / 500
    assert op.prepare_scale(500, 1000, 800, 1333) == 1333 / 1000
    assert op.prepare_scale(500, 1000, 0, 0) == 1.


def test_golden_reference_prepare():
    """The reference's prepare ran with the wheel's default (IPP on): values agree to the
    5th significant digit (pixel range 0..255), sizes and scales exactly."""
    g = np.load(os.path.join(HERE, 'golden', 'prepare.npz'))
    for k, (H, W, lo, hi) in enumerate(g['cases']):
        out, sizes, scales = op.prepare([g['img_%d' % k]], int(lo), int(hi), g['mean'])
        assert sizes == [(H, W)]
        assert scales[0] == float(g['scale_%d' % k])
        assert out[0].shape == g['out_%d' % k].shape
        np.testing.assert_allclose(out[0], g['out_%d' % k], rtol=0, atol=255 * 1e-5)
