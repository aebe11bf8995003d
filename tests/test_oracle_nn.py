"""oracle/nn.py (the NumPy restatement of the Chainer layers and losses on the hot path)
cross-checked, forward and backward, against torch CPU ops and autograd -- an
implementation nobody here wrote.  chainer itself is absent: this is the
pin these functions have."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nn as onn

RS = np.random.RandomState


def t(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).double().requires_grad_(grad)


def close(got, want, tol=2e-5):
    want = want.detach().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    scale = max(float(np.abs(want).max()), 1e-6)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= tol * scale


@pytest.mark.parametrize('k,stride,pad', [(1, 1, 0), (3, 1, 1), (1, 2, 0), (7, 2, 3), (3, 2, 1)])
def test_conv2d_forward_backward(k, stride, pad):
    rs = RS(k * 10 + stride)
    x = rs.standard_normal((2, 5, 13, 17)).astype(np.float32)
    W = rs.standard_normal((6, 5, k, k)).astype(np.float32)
    b = rs.standard_normal(6).astype(np.float32)
    xt, Wt, bt = t(x, True), t(W, True), t(b, True)
    yt = F.conv2d(xt, Wt, bt, stride=stride, padding=pad)
    y = onn.conv2d(x, W, b, stride, pad)
    close(y, yt)
    gy = rs.standard_normal(y.shape).astype(np.float32)
    yt.backward(t(gy))
    gx, gW, gb = onn.conv2d_backward(x, W, gy, stride, pad)
    close(gx, xt.grad)
    close(gW, Wt.grad)
    close(gb, bt.grad)


def test_deconv2d_forward_backward():
    rs = RS(1)
    x = rs.standard_normal((3, 8, 7, 7)).astype(np.float32)
    W = rs.standard_normal((8, 4, 2, 2)).astype(np.float32)        # (in, out, kh, kw)
    b = rs.standard_normal(4).astype(np.float32)
    xt, Wt, bt = t(x, True), t(W, True), t(b, True)
    yt = F.conv_transpose2d(xt, Wt, bt, stride=2)
    y = onn.deconv2d(x, W, b, 2)
    close(y, yt)
    gy = rs.standard_normal(y.shape).astype(np.float32)
    yt.backward(t(gy))
    gx, gW, gb = onn.deconv2d_backward(x, W, gy, 2)
    close(gx, xt.grad)
    close(gW, Wt.grad)
    close(gb, bt.grad)


def test_linear_forward_backward():
    rs = RS(2)
    x = rs.standard_normal((9, 4, 1, 1)).astype(np.float32)
    W = rs.standard_normal((7, 4)).astype(np.float32)
    b = rs.standard_normal(7).astype(np.float32)
    xt, Wt, bt = t(x, True), t(W, True), t(b, True)
    yt = F.linear(xt.reshape(9, -1), Wt, bt)
    close(onn.linear(x, W, b), yt)
    gy = rs.standard_normal((9, 7)).astype(np.float32)
    yt.backward(t(gy))
    gx, gW, gb = onn.linear_backward(x, W, gy)
    close(gx, xt.grad)
    close(gW, Wt.grad)
    close(gb, bt.grad)


@pytest.mark.parametrize('h,w', [(12, 16), (13, 17), (200, 334 // 2)])
def test_max_pooling_cover_all_is_ceil_mode(h, w):
    """max_pooling_2d(3, stride=2, pad=1) with Chainer's default cover_all=True: the output
    size rounds up (resnet_extractor.py:67-69; 400 x 667 -> 201 x 334 in SURVEY.md 8a)."""
    x = RS(h).standard_normal((2, 3, h, w)).astype(np.float32)
    y = onn.max_pooling_2d(x, 3, 2, 1, cover_all=True)
    assert y.shape[2:] == ((h + 2 - 3 + 1) // 2 + 1, (w + 2 - 3 + 1) // 2 + 1)
    # torch's ceil_mode drops a last window that starts in the right padding; Chainer keeps
    # it -- compare on the windows both produce, and check the extra ones by hand
    yt = F.max_pool2d(t(x), 3, 2, 1, ceil_mode=True).numpy()
    hh, ww = min(y.shape[2], yt.shape[2]), min(y.shape[3], yt.shape[3])
    np.testing.assert_array_equal(y[:, :, :hh, :ww], yt[:, :, :hh, :ww].astype(np.float32))
    pad = np.full((2, 3, h + 4, w + 4), -np.inf, np.float32)
    pad[:, :, 1:h + 1, 1:w + 1] = x
    for oy in range(hh, y.shape[2]):
        for ox in range(y.shape[3]):
            np.testing.assert_array_equal(
                y[:, :, oy, ox], pad[:, :, 2 * oy:2 * oy + 3, 2 * ox:2 * ox + 3].max(axis=(2, 3)))
    assert onn.max_pooling_2d(x, 3, 2, 1, cover_all=False).shape[2:] == \
        ((h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1)


def test_average_pooling_forward_backward():
    rs = RS(3)
    x = rs.standard_normal((4, 6, 7, 7)).astype(np.float32)
    xt = t(x, True)
    yt = F.avg_pool2d(xt, 7, 7)
    y = onn.average_pooling_2d(x, 7, 7)
    close(y, yt)
    gy = rs.standard_normal(y.shape).astype(np.float32)
    yt.backward(t(gy))
    close(onn.average_pooling_2d_backward(x.shape, gy, 7, 7), xt.grad)


def test_affine_channel_forward_backward():
    rs = RS(4)
    x = rs.standard_normal((2, 5, 6, 7)).astype(np.float32)
    W = rs.uniform(0.5, 1.5, 5).astype(np.float32)
    b = rs.standard_normal(5).astype(np.float32)
    xt, Wt, bt = t(x, True), t(W, True), t(b, True)
    yt = Wt.view(1, -1, 1, 1) * xt + bt.view(1, -1, 1, 1)
    close(onn.affine_channel_2d(x, W, b), yt)
    gy = rs.standard_normal(x.shape).astype(np.float32)
    yt.backward(t(gy))
    gx, gW, gb = onn.affine_channel_2d_backward(x, W, gy)
    close(gx, xt.grad)
    close(gW, Wt.grad)
    close(gb, bt.grad)


def test_sigmoid_cross_entropy_with_ignore():
    rs = RS(5)
    x = (rs.standard_normal((6, 50)) * 4).astype(np.float32)
    tt = rs.randint(-1, 2, x.shape).astype(np.int32)              # -1 = ignore
    xt = t(x, True)
    valid = torch.from_numpy(tt != -1)
    lt = F.binary_cross_entropy_with_logits(xt[valid], torch.from_numpy(tt).double()[valid],
                                            reduction='sum') / int(valid.sum())
    loss, gx = onn.sigmoid_cross_entropy(x, tt)
    assert abs(float(loss) - lt.item()) <= 2e-6 * max(abs(lt.item()), 1.)
    lt.backward()
    close(gx, xt.grad, tol=1e-5)
    # everything ignored: zero loss and gradient, no division by zero
    loss, gx = onn.sigmoid_cross_entropy(x, np.full(x.shape, -1, np.int32))
    assert float(loss) == 0. and not gx.any()


def test_softmax_cross_entropy_with_ignore():
    rs = RS(6)
    x = (rs.standard_normal((40, 81)) * 3).astype(np.float32)
    tt = rs.randint(-1, 81, 40).astype(np.int32)
    xt = t(x, True)
    lt = F.cross_entropy(xt, torch.from_numpy(tt).long(), ignore_index=-1)
    loss, gx = onn.softmax_cross_entropy(x, tt)
    assert abs(float(loss) - lt.item()) <= 2e-6 * max(abs(lt.item()), 1.)
    lt.backward()
    close(gx, xt.grad, tol=1e-5)
    close(onn.softmax(x), F.softmax(t(x), dim=1), tol=1e-6)


@pytest.mark.parametrize('sigma', [1., 3.])
def test_smooth_l1_and_loc_loss(sigma):
    """_smooth_l1_loss / _fast_rcnn_loc_loss (mask_rcnn_train_chain.py:192-213): the
    Huber loss with beta = 1 / sigma^2, summed, over rows with label > 0, divided by the
    number of rows with label >= 0."""
    rs = RS(int(sigma))
    pred = rs.standard_normal((30, 4)).astype(np.float32)
    gt = rs.standard_normal((30, 4)).astype(np.float32)
    label = rs.randint(-1, 3, 30).astype(np.int32)
    pt = t(pred, True)
    pos = torch.from_numpy(label > 0)
    lt = F.smooth_l1_loss(pt[pos], t(gt)[pos], reduction='sum', beta=1. / sigma ** 2)
    lt = lt / int((label >= 0).sum())
    loss, g = onn.fast_rcnn_loc_loss(pred, gt, label, sigma)
    assert abs(float(loss) - lt.item()) <= 2e-6 * max(abs(lt.item()), 1.)
    lt.backward()
    close(g, pt.grad, tol=1e-5)


def test_affine_channel_matches_reference_golden(golden_dir):
    """tests/golden/affine_channel.npz: functions/affine_channel_2d.py run verbatim."""
    import os
    g = np.load(os.path.join(golden_dir, 'affine_channel.npz'))
    y = onn.affine_channel_2d(g['x'], g['W'].reshape(-1), g['b'].reshape(-1))
    np.testing.assert_array_equal(y, g['y'])
    gx, gW, gb = onn.affine_channel_2d_backward(g['x'], g['W'].reshape(-1), g['gy'])
    np.testing.assert_array_equal(gx, g['gx'])
    np.testing.assert_allclose(gW, g['gW'].reshape(-1), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gb, g['gb'].reshape(-1), rtol=1e-6, atol=1e-6)
