"""AffineChannel2D as a stand-alone operator (a5) and the BatchNormalization ->
AffineChannel2D fold (a4) on the B200, through the C ABI (csrc/affine.cu), against golden
vectors produced by the reference's own files run verbatim (tests/golden/make_golden.py)
and against the oracle on larger shapes.

Tolerances: the forward is two IEEE operations per element -> bit-exact; ``gx = W * gy`` is
one -> bit-exact; ``gW`` / ``gb`` are fp32 sums over N*H*W elements in a different (fixed)
order than NumPy's pairwise sum -> 1e-5 of max|ref| (north star: 1e-3); the fold is four
IEEE operations per channel -> bit-exact."""
import os

import numpy as np
import pytest
import torch

from chainer_mask_rcnn_b200 import functions, models
from chainer_mask_rcnn_b200.models import resnet_extractor as rx
from oracle import model as om
from oracle import nn as onn

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def test_reference_golden_forward_backward(golden_dir):
    g = np.load(os.path.join(golden_dir, 'affine_channel.npz'))
    y = functions.affine_channel_2d(g['x'], g['W'], g['b'])
    assert isinstance(y, np.ndarray) and y.dtype == np.float32
    np.testing.assert_array_equal(y, g['y'])
    x = torch.from_numpy(g['x']).cuda().requires_grad_(True)
    W = torch.from_numpy(g['W']).cuda().requires_grad_(True)
    b = torch.from_numpy(g['b']).cuda().requires_grad_(True)
    out = functions.affine_channel_2d(x, W, b)
    out.backward(torch.from_numpy(g['gy']).cuda())
    np.testing.assert_array_equal(x.grad.cpu().numpy(), g['gx'])
    assert W.grad.shape == g['gW'].shape and b.grad.shape == g['gb'].shape
    assert _rel(W.grad.cpu().numpy(), g['gW']) <= 1e-5
    assert _rel(b.grad.cpu().numpy(), g['gb']) <= 1e-5


@pytest.mark.parametrize('shape', [(2, 64, 51, 84), (3, 5, 7, 9), (1, 1024, 13, 3), (2, 256, 1, 1),
                                   (1, 3, 200, 333)])
def test_against_oracle(shape):
    """Odd plane sizes (unaligned vector lanes), few and many channels, 1x1 planes."""
    rs = np.random.RandomState(sum(shape))
    N, C, H, Wd = shape
    x = rs.standard_normal(shape).astype(np.float32)
    W = rs.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    b = rs.standard_normal((1, C, 1, 1)).astype(np.float32)
    gy = rs.standard_normal(shape).astype(np.float32)
    want = onn.affine_channel_2d(x, W.reshape(-1), b.reshape(-1))
    xt, Wt, bt = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (x, W, b))
    y = functions.affine_channel_2d(xt, Wt, bt)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), want)
    y.backward(torch.from_numpy(gy).cuda())
    w_gx, w_gW, w_gb = onn.affine_channel_2d_backward(x, W.reshape(-1), gy)
    np.testing.assert_array_equal(xt.grad.cpu().numpy(), w_gx)
    assert _rel(Wt.grad.cpu().numpy().reshape(-1), np.asarray(w_gW).reshape(-1)) <= 1e-5
    assert _rel(bt.grad.cpu().numpy().reshape(-1), np.asarray(w_gb).reshape(-1)) <= 1e-5
    # deterministic reduction: a second run gives the same bits
    xt.grad = Wt.grad = bt.grad = None
    y2 = functions.affine_channel_2d(xt, Wt, bt)
    gW1 = None
    for _ in range(2):
        Wt.grad = None
        functions.affine_channel_2d(xt, Wt, bt).backward(torch.from_numpy(gy).cuda())
        if gW1 is None:
            gW1 = Wt.grad.clone()
    assert torch.equal(gW1, Wt.grad) and torch.equal(y2, y)


def test_type_errors_like_reference():
    x = np.zeros((1, 2, 4, 4), np.float32)
    with pytest.raises(TypeError):
        functions.affine_channel_2d(x.astype(np.float64), np.ones((1, 2, 1, 1), np.float32),
                                    np.ones((1, 2, 1, 1), np.float32))
    with pytest.raises(TypeError):
        functions.affine_channel_2d(x, np.ones((1, 3, 1, 1), np.float32),
                                    np.ones((1, 3, 1, 1), np.float32))


# ----------------------------------------------------------------- a4: BN fold --
LINKS = ('bn1', 'res2/a/bn1', 'res2/a/bn4', 'res3/b2/bn3')


def test_get_affine_from_bn_bit_exact_with_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'bn_fold.npz'))
    for link in LINKS:
        bn = {n: g['%s/%s' % (link, n)] for n in ('gamma', 'beta', 'avg_mean', 'avg_var')}
        W, b = rx._get_affine_from_bn(bn)
        np.testing.assert_array_equal(W.cpu().numpy(), g[link + '/W'])
        np.testing.assert_array_equal(b.cpu().numpy(), g[link + '/b'])

    class BN(object):           # attribute form, arrays wrapped like chainer.Parameter
        pass
    bn = BN()
    bn.gamma = type('V', (), {'data': g['bn1/gamma']})()
    bn.beta = type('V', (), {'data': g['bn1/beta']})()
    bn.avg_mean, bn.avg_var = g['bn1/avg_mean'], g['bn1/avg_var']
    W, b = rx._get_affine_from_bn(bn)
    np.testing.assert_array_equal(W.cpu().numpy(), g['bn1/W'])


def test_convert_bn_to_affine_flat_snapshot(golden_dir):
    g = np.load(os.path.join(golden_dir, 'bn_fold.npz'))
    params = {k: g[k] for k in g.files if k.rsplit('/', 1)[1] in
              ('gamma', 'beta', 'avg_mean', 'avg_var')}
    params['bn1/N'] = np.int64(3)
    params['conv1/W'] = np.ones((2, 3, 7, 7), np.float32)
    out = rx._convert_bn_to_affine(params)
    assert sorted(out) == sorted(['conv1/W'] + [l + s for l in LINKS for s in ('/W', '/b')])
    for link in LINKS:
        np.testing.assert_array_equal(out[link + '/W'].cpu().numpy(), g[link + '/W'])


def test_load_imagenet_resnet_folds_and_places_every_layer():
    """A synthetic Chainer ResNet50Layers snapshot (BatchNormalization statistics, BGR conv1)
    loaded through MaskRCNNResNet.load_imagenet_resnet equals the oracle's conversion:
    conv1 flipped to RGB, every BN folded, conv1..res4 under extractor/, res5 under head/."""
    base = 8
    rs = np.random.RandomState(0)
    cfg = om.Config(n_layers=50, n_fg_class=3, anchor_scales=(4, 8), roi_size=14, base=base)
    affine = om.make_params(cfg, rs)
    snap = {}
    for name, v in affine.items():
        top = name.split('/')[0]
        if top not in ('extractor', 'head') or name.split('/')[1] not in (
                'conv1', 'bn1', 'res2', 'res3', 'res4', 'res5'):
            continue
        key = name.split('/', 1)[1]
        if '/bn' in '/' + key and key.endswith('/W'):
            root = key[:-2]
            c = v.shape[0]
            snap[root + '/gamma'] = rs.uniform(0.5, 1.5, c).astype(np.float32)
            snap[root + '/beta'] = rs.standard_normal(c).astype(np.float32)
            snap[root + '/avg_mean'] = rs.standard_normal(c).astype(np.float32)
            snap[root + '/avg_var'] = rs.uniform(1e-6, 4., c).astype(np.float32)
            snap[root + '/N'] = np.int64(1)
        elif '/bn' in '/' + key:
            continue
        else:
            snap[key] = v
    snap['fc6/W'] = np.zeros((10, 32 * base), np.float32)
    want = onn.convert_bn_to_affine(snap)
    want['conv1/W'] = want['conv1/W'][:, ::-1]
    model = models.MaskRCNNResNet(50, 3, anchor_scales=(4, 8), roi_size=14, base_channels=base)
    before = model.state_dict()
    model.load_imagenet_resnet(snap)
    got = model.state_dict()
    for key, v in want.items():
        if key.startswith('fc6'):
            continue
        name = ('head/' if key.startswith('res5') else 'extractor/') + key
        np.testing.assert_array_equal(got[name], v, err_msg=name)
    for name in ('rpn/conv1/W', 'head/score/W', 'head/mask/b'):     # untouched
        np.testing.assert_array_equal(got[name], before[name])
    bad = dict(snap)
    del bad['res3/a/bn2/gamma']
    with pytest.raises(KeyError):
        model.load_imagenet_resnet(bad)
