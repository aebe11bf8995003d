"""BatchNormalization -> AffineChannel2D (a4): the oracle's restatement against golden
vectors made by the reference's own ``_get_affine_from_bn`` / ``_convert_bn_to_affine``
(models/resnet_extractor.py:16-44) run verbatim, and -- when the reference tree is
present -- against that code again."""
import os

import numpy as np
import pytest

from oracle import nn as onn
from oracle import ref_loader

LINKS = ('bn1', 'res2/a/bn1', 'res2/a/bn4', 'res3/b2/bn3')


def test_oracle_fold_is_bit_exact_with_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'bn_fold.npz'))
    for link in LINKS:
        W, b = onn.bn_to_affine(g[link + '/gamma'], g[link + '/beta'], g[link + '/avg_mean'],
                                g[link + '/avg_var'])
        assert W.dtype == np.float32 and b.dtype == np.float32
        np.testing.assert_array_equal(W, g[link + '/W'])
        np.testing.assert_array_equal(b, g[link + '/b'])


def test_oracle_convert_walks_a_flat_snapshot(golden_dir):
    g = np.load(os.path.join(golden_dir, 'bn_fold.npz'))
    params = {k: g[k] for k in g.files if k.rsplit('/', 1)[1] in
              ('gamma', 'beta', 'avg_mean', 'avg_var')}
    params['bn1/N'] = np.int64(7)                        # Chainer's batch counter
    params['conv1/W'] = np.ones((2, 3, 7, 7), np.float32)
    out = onn.convert_bn_to_affine(params)
    assert sorted(out) == sorted(['conv1/W'] + [l + s for l in LINKS for s in ('/W', '/b')])
    for link in LINKS:
        np.testing.assert_array_equal(out[link + '/W'], g[link + '/W'])
        np.testing.assert_array_equal(out[link + '/b'], g[link + '/b'])


@pytest.mark.skipif(not ref_loader.reference_available(), reason='needs /root/reference')
def test_reference_code_reproduces_the_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'bn_fold.npz'))
    tree = {l: tuple(g['%s/%s' % (l, n)] for n in ('gamma', 'beta', 'avg_mean', 'avg_var'))
            for l in LINKS}
    for link, (W, b) in ref_loader.ref_convert_bn_to_affine(tree).items():
        np.testing.assert_array_equal(W, g[link + '/W'])
        np.testing.assert_array_equal(b, g[link + '/b'])
