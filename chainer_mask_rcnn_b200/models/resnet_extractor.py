"""ResNet-C4 feature extractor (conv1 .. res4) on the tensor-core kernels.

Mirrors ``ResNetExtractorBase`` / ``ResNet50Extractor`` / ``ResNet101Extractor``
(chainer_mask_rcnn/models/resnet_extractor.py:47-124): conv1 (7x7, stride 2, pad 3,
with bias) -> AffineChannel2D -> ReLU -> max_pooling_2d(3, stride=2, pad=1) [Chainer's
cover_all=True] -> res2 -> res3 -> res4, every BatchNormalization replaced by a frozen
AffineChannel2D (:16-44), gradients stopped after res2 (``freeze_at``, :86-87).

Input is the reference's (B, 3, H, W) float32 image batch; the returned feature map
is shaped (B, 1024, H/16, W/16) like the reference's, as a channels-last view (the
kernels work on NHWC memory).
"""
import ctypes

import numpy as np
import torch

from . import engine as E
from .layers import BuildingBlock, Conv

N_BLOCKS = {50: (3, 4, 6), 101: (3, 4, 23)}

BN_EPS = 1e-5       # resnet_extractor.py:23
_BN_LEAVES = ('gamma', 'beta', 'avg_mean', 'avg_var')


def _get_affine_from_bn(bn):
    """BatchNormalization -> AffineChannel2D (resnet_extractor.py:16-29):
    ``W = gamma / sqrt(avg_var + 1e-5)``, ``b = beta - avg_mean * W``, computed on the device
    by ``cmr_bn_fold`` one IEEE fp32 operation at a time (bit-exact with the reference's
    NumPy expression).

    ``bn``: an object with the reference's attribute names (``gamma`` / ``beta`` arrays or
    objects holding ``.data``, ``avg_mean``, ``avg_var``) or a mapping with those keys.
    -> (W, b) float32 CUDA tensors of shape (C,)."""
    from .mask_rcnn import as_device_f32

    def field(name):
        v = bn[name] if isinstance(bn, dict) else getattr(bn, name)
        v = getattr(v, 'data', v) if not isinstance(v, (np.ndarray, torch.Tensor)) else v
        return as_device_f32(np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) \
            .contiguous().view(-1)

    gamma, beta, mean, var = (field(n) for n in _BN_LEAVES)
    C = gamma.numel()
    if not (beta.numel() == mean.numel() == var.numel() == C):
        raise ValueError('gamma, beta, avg_mean and avg_var must have the same size')
    W = torch.empty_like(gamma)
    b = torch.empty_like(gamma)
    E._lib.call('cmr_bn_fold', E._p(gamma), E._p(beta), E._p(mean), E._p(var),
                ctypes.c_float(BN_EPS), C, E._p(W), E._p(b), E.stream())
    return W, b


def _convert_bn_to_affine(params):
    """The reference walks a chain and swaps every ``L.BatchNormalization`` for an
    ``AffineChannel2D`` in place (resnet_extractor.py:32-44).  Here parameters are a flat
    name -> array mapping (the npz snapshot): every ``<link>/{gamma,beta,avg_mean,avg_var}``
    group becomes ``<link>/{W,b}`` (Chainer's batch counter ``<link>/N`` is dropped), every
    other entry is passed through.  -> new dict; folded values are CUDA tensors."""
    out = {}
    for key, value in params.items():
        root, _, leaf = key.rpartition('/')
        if leaf in ('beta', 'avg_mean', 'avg_var', 'N') and (root + '/gamma') in params:
            continue
        if leaf == 'gamma':
            out[root + '/W'], out[root + '/b'] = _get_affine_from_bn(
                {n: params[root + '/' + n] for n in _BN_LEAVES})
        else:
            out[key] = value
    return out


class _Stem(Conv):
    """conv1 read as an implicit GEMM over RGB0-packed rows (see cmr_pack_image_nhwc4):
    K = 7 filter rows x 32 floats (8 pixels x 4 channels; the 8th pixel's weights are 0)."""

    def prep_frozen(self):
        super(_Stem, self).prep_frozen()
        c = self.ctx
        w = c.param(self.W)                                   # (64, 7, 7, 3) OHWI
        packed = torch.zeros((self.cout, 7, 8, 4), dtype=torch.float32, device=c.device)
        packed[:, :, :7, :3] = w
        self.w_packed = E.round_tf32(packed.view(self.cout, 7, 1, 32).contiguous())

    def forward(self, x_nchw):
        B, _, H, W = x_nchw.shape
        if self.ctx.precision == 'tf32x3':
            # parity mode: a plain 7x7 convolution over [hi | lo | hi] pixels of 3 -> 32
            # zero-padded channels each (K = 49 x 96)
            c = self.ctx
            px = torch.empty((B, H * W, 3), dtype=torch.float32, device=x_nchw.device)
            E._lib.call('cmr_transpose_batched', E._p(x_nchw.contiguous()), B, 3, H * W, E._p(px),
                        E.stream())
            x96 = E.split3(px.view(B, H, W, 3), order=0, c_pad=32)
            if getattr(self, '_w96', (None, None))[0] != c.version:
                self._w96 = (c.version, E.split3(c.param(self.W), order=1, c_pad=32))
            scale, bias = self._epilogue()
            return E.conv_gemm(x96, self._w96[1], self.cout, 7, 7, 2, 3, scale=scale, bias=bias,
                               relu=True, round_out=False)
        oh, ow = E.conv_out(H, 7, 2, 3), E.conv_out(W, 7, 2, 3)
        hp = H + 6
        wp = max(W + 6, 2 * (ow - 1) + 8)
        wp += (-wp) % 4
        packed = torch.empty((B, hp, wp, 4), dtype=torch.float32, device=x_nchw.device)
        E._lib.call('cmr_pack_image_nhwc4', E._p(x_nchw), B, H, W, hp, wp, 3, 3, E._p(packed),
                    E.stream())
        scale, bias = self._epilogue()
        return E.conv_gemm(packed, self.w_packed, self.cout, 7, 1, 2, 0, scale=scale, bias=bias,
                           relu=True, in_c=32, in_ld=4, out_hw=(oh, ow))


class ResNetExtractorBase(object):

    target_layer = 'res4'
    freeze_at = 'res2'

    def __init__(self, ctx, n_layers, base=64, root='extractor'):
        n2, n3, n4 = N_BLOCKS[n_layers]
        b = base
        self.ctx = ctx
        self.conv1 = _Stem(ctx, root + '/conv1', 3, b, 7, 2, 3, trainable=False, bias=True,
                           affine=root + '/bn1')
        self.res2 = BuildingBlock(ctx, root + '/res2', n2, b, b, 4 * b, 1, trainable=False)
        # res3.a's input gradient is never needed: backprop stops at res2
        self.res3 = BuildingBlock(ctx, root + '/res3', n3, 4 * b, 2 * b, 8 * b, 2, need_gx=False)
        self.res4 = BuildingBlock(ctx, root + '/res4', n4, 8 * b, 4 * b, 16 * b, 2)
        self.out_channels = 16 * b

    def forward_nhwc(self, x_nchw):
        h = self.conv1.forward(x_nchw)
        h = E.max_pool(h, 3, 2, 1, cover_all=True)
        rec = self.ctx.recording
        self.ctx.recording = False          # nothing below res3 is differentiated
        h = self.res2.forward(h)
        self.ctx.recording = rec
        h = self.res3.forward(h)
        return self.res4.forward(h)

    def __call__(self, x):
        from .mask_rcnn import as_device_f32
        self.ctx.prepare(backward=False)
        return E.as_nchw_view(self.forward_nhwc(as_device_f32(x)))

    def backward(self, g_feat, progress=None):
        """g_feat: dL/d(res4 output) * ReLU mask, NHWC, tf32-rounded.  ``progress``: see
        BuildingBlock.backward."""
        g = self.res4.backward(g_feat, input_is_relu=True, progress=progress)
        self.res3.backward(g, input_is_relu=True, progress=progress)


class ResNet50Extractor(ResNetExtractorBase):
    def __init__(self, ctx, **kw):
        super(ResNet50Extractor, self).__init__(ctx, 50, **kw)


class ResNet101Extractor(ResNetExtractorBase):
    def __init__(self, ctx, **kw):
        super(ResNet101Extractor, self).__init__(ctx, 101, **kw)
