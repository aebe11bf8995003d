"""Every distinct tensor-core GEMM of BASELINE.json configs[1] (R50-C4 train step, bs 2,
3 x 800 x 1333, 1024 sampled RoIs) at FULL SIZE with raw fp32 operands against an fp64
reference of the same operation: the north-star bound

        max|got - want| / max|want|  <=  1e-3

One eager train step of the full-size model is run with engine.conv_gemm / engine.wgrad_tap
hooked to record every launch's geometry (the same enumeration tools/layer_bench.py times).
Each distinct geometry is then re-launched through the C ABI on random raw fp32 operands of
that exact shape -- rounded to TF32 first, as the producing layer's epilogue / the weight
preparation does in the model -- and compared with torch's fp64 convolution (forward / data
gradient GEMMs) or fp64 per-tap matrix products (weight-gradient GEMMs) of the UNROUNDED
operands.  Epilogue terms (affine, bias, residual, ReLU mask) have their own tests
(test_gpu_conv_tc.py); here the contraction itself is checked on every shape, tile width,
CTA-pair / single-CTA path and split the step uses."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from chainer_mask_rcnn_b200 import models, optimizers
from chainer_mask_rcnn_b200.models import engine as E

pytestmark = pytest.mark.gpu

TOL = 1e-3


def rel(got, want):
    return float((got.double() - want).abs().max() / want.abs().max())


@pytest.fixture(scope='module')
def shapes():
    import bench
    model = models.MaskRCNNResNet(50, 80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333)
    chain = models.MaskRCNNTrainChain(model)
    opt = optimizers.MomentumSGD(lr=0.0025).setup(chain)
    imgs, bboxes, labels, masks, scales = bench.synth_batch(0)
    masks = models.utils.PackedMasks.from_numpy(np.stack(masks)).to('cuda')
    x = torch.from_numpy(imgs).cuda()
    gemms, wgrads = {}, {}
    conv0, wgrad0 = E.conv_gemm, E.wgrad_tap

    def conv_hook(x, w, n, kh=1, kw=1, stride=1, pad=0, **kw_):
        out = conv0(x, w, n, kh, kw, stride, pad, **kw_)
        key = (tuple(x.shape), n, kh, kw, stride, pad, kw_.get('in_c'), kw_.get('in_ld'),
               kw_.get('out_hw'), kw_.get('d_stride', 1), kw_.get('tap_cols', 0),
               tuple(out.shape))
        gemms.setdefault(key, 0)
        gemms[key] += 1
        return out

    def wgrad_hook(gy, x, gw, rows, cols, loop_hw, gw_ld, **kw_):
        wgrad0(gy, x, gw, rows, cols, loop_hw, gw_ld, **kw_)
        key = (tuple(gy.shape), tuple(x.shape), rows, cols, tuple(loop_hw), gw_ld,
               kw_.get('gw_col0', 0), kw_.get('gy_stride', 1), tuple(kw_.get('gy_off', (0, 0))),
               kw_.get('gy_c0', 0), kw_.get('x_stride', 1), tuple(kw_.get('x_off', (0, 0))),
               kw_.get('x_c0', 0), tuple(kw_.get('taps', (1, 1))))
        wgrads.setdefault(key, 0)
        wgrads[key] += 1

    E.conv_gemm, E.wgrad_tap = conv_hook, wgrad_hook
    try:
        opt.update(chain, x, bboxes, labels, masks, scales)
        torch.cuda.synchronize()
    finally:
        E.conv_gemm, E.wgrad_tap = conv0, wgrad0
    del model, chain, opt
    torch.cuda.empty_cache()
    return gemms, wgrads


def _gemm_case(key, g):
    (xs, n, kh, kw, stride, pad, in_c, in_ld, out_hw, d_stride, tap_cols, outs) = key
    B, H, W, C = xs
    x = torch.randn(xs, device='cuda', generator=g)
    if in_c is not None and in_ld is not None and in_ld < in_c:
        # the stem: rows of RGB0 pixels read 32 floats (8 pixels) per filter row; as a
        # convolution: a (7 x 8)-pixel window over 4 channels, stride 2
        w = torch.randn((n, kh, 1, in_c), device='cuda', generator=g) / (kh * in_c) ** 0.5
        px = in_c // in_ld
        wd = w.view(n, kh, px, in_ld).permute(0, 3, 1, 2).double()
        want = F.conv2d(x.permute(0, 3, 1, 2).double(), wd, stride=stride)
        want = want[:, :, :out_hw[0], :out_hw[1]].permute(0, 2, 3, 1)
        got = E.conv_gemm(E.round_tf32(x.clone()), E.round_tf32(w.clone()), n, kh, kw, stride, pad,
                          in_c=in_c, in_ld=in_ld, out_hw=out_hw, round_out=False)
        return rel(got, want)
    w = torch.randn((n, kh, kw, C), device='cuda', generator=g) / (kh * kw * C) ** 0.5
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(),
                    stride=stride, padding=pad).permute(0, 2, 3, 1)
    xr, wr = E.round_tf32(x.clone()), E.round_tf32(w.clone())
    if tap_cols:            # fused 2x2 deconvolution: four column blocks, pixel-shuffled
        got = torch.zeros(outs, device='cuda')
        E.conv_gemm(xr, wr, n, out=got, round_out=False, d_stride=2, tap_cols=tap_cols)
        oh, ow = want.shape[1:3]
        want = want.view(B, oh, ow, 2, 2, tap_cols).permute(0, 1, 3, 2, 4, 5) \
            .reshape(B, 2 * oh, 2 * ow, tap_cols)
        return rel(got, want)
    if d_stride > 1:        # data gradient of a stride-2 1x1 convolution: scattered rows
        got = torch.zeros(outs, device='cuda')
        E.conv_gemm(xr, wr, n, kh, kw, stride, pad, out=got, round_out=False, d_stride=d_stride)
        oh, ow = want.shape[1:3]
        sub = got[:, ::d_stride, ::d_stride][:, :oh, :ow].clone()
        got[:, ::d_stride, ::d_stride] = 0
        assert float(got.abs().max()) == 0.                         # nothing written elsewhere
        return rel(sub, want)
    got = torch.zeros(outs, device='cuda')
    E.conv_gemm(xr, wr, n, kh, kw, stride, pad, out=got, round_out=False)
    return rel(got[..., :n], want)


def test_every_forward_and_data_gradient_gemm_shape(shapes):
    gemms, _ = shapes
    assert len(gemms) >= 30
    g = torch.Generator(device='cuda').manual_seed(0)
    errs = {k: _gemm_case(k, g) for k in sorted(gemms, key=repr)}
    bad = {k: v for k, v in errs.items() if not v <= TOL}
    assert not bad, bad
    # the step's GEMM inventory: SURVEY.md 8d counts 2.153 + 2.075 TFLOP of forward + data
    # gradient work per step; the recorded launches add up to it
    flops = 0.
    for (xs, n, kh, kw, stride, pad, in_c, in_ld, out_hw, d_stride, tap_cols, outs), cnt in gemms.items():
        oh, ow = out_hw or (E.conv_out(xs[1], kh, stride, pad), E.conv_out(xs[2], kw, stride, pad))
        flops += cnt * 2. * xs[0] * oh * ow * n * kh * kw * (in_c or xs[3])
    assert 4.15e12 <= flops <= 4.5e12, flops      # padded K of the fused banks included


def _wgrad_case(key, g):
    (gys, xs, rows, cols, loop_hw, gw_ld, col0, gy_stride, gy_off, gy_c0, x_stride, x_off, x_c0,
     taps) = key
    gy = torch.randn(gys, device='cuda', generator=g)
    x = torch.randn(xs, device='cuda', generator=g)
    lh, lw = loop_hw
    T = taps[0] * taps[1]
    gw = torch.zeros((rows, max(gw_ld, col0 + T * cols)), device='cuda')
    E.wgrad_tap(E.round_tf32(gy.clone()), E.round_tf32(x.clone()), gw, rows, cols, loop_hw, gw_ld,
                gw_col0=col0, gy_stride=gy_stride, gy_off=gy_off, gy_c0=gy_c0,
                x_stride=x_stride, x_off=x_off, x_c0=x_c0, taps=taps)
    # fp64 reference: per tap, (rows x pixels) @ (pixels x cols) on the gathered pixels
    gsel = gy[:, gy_off[0]:gy_off[0] + gy_stride * lh:gy_stride,
              gy_off[1]:gy_off[1] + gy_stride * lw:gy_stride, gy_c0:gy_c0 + rows]
    assert gsel.shape[1:3] == (lh, lw)
    gm = gsel.reshape(-1, rows).double()
    worst = 0.
    pad_lo = max(0, -x_off[0], -x_off[1])
    pad_hi = max(0, x_off[0] + taps[0] - 1 + x_stride * (lh - 1) + 1 - xs[1],
                 x_off[1] + taps[1] - 1 + x_stride * (lw - 1) + 1 - xs[2])
    xp = F.pad(x[..., x_c0:x_c0 + cols], (0, 0, pad_lo, pad_hi, pad_lo, pad_hi))
    want = torch.empty((rows, T, cols), device='cuda', dtype=torch.float64)
    for fr in range(taps[0]):
        for fs in range(taps[1]):
            y0, x0 = x_off[0] + fr + pad_lo, x_off[1] + fs + pad_lo
            xsel = xp[:, y0:y0 + x_stride * lh:x_stride, x0:x0 + x_stride * lw:x_stride]
            want[:, fr * taps[1] + fs] = gm.t() @ xsel.reshape(-1, cols).double()
    got = gw[:, col0:col0 + T * cols].view(rows, T, cols)
    worst = rel(got, want)
    return worst


def test_every_weight_gradient_gemm_shape(shapes):
    _, wgrads = shapes
    assert len(wgrads) >= 20
    g = torch.Generator(device='cuda').manual_seed(1)
    errs = {k: _wgrad_case(k, g) for k in sorted(wgrads, key=repr)}
    bad = {k: v for k, v in errs.items() if not v <= TOL}
    assert not bad, bad
