"""tcgen05 implicit-GEMM convolution vs an fp32 reference of the same op.

Operands are rounded to TF32 first (what the producing layer's epilogue / the
weight preparation does in the model), so the products are exact in fp32 and the
only difference left is the accumulation order: tolerance 1e-4 of max|ref|.  A
second check feeds raw fp32 operands and asserts the north-star bound (<= 1e-3 of
max|ref|) for the tensor core's own truncation.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from chainer_mask_rcnn_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tma_im2col', 'cp_async'], autouse=True)
def a_operand_path(request):
    """Every case runs with the activation operand fetched by im2col-mode TMA (the
    default) and by the cp.async gather fallback."""
    lib = _lib.load()
    old = lib.cmr_set_im2col_tma(1 if request.param == 'tma_im2col' else 0)
    yield request.param
    lib.cmr_set_im2col_tma(old)


def round_tf32(t):
    out = torch.empty_like(t)
    _lib.call('cmr_round_tf32', _lib.ptr(t), _lib.ptr(out), t.numel(), _lib.stream_ptr())
    return out


def conv_tc(x_nhwc, w_ohwi, stride, pad, scale=None, bias=None, addend=None, mask=None,
            relu=False, round_out=False, tile_n=0, d_stride=1, d_off=(0, 0), d_hw=None):
    B, H, W, C = x_nhwc.shape
    N, kh, kw, _ = w_ohwi.shape
    oh = (H + 2 * pad - kh) // stride + 1
    ow = (W + 2 * pad - kw) // stride + 1
    dh, dw = d_hw if d_hw else (oh, ow)
    d = torch.zeros((B, dh, dw, N), device='cuda')
    desc = _lib.ConvDesc(B, H, W, C, C, oh, ow, kh, kw, stride, pad, N, dh, dw, N, d_stride,
                         d_off[0], d_off[1], int(relu), int(round_out), tile_n)
    _lib.call('cmr_conv_gemm_tc', ctypes.byref(desc), _lib.ptr(x_nhwc), _lib.ptr(w_ohwi),
              _lib.ptr(d), _lib.ptr(scale), _lib.ptr(bias), _lib.ptr(addend), _lib.ptr(mask),
              _lib.stream_ptr())
    return d


def ref_conv(x_nhwc, w_ohwi, stride, pad):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).double(), w_ohwi.permute(0, 3, 1, 2).double(),
                     stride=stride, padding=pad)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return y.permute(0, 2, 3, 1).float().contiguous()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


CASES = [
    # B, H,  W,  C,   N,  k, s, p, tile
    (1, 8, 16, 32, 64, 1, 1, 0, 64),        # one tile, one k-block
    (1, 8, 16, 64, 64, 1, 1, 0, 64),        # two k-blocks
    (2, 13, 17, 128, 128, 1, 1, 0, 128),    # ragged M, ring wraps (4 k-blocks, 3 stages)
    (2, 13, 17, 256, 256, 1, 1, 0, 256),    # BN = 256
    (1, 20, 23, 64, 96, 3, 1, 1, 128),      # 3x3 pad 1, N not a tile multiple
    (2, 21, 19, 64, 128, 1, 2, 0, 128),     # 1x1 stride 2 (res3/res4 'a' blocks)
    (1, 7, 7, 512, 512, 3, 1, 1, 0),        # res5 3x3, K = 4608, auto tile
    (3, 14, 14, 1024, 80, 1, 1, 0, 0),      # mask head N = 80
    (1, 51, 84, 1024, 75, 1, 1, 0, 0),      # RPN loc+score fused, N = 75 (scalar stores)
    (3, 13, 18, 64, 64, 1, 2, 0, 64),       # stride 2, even width, tile crosses images
    (2, 28, 28, 64, 64, 2, 2, 0, 64),       # 2x2 stride 2 (deconv6 data gradient)
    (5, 7, 7, 96, 64, 3, 1, 1, 64),         # 128-row tiles spanning 3 images, K = 27 k-blocks
    (2, 10, 9, 32, 64, 3, 2, 1, 64),        # 3x3 stride 2 pad 1
    (1, 5, 6, 64, 64, 5, 1, 2, 64),         # 5x5 pad 2
    (8, 48, 50, 1024, 256, 1, 1, 0, 256),   # CTA-pair path (cta_group::2): 75 x 1 tiles of 256 rows
    (4, 70, 70, 128, 512, 3, 1, 1, 256),    # CTA-pair path, 3x3, ragged last tile, 2 column tiles
    (2, 51, 84, 128, 256, 3, 1, 1, 128),    # res4 3x3 geometry, 128-wide tiles
    (8, 48, 50, 256, 512, 1, 1, 0, 256),    # short reduction (K = 256), 150 x 2 tiles: 3 epilogue groups, 2 tiles per CTA
]


@pytest.mark.parametrize('B,H,W,C,N,k,s,p,tile', CASES)
def test_conv_matches_fp32_reference(B, H, W, C, N, k, s, p, tile):
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + H * 10 + C)
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    w = round_tf32(torch.randn((N, k, k, C), device='cuda', generator=g) / (k * C ** 0.5))
    want = ref_conv(x, w, s, p)
    got = conv_tc(x, w, s, p, tile_n=tile)
    assert got.shape == want.shape
    assert rel(got, want) <= 1e-4


def test_fused_epilogue():
    g = torch.Generator(device='cuda').manual_seed(5)
    B, H, W, C, N = 2, 9, 11, 64, 128
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    w = round_tf32(torch.randn((N, 3, 3, C), device='cuda', generator=g) / 24)
    scale = torch.rand((N,), device='cuda', generator=g) + 0.5
    bias = torch.randn((N,), device='cuda', generator=g)
    addend = torch.randn((B, H, W, N), device='cuda', generator=g)
    mask = torch.randn((B, H, W, N), device='cuda', generator=g)
    base = ref_conv(x, w, 1, 1)
    # forward epilogue: affine (conv first, then W*x+b as the reference), residual, relu
    want = torch.relu(base * scale + bias + addend)
    got = conv_tc(x, w, 1, 1, scale=scale, bias=bias, addend=addend, relu=True)
    assert rel(got, want) <= 1e-4
    # backward epilogue: add the other branch's gradient, then the ReLU mask
    want = (base + addend) * (mask > 0)
    got = conv_tc(x, w, 1, 1, addend=addend, mask=mask)
    assert rel(got, want) <= 1e-4
    # tf32-rounded output is idempotent under rounding
    got = conv_tc(x, w, 1, 1, round_out=True)
    assert torch.equal(got, round_tf32(got))
    assert rel(got, base) <= 1e-3


def test_short_reduction_epilogue_many_tiles():
    """res5 conv3 geometry (1x1, K = 512, wide N) with several tiles per persistent CTA
    (both TMEM accumulator buffers in flight, three epilogue groups): affine + residual +
    ReLU, and the masked data gradient."""
    g = torch.Generator(device='cuda').manual_seed(9)
    B, H, W, C, N = 400, 7, 7, 512, 1024          # 77 pair tiles x 4 column tiles
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    w = round_tf32(torch.randn((N, 1, 1, C), device='cuda', generator=g) / 22)
    scale = torch.rand((N,), device='cuda', generator=g) + 0.5
    bias = torch.randn((N,), device='cuda', generator=g)
    addend = torch.randn((B, H, W, N), device='cuda', generator=g)
    mask = torch.randn((B, H, W, N), device='cuda', generator=g)
    base = ref_conv(x, w, 1, 0)
    got = conv_tc(x, w, 1, 0, scale=scale, bias=bias, addend=addend, relu=True, tile_n=256)
    assert rel(got, torch.relu(base * scale + bias + addend)) <= 1e-4
    got = conv_tc(x, w, 1, 0, addend=addend, mask=mask, tile_n=256)
    assert rel(got, (base + addend) * (mask > 0)) <= 1e-4


def test_strided_scatter_output():
    """Deconvolution2D(2, stride 2) = four 1x1 GEMMs written with a pixel-shuffle store
    (d_stride 2, offsets (dy, dx)); also the data gradient of a stride-2 1x1 conv."""
    g = torch.Generator(device='cuda').manual_seed(6)
    B, H, W, C, N = 2, 7, 7, 64, 64
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    wt = round_tf32(torch.randn((C, N, 2, 2), device='cuda', generator=g) / 8)   # (in, out, kh, kw)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), stride=2)
    want = want.permute(0, 2, 3, 1).float()
    out = torch.zeros((B, 2 * H, 2 * W, N), device='cuda')
    for dy in range(2):
        for dx in range(2):
            w_tap = wt[:, :, dy, dx].t().contiguous().view(N, 1, 1, C)
            desc = _lib.ConvDesc(B, H, W, C, C, H, W, 1, 1, 1, 0, N, 2 * H, 2 * W, N, 2, dy, dx,
                                 0, 0, 0)
            _lib.call('cmr_conv_gemm_tc', ctypes.byref(desc), _lib.ptr(x), _lib.ptr(w_tap),
                      _lib.ptr(out), None, None, None, None, _lib.stream_ptr())
    assert rel(out, want) <= 1e-4


def test_raw_fp32_operands_meet_north_star_bound():
    g = torch.Generator(device='cuda').manual_seed(7)
    x = torch.randn((2, 25, 42, 256), device='cuda', generator=g)
    w = torch.randn((256, 3, 3, 256), device='cuda', generator=g) / 48
    want = ref_conv(x, w, 1, 1)
    got = conv_tc(round_tf32(x), round_tf32(w), 1, 1)
    assert rel(got, want) <= 1e-3


def test_unsupported_channels_is_an_error_not_a_fallback():
    x = torch.zeros((1, 4, 4, 3), device='cuda')
    w = torch.zeros((8, 1, 1, 3), device='cuda')
    with pytest.raises(_lib.CmrError):
        conv_tc(x, w, 1, 0)


def test_row_group_broadcast_epilogue():
    """cmr_conv_gemm_tc_ex: v += bcast[row // group] * scale between the addend and the ReLU
    mask (the average-pooling backward fused into a data-gradient GEMM)."""
    g = torch.Generator(device='cuda').manual_seed(8)
    B, H, W, C, N = 6, 7, 7, 64, 128
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    w = round_tf32(torch.randn((N, 1, 1, C), device='cuda', generator=g) / 8)
    bc = torch.randn((B, N), device='cuda', generator=g)
    mask = torch.randn((B, H, W, N), device='cuda', generator=g)
    want = (ref_conv(x, w, 1, 0) + bc.view(B, 1, 1, N) / 49.) * (mask > 0)
    d = torch.zeros((B, H, W, N), device='cuda')
    desc = _lib.ConvDesc(B, H, W, C, C, H, W, 1, 1, 1, 0, N, H, W, N, 1, 0, 0, 0, 0, 0)
    _lib.call('cmr_conv_gemm_tc_ex', ctypes.byref(desc), _lib.ptr(x), _lib.ptr(w), _lib.ptr(d),
              None, None, None, _lib.ptr(mask), _lib.ptr(bc), 49, 1.0 / 49, _lib.stream_ptr())
    assert rel(d, want) <= 1e-4


def test_fused_deconvolution_tap_columns():
    """Deconvolution2D(2, stride 2) + bias + ReLU as ONE GEMM with N = 4 * cout whose column
    blocks are pixel-shuffled by the epilogue (cmr_conv_desc.tap_cols)."""
    g = torch.Generator(device='cuda').manual_seed(9)
    B, H, W, C, N = 3, 7, 7, 64, 64
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    wt = round_tf32(torch.randn((C, N, 2, 2), device='cuda', generator=g) / 8)   # (in, out, kh, kw)
    bias = torch.randn((N,), device='cuda', generator=g)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), bias.double(), stride=2)
    want = torch.relu(want).permute(0, 2, 3, 1).float()
    w_taps = wt.permute(2, 3, 1, 0).contiguous().view(4 * N, 1, 1, C)      # (tap, out, in)
    out = torch.zeros((B, 2 * H, 2 * W, N), device='cuda')
    desc = _lib.ConvDesc(B, H, W, C, C, H, W, 1, 1, 1, 0, 4 * N, 2 * H, 2 * W, N, 2, 0, 0, 1, 0, 0, N)
    _lib.call('cmr_conv_gemm_tc', ctypes.byref(desc), _lib.ptr(x), _lib.ptr(w_taps), _lib.ptr(out),
              None, _lib.ptr(bias), None, None, _lib.stream_ptr())
    assert rel(out, want) <= 1e-4


def test_split_tf32x3_and_fp32_level_convolution():
    """cmr_split_tf32x3: hi + lo reproduces x to 2^-21 |x|, both parts are TF32 values; a
    convolution on [hi|lo|hi] x [hi|hi|lo] operands matches the fp64 reference of the RAW fp32
    operands to 5e-5 (the plain TF32 path: ~4e-4; the residual is the tensor core's truncating
    accumulator, not the operands)."""
    from chainer_mask_rcnn_b200.models import engine as E
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randn((2, 25, 42, 256), device='cuda', generator=g)
    w = torch.randn((256, 3, 3, 256), device='cuda', generator=g) / 48
    x3 = E.split3(x)
    hi, lo, hi2 = x3[..., :256], x3[..., 256:512], x3[..., 512:]
    assert torch.equal(hi, hi2) and torch.equal(hi, round_tf32(x.clone()))
    assert torch.equal(lo.contiguous(), round_tf32(lo.contiguous().clone()))
    assert float(((hi + lo) - x).abs().max() / x.abs().max()) <= 2 ** -21
    w3 = E.split3(w, order=1)
    assert torch.equal(w3[..., 256:512], w3[..., :256])
    want = ref_conv(x, w, 1, 1)
    got = E.conv_gemm(x3, w3, 256, 3, 3, 1, 1, round_out=False)
    e3 = rel(got, want)
    e1 = rel(conv_tc(round_tf32(x.clone()), round_tf32(w.clone()), 1, 1), want)
    print('tf32x3 conv error %.2e, tf32 %.2e' % (e3, e1))
    assert e3 <= 5e-5 and e3 < 0.25 * e1
    # zero-padded narrow rows (the stem: 3 -> 32 channels per part)
    px = torch.randn((5, 7, 3), device='cuda', generator=g)
    p3 = E.split3(px, c_pad=32)
    assert p3.shape == (5, 7, 96) and float(p3[..., 3:32].abs().max()) == 0.
    assert torch.equal(p3[..., :3], round_tf32(px.clone()))


@pytest.mark.parametrize('geom', [
    # (B, H, W, C, N, k, pad): tiles leave a last wave that fills less than half of the SMs
    (430, 7, 7, 1024, 256, 1, 0),      # 83 pair tiles of 256 x 256 on 74 SM pairs, 4 K-parts
    (186, 7, 7, 128, 256, 3, 1),       # 3x3, 36 pair tiles ... single wave: no split (control)
    (4, 75, 76, 1024, 128, 1, 0),      # 179 single-CTA tiles: not a pair launch, no split (control)
    (410, 7, 7, 512, 512, 3, 1),       # 79 x 2 pair tiles, im2col K ranges starting mid-filter
])
def test_k_split_tail_matches_unsplit(geom, a_operand_path):
    """cmr_conv_gemm_tc_ws: the tiles of a short last wave computed as K-parts through the
    workspace give the un-split kernel's result up to the fp32 summation order, for the forward
    epilogue (affine, residual, ReLU, tf32 rounding) and the backward one (addend, mask)."""
    B, H, W, C, N, k, pad = geom
    g = torch.Generator(device='cuda').manual_seed(21)
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    w = round_tf32(torch.randn((N, k, k, C), device='cuda', generator=g) / (k * C ** 0.5))
    scale = torch.rand((N,), device='cuda', generator=g) + 0.5
    bias = torch.randn((N,), device='cuda', generator=g)
    addend = torch.randn((B, H, W, N), device='cuda', generator=g)
    mask = torch.randn((B, H, W, N), device='cuda', generator=g)
    ws = torch.empty((int(_lib.load().cmr_conv_gemm_ws_bytes()),), dtype=torch.uint8,
                     device='cuda')

    def run(ws_t, **kw):
        d = torch.full((B, H, W, N), float('nan'), device='cuda')
        desc = _lib.ConvDesc(B, H, W, C, C, H, W, k, k, 1, pad, N, H, W, N, 1, 0, 0,
                             int(kw.get('relu', False)), int(kw.get('round_out', False)), 0)
        _lib.call('cmr_conv_gemm_tc_ws', ctypes.byref(desc), _lib.ptr(x), _lib.ptr(w),
                  _lib.ptr(d), _lib.ptr(kw.get('scale')), _lib.ptr(kw.get('bias')),
                  _lib.ptr(kw.get('addend')), _lib.ptr(kw.get('mask')), None, 1, 0.0,
                  _lib.ptr(ws_t), ws_t.numel() if ws_t is not None else 0, _lib.stream_ptr())
        return d

    base = ref_conv(x, w, 1, pad)
    for kw, want in ((dict(scale=scale, bias=bias, addend=addend, relu=True),
                      torch.relu(base * scale + bias + addend)),
                     (dict(addend=addend, mask=mask), (base + addend) * (mask > 0)),
                     (dict(), base)):
        plain = run(None, **kw)
        split = run(ws, **kw)
        assert torch.isfinite(split).all()
        assert rel(split, plain) <= 1e-5      # (the tensor core adds with truncation)
        assert rel(split, want) <= 1e-4
        again = run(ws, **kw)
        assert torch.equal(split, again)          # fixed summation order
    got = run(ws, round_out=True)
    assert torch.equal(got, round_tf32(got))


def test_engine_hands_out_one_split_tail_workspace_per_slot():
    """models.engine: with the K-split tail switched on the convolutions of a chain get that
    chain's workspace (E.ws_slot), results stay those of the un-split launch up to the
    summation order and are identical from slot to slot."""
    from chainer_mask_rcnn_b200.models import engine as E
    g = torch.Generator(device='cuda').manual_seed(33)
    x = round_tf32(torch.randn((430, 7, 7, 1024), device='cuda', generator=g))
    w = round_tf32(torch.randn((256, 1, 1, 1024), device='cuda', generator=g) / 32)
    old, old_ws = E._split_tail[0], dict(E._conv_ws)
    try:
        E._split_tail[0] = False
        assert E.conv_workspace(x.device) is None
        plain = E.conv_gemm(x, w, 256, relu=True, round_out=False)
        E._split_tail[0] = True
        E._conv_ws.clear()
        split0 = E.conv_gemm(x, w, 256, relu=True, round_out=False)
        with E.ws_slot(1):
            split1 = E.conv_gemm(x, w, 256, relu=True, round_out=False)
        assert sorted(k[1] for k in E._conv_ws) == [0, 1]
    finally:
        E._split_tail[0] = old
        E._conv_ws.clear()
        E._conv_ws.update(old_ws)
    assert rel(split0, plain) <= 1e-5
    assert torch.equal(split0, split1)
