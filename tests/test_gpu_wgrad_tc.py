"""tcgen05 weight-gradient kernel (MN-major operands) vs fp32 autograd of the same op.
Operands are pre-rounded to TF32 so only the accumulation order differs."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from chainer_mask_rcnn_b200 import _lib
from test_gpu_conv_tc import rel, round_tf32

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tma_im2col', 'cp_async'], autouse=True)
def operand_path(request):
    """Every case runs on the im2col-TMA kernel (default) and on the cp.async fallback."""
    lib = _lib.load()
    old = lib.cmr_set_im2col_tma(1 if request.param == 'tma_im2col' else 0)
    yield request.param
    lib.cmr_set_im2col_tma(old)


def wgrad_conv(x, gy, kh, kw, stride, pad, row_scale=None, splits=0):
    """x (B,H,W,C), gy (B,oh,ow,N) NHWC -> gW (N, kh, kw, C)."""
    B, H, W, C = x.shape
    _, oh, ow, N = gy.shape
    gw = torch.zeros((N, kh, kw, C), device='cuda')
    for fr in range(kh):
        for fs in range(kw):
            d = _lib.WgradDesc(B, oh, ow, oh, ow, N, 1, 0, 0, 0, H, W, C, stride, fr - pad,
                               fs - pad, 0, N, C, kh * kw * C, (fr * kw + fs) * C, splits, 1, 1)
            _lib.call('cmr_conv_wgrad_tc', ctypes.byref(d), _lib.ptr(gy), _lib.ptr(x),
                      _lib.ptr(gw), _lib.ptr(row_scale), _lib.stream_ptr())
    return gw


def ref_wgrad(x, gy, kh, kw, stride, pad):
    B, H, W, C = x.shape
    N = gy.shape[3]
    w = torch.zeros((N, C, kh, kw), device='cuda', dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w, stride=stride, padding=pad)
    y.backward(gy.permute(0, 3, 1, 2).double())
    return w.grad.permute(0, 2, 3, 1).float().contiguous()


CASES = [
    # B, H,  W,   C,   N, k, s, p, splits
    (1, 4, 8, 64, 128, 1, 1, 0, 1),        # one k-block, one tile
    (2, 13, 17, 128, 128, 1, 1, 0, 1),     # ragged pixel count, several k-blocks
    (2, 13, 17, 128, 256, 1, 1, 0, 0),     # two row tiles, automatic split
    (2, 20, 23, 128, 128, 3, 1, 1, 0),     # 3x3 pad 1 (nine taps)
    (2, 21, 19, 256, 128, 1, 2, 0, 3),     # 1x1 stride 2, explicit 3-way split
    (16, 7, 7, 512, 512, 3, 1, 1, 0),      # res5 conv2 shape (fewer RoIs)
    (1, 25, 42, 1024, 76, 1, 1, 0, 0),     # RPN loc+score rows = 76 (ragged row tile)
    (64, 1, 1, 2048, 408, 1, 1, 0, 0),     # Linear layers fused (rows 408), 64 RoIs
    (3, 14, 18, 192, 64, 3, 2, 1, 0),      # 3x3 stride 2 pad 1, cols = 192 (ragged 256 tile)
    (2, 51, 84, 256, 256, 1, 1, 0, 0),     # res4 1x1: many splits of a 256-wide tile
]


@pytest.mark.parametrize('B,H,W,C,N,k,s,p,splits', CASES)
def test_wgrad_matches_autograd(B, H, W, C, N, k, s, p, splits):
    g = torch.Generator(device='cuda').manual_seed(B + H * 7 + C)
    oh = (H + 2 * p - k) // s + 1
    ow = (W + 2 * p - k) // s + 1
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    gy = round_tf32(torch.randn((B, oh, ow, N), device='cuda', generator=g))
    want = ref_wgrad(x, gy, k, k, s, p)
    got = wgrad_conv(x, gy, k, k, s, p, splits=splits)
    assert rel(got, want) <= 1e-4


@pytest.mark.parametrize('B,H,W,C,N,k,p', [(2, 20, 23, 128, 128, 3, 1), (16, 7, 7, 512, 512, 3, 1),
                                           (2, 13, 17, 64, 96, 3, 1), (1, 9, 9, 64, 64, 7, 3)])
def test_all_taps_in_one_launch(B, H, W, C, N, k, p):
    g = torch.Generator(device='cuda').manual_seed(C + k)
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    gy = round_tf32(torch.randn((B, H, W, N), device='cuda', generator=g))
    want = ref_wgrad(x, gy, k, k, 1, p)
    gw = torch.zeros((N, k, k, C), device='cuda')
    d = _lib.WgradDesc(B, H, W, H, W, N, 1, 0, 0, 0, H, W, C, 1, -p, -p, 0, N, C, k * k * C, 0, 0,
                       k, k)
    _lib.call('cmr_conv_wgrad_tc', ctypes.byref(d), _lib.ptr(gy), _lib.ptr(x), _lib.ptr(gw), None,
              _lib.stream_ptr())
    assert rel(gw, want) <= 1e-4


def test_row_scale_and_accumulation():
    g = torch.Generator(device='cuda').manual_seed(3)
    x = round_tf32(torch.randn((2, 9, 11, 128), device='cuda', generator=g))
    gy = round_tf32(torch.randn((2, 9, 11, 128), device='cuda', generator=g))
    scale = torch.rand((128,), device='cuda', generator=g) + 0.5
    want = ref_wgrad(x, gy, 1, 1, 1, 0) * scale.view(-1, 1, 1, 1)
    got = wgrad_conv(x, gy, 1, 1, 1, 0, row_scale=scale)
    assert rel(got, want) <= 1e-4
    # gw is accumulated into: a second call doubles it
    d = _lib.WgradDesc(2, 9, 11, 9, 11, 128, 1, 0, 0, 0, 9, 11, 128, 1, 0, 0, 0, 128, 128, 128, 0, 0, 1, 1)
    _lib.call('cmr_conv_wgrad_tc', ctypes.byref(d), _lib.ptr(gy), _lib.ptr(x), _lib.ptr(got),
              _lib.ptr(scale), _lib.stream_ptr())
    assert rel(got, 2 * want) <= 1e-4


def test_deconv_tap_weight_gradient():
    """Deconvolution2D(k=2, stride=2): gW[tap][o][c] = sum gy[2y+dy, 2x+dx, o] * x[y, x, c]."""
    g = torch.Generator(device='cuda').manual_seed(4)
    B, H, W, C, N = 8, 7, 7, 256, 128
    x = round_tf32(torch.randn((B, H, W, C), device='cuda', generator=g))
    gy = round_tf32(torch.randn((B, 2 * H, 2 * W, N), device='cuda', generator=g))
    wt = torch.zeros((C, N, 2, 2), device='cuda', dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt, stride=2)
    y.backward(gy.permute(0, 3, 1, 2).double())
    want = wt.grad.permute(2, 3, 1, 0).float().contiguous()      # (dy, dx, o, c)
    got = torch.zeros((2, 2, N, C), device='cuda')
    for dy in range(2):
        for dx in range(2):
            d = _lib.WgradDesc(B, H, W, 2 * H, 2 * W, N, 2, dy, dx, 0, H, W, C, 1, 0, 0, 0,
                               N, C, C, 0, 0, 1, 1)
            _lib.call('cmr_conv_wgrad_tc', ctypes.byref(d), _lib.ptr(gy), _lib.ptr(x),
                      _lib.ptr(got[dy, dx]), None, _lib.stream_ptr())
    assert rel(got, want) <= 1e-4
