"""Per-channel affine ``y = W * x + b`` (frozen batch-norm replacement).

Mirrors ``chainer_mask_rcnn/functions/affine_channel_2d.py:8-66``: class
``AffineChannel2DFunction`` (forward :17-20, backward :48-55) and the wrapper
``affine_channel_2d(x, W, b)`` with ``W`` and ``b`` shaped (1, C, 1, 1).  Inside the model
this operator never runs on its own -- it is the epilogue of the convolution kernels
(csrc/conv_tc.cu); the stand-alone operator runs ``cmr_affine_channel_fwd / _bwd``
(csrc/affine.cu): the forward rounds like NumPy's ``W * x + b`` (bit-exact), the backward
writes ``gx = W * gy`` and reduces ``gW = sum(x * gy)``, ``gb = sum(gy)`` over (n, h, w) in
one pass over ``x`` and ``gy`` with a fixed-order (deterministic) second stage.
"""
import torch

from .. import _lib
from .._array import from_device, InvalidType, to_device


def _check(x, W, b):
    for a in (x, W, b):
        if a.dtype != torch.float32 or a.dim() != 4:
            raise InvalidType('affine_channel_2d expects 4-d float32 arrays, got {} {}'.format(
                a.dtype, tuple(a.shape)))
    C = x.shape[1]
    if tuple(W.shape) != (1, C, 1, 1) or tuple(b.shape) != (1, C, 1, 1):
        raise InvalidType('W and b must be shaped (1, {}, 1, 1), got {} and {}'.format(
            C, tuple(W.shape), tuple(b.shape)))


def _forward(x, W, b):
    N, C, H, Wd = x.shape
    y = torch.empty_like(x)
    _lib.call('cmr_affine_channel_fwd', _lib.ptr(x), _lib.ptr(W), _lib.ptr(b), N, C, H * Wd,
              _lib.ptr(y), _lib.stream_ptr())
    return y


def _backward(x, W, gy):
    N, C, H, Wd = x.shape
    gx = torch.empty_like(x)
    gW = torch.empty((1, C, 1, 1), dtype=torch.float32, device=x.device)
    gb = torch.empty((1, C, 1, 1), dtype=torch.float32, device=x.device)
    if N == 0:
        return gx, gW.zero_(), gb.zero_()
    nbytes = _lib.load().cmr_affine_channel_bwd_workspace_bytes(N, C)
    ws = torch.empty((max(nbytes // 4, 1),), dtype=torch.float32, device=x.device)
    _lib.call('cmr_affine_channel_bwd', _lib.ptr(x), _lib.ptr(W), _lib.ptr(gy), N, C, H * Wd,
              _lib.ptr(gx), _lib.ptr(gW), _lib.ptr(gb), _lib.ptr(ws), ws.numel() * 4,
              _lib.stream_ptr())
    return gx, gW, gb


class AffineChannel2DFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return _forward(x, W, b)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        return _backward(x, W, gy.contiguous())


def affine_channel_2d(x, W, b):
    """y = W * x + b for x (N, C, H, W), W and b (1, C, 1, 1); float32 torch CUDA tensors
    (differentiable with respect to all three) or NumPy arrays (copied to the GPU and
    back)."""
    x, as_np = to_device(x)
    W, _ = to_device(W)
    b, _ = to_device(b)
    _check(x, W, b)
    return from_device(AffineChannel2DFunction.apply(x, W, b), as_np)
