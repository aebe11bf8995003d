"""Per-channel affine ``y = W * x + b`` (frozen batch-norm replacement).

Mirrors ``chainer_mask_rcnn/functions/affine_channel_2d.py:8-66``.  Inside the
model this operator never runs on its own: it is the epilogue of the convolution
kernels (csrc/conv_*.cu).  The stand-alone function is kept for the operator
surface; it is element-wise plumbing expressed with torch on the device.
"""
import torch

from .._array import from_device, InvalidType, to_device


class AffineChannel2DFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return W * x + b

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gx = W * gy
        gW = (x * gy).sum(dim=(0, 2, 3), keepdim=True)
        gb = gy.sum(dim=(0, 2, 3), keepdim=True)
        return gx, gW, gb


def affine_channel_2d(x, W, b):
    x, as_np = to_device(x)
    W, _ = to_device(W)
    b, _ = to_device(b)
    for a in (x, W, b):
        if not a.dtype.is_floating_point or a.dim() != 4:
            raise InvalidType('affine_channel_2d expects 4-d floating arrays')
    if W.shape[1] != b.shape[1]:
        raise InvalidType('W and b must have the same number of channels')
    return from_device(AffineChannel2DFunction.apply(x, W, b), as_np)
