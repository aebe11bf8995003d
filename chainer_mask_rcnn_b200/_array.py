"""Array plumbing: torch CUDA tensors in, same kind of array out."""
import numpy as np
import torch


class InvalidType(TypeError):
    """Mirror of chainer.utils.type_check.InvalidType (dtype / ndim / shape)."""


def device():
    if not torch.cuda.is_available():
        raise RuntimeError('chainer_mask_rcnn_b200 needs a CUDA device (B200); '
                           'there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(a, dtype=None, keep_layout=False):
    """-> (contiguous CUDA tensor, was_numpy).  keep_layout: a dense tensor is returned with
    the strides it has (a channels-last feature map stays channels-last)."""
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device(), non_blocking=False)
        was_numpy = True
    elif isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(device())
        was_numpy = False
    else:
        raise InvalidType('expected numpy.ndarray or torch.Tensor, got {}'.format(type(a)))
    if dtype is not None and t.dtype != dtype:
        raise InvalidType('expected dtype {}, got {}'.format(dtype, t.dtype))
    if keep_layout and t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous():
        return t, was_numpy
    return t.contiguous(), was_numpy


def from_device(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t
