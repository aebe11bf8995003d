"""``concat_examples`` with the reference's signature and batch format.

Mirrors ``chainer_mask_rcnn/datasets/concat_examples.py:6-34`` (which builds on
``chainer.dataset.convert._concat_arrays``): element ``i`` of every example is stacked into
one array when ``i`` is in ``indices_concat`` -- arrays of different shapes are placed in
the top-left corner of an array of the largest shape filled with ``padding[i]`` -- and stays
a list otherwise; elements in ``indices_to_device`` are sent to ``device``.  The training
script uses ``padding=0, indices_concat=[0, 2, 3, 4], indices_to_device=[0, 1]``
(examples/train_common.py): images (B,3,H,W) float32 on the device, boxes a list of device
arrays, labels (B,G) int32, masks (B,G,H,W) int32 and scales (B,) on the host.

B200 additions, both optional:
  pinned=True   stacked host arrays live in page-locked memory (NumPy views of pinned torch
                tensors), so ``optimizers.GraphedUpdater`` uploads them asynchronously while
                the previous iteration computes.  The 341 MB int32 mask block of a COCO batch
                is the one input whose copy is worth hiding.
  canvas=(H,W)  spatial axes of image-like elements (ndim >= 3) are padded up to this fixed
                size instead of the batch maximum, so every batch has one geometry and the
                updater replays one CUDA graph.
"""
import numpy as np
import torch


def pinned_empty(shape, dtype):
    """NumPy array of page-locked host memory (a view of a pinned torch tensor, which it
    keeps alive)."""
    t = torch.empty(tuple(int(s) for s in shape), dtype=_torch_dtype(dtype))
    if torch.cuda.is_available():
        t = t.pin_memory()
    return t.numpy()


def _torch_dtype(dtype):
    return torch.from_numpy(np.empty((0,), dtype=np.bool_ if dtype == bool else dtype)).dtype


def _concat_arrays(arrays, padding, pinned, canvas):
    arrays = [np.asarray(a) for a in arrays]
    first = arrays[0]
    same = all(a.shape == first.shape for a in arrays)
    shape = np.array(first.shape, dtype=int)
    for a in arrays[1:]:
        if a.ndim != first.ndim:
            raise ValueError('arrays of a batch element must have the same ndim')
        shape = np.maximum(shape, a.shape)
    if canvas is not None and first.ndim >= 3:
        if shape[-2] > canvas[0] or shape[-1] > canvas[1]:
            raise ValueError('canvas {} is smaller than an example {}'.format(
                tuple(canvas), tuple(shape[-2:])))
        shape[-2:] = canvas
        same = same and tuple(first.shape[-2:]) == tuple(canvas)
    if not same and padding is None:
        raise ValueError('arrays of different shapes need a padding value')
    full = (len(arrays),) + tuple(int(s) for s in shape)
    out = pinned_empty(full, first.dtype) if pinned else np.empty(full, first.dtype)
    if same:
        for i, a in enumerate(arrays):
            out[i] = a
        return out
    out[...] = padding
    for i, a in enumerate(arrays):
        out[(i,) + tuple(slice(0, s) for s in a.shape)] = a
    return out


def _to_device(device, a):
    if device is None or (isinstance(device, int) and device < 0):
        return a
    dev = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dev, non_blocking=True)


def concat_examples(batch, device=None, padding=None, indices_concat=None,
                    indices_to_device=None, pinned=False, canvas=None):
    if len(batch) == 0:
        raise ValueError('batch is empty')
    elem_size = len(batch[0])
    if indices_concat is None:
        indices_concat = range(elem_size)
    if indices_to_device is None:
        indices_to_device = range(elem_size)
    if not isinstance(padding, tuple):
        padding = [padding] * elem_size
    result = []
    for i in range(elem_size):
        res = [example[i] for example in batch]
        if i in indices_concat:
            res = _concat_arrays(res, padding[i], pinned, canvas)
        if i in indices_to_device:
            if i in indices_concat:
                res = _to_device(device, res)
            else:
                res = [_to_device(device, r) for r in res]
        result.append(res)
    return tuple(result)
