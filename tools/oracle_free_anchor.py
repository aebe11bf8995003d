"""Shifted anchors for benchmarks, built with the product package only."""
import numpy as np

from chainer_mask_rcnn_b200.utils import generate_anchor_base


def anchors(fh, fw, scales, stride=16):
    base = generate_anchor_base(stride, (0.5, 1, 2), scales)
    sy = np.arange(fh, dtype=np.float32) * stride
    sx = np.arange(fw, dtype=np.float32) * stride
    shift = np.stack(np.broadcast_arrays(sy[:, None], sx[None, :], sy[:, None], sx[None, :]), axis=-1)
    return (shift.reshape(-1, 1, 4) + base[None]).reshape(-1, 4).astype(np.float32)
