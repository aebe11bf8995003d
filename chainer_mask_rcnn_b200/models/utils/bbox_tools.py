"""Host-side box arithmetic used by the target creators (NumPy, fp32,
boxes are (y_min, x_min, y_max, x_max)).

These mirror the chainercv helpers the reference imports at
chainer_mask_rcnn/models/utils/proposal_target_creator.py:18-20 and
models/mask_rcnn.py:38 (``bbox2loc``, ``bbox_iou``, ``loc2bbox``).  The target
creators are host-side glue in the reference as well;
the device versions of decode/IoU live in csrc/nms.cu.
"""
import numpy as np


def _hw_center(box):
    h = box[:, 2] - box[:, 0]
    w = box[:, 3] - box[:, 1]
    return h, w, box[:, 0] + 0.5 * h, box[:, 1] + 0.5 * w


def bbox_iou(bbox_a, bbox_b):
    """(N,4), (K,4) -> (N,K) IoU; areas without the +1 pixel convention."""
    if bbox_a.shape[1] != 4 or bbox_b.shape[1] != 4:
        raise IndexError('boxes must have 4 columns')
    top_left = np.maximum(bbox_a[:, None, :2], bbox_b[None, :, :2])
    bottom_right = np.minimum(bbox_a[:, None, 2:], bbox_b[None, :, 2:])
    overlap = (top_left < bottom_right).all(axis=2)
    inter = np.prod(bottom_right - top_left, axis=2) * overlap
    area_a = np.prod(bbox_a[:, 2:] - bbox_a[:, :2], axis=1)
    area_b = np.prod(bbox_b[:, 2:] - bbox_b[:, :2], axis=1)
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def bbox2loc(src_bbox, dst_bbox):
    """Offsets (dy, dx, dh, dw) that move ``src_bbox`` onto ``dst_bbox``."""
    sh, sw, scy, scx = _hw_center(src_bbox)
    dh, dw, dcy, dcx = _hw_center(dst_bbox)
    eps = np.finfo(sh.dtype).eps
    sh = np.maximum(sh, eps)
    sw = np.maximum(sw, eps)
    return np.stack(((dcy - scy) / sh, (dcx - scx) / sw, np.log(dh / sh), np.log(dw / sw)),
                    axis=1)


def loc2bbox(src_bbox, loc):
    """Inverse of :func:`bbox2loc`."""
    if src_bbox.shape[0] == 0:
        return np.zeros((0, 4), dtype=loc.dtype)
    sh, sw, scy, scx = _hw_center(src_bbox.astype(loc.dtype, copy=False))
    cy = loc[:, 0] * sh + scy
    cx = loc[:, 1] * sw + scx
    h = np.exp(loc[:, 2]) * sh
    w = np.exp(loc[:, 3]) * sw
    half = loc.dtype.type(0.5)
    return np.stack((cy - half * h, cx - half * w, cy + half * h, cx + half * w), axis=1)
