"""Box utilities with chainercv's interface, running in libcmr_b200.

``non_maximum_suppression`` replaces ``chainercv.utils.non_maximum_suppression``
as used at chainer_mask_rcnn/models/mask_rcnn.py:39,193-194;
``generate_anchor_base`` replaces the chainercv helper imported at
models/region_proposal_network.py:20-21 (init-time host arithmetic).
"""
import numpy as np
import torch

from .. import _lib
from .._array import from_device, to_device


def generate_anchor_base(base_size=16, ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32)):
    """(len(ratios) * len(anchor_scales), 4) float32 anchors (y1, x1, y2, x2) around
    the centre of a ``base_size`` cell; index = i_ratio * len(scales) + j_scale."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(anchor_scales, dtype=np.float64)
    h = (base_size * scales[None, :] * np.sqrt(ratios)[:, None]).ravel()
    w = (base_size * scales[None, :] * np.sqrt(1. / ratios)[:, None]).ravel()
    c = base_size / 2.
    return np.stack([c - h / 2., c - w / 2., c + h / 2., c + w / 2.], axis=1).astype(np.float32)


def _nms_device(bbox, thresh, limit, want_mask=False):
    n = bbox.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=bbox.device)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=bbox.device)
    ws_bytes = _lib.load().cmr_nms_workspace_bytes(n)
    ws = torch.empty((ws_bytes // 8,), dtype=torch.int64, device=bbox.device)
    if want_mask:
        ws.zero_()
    _lib.call('cmr_nms', _lib.ptr(bbox), n, float(thresh), int(limit or 0), _lib.ptr(keep),
              _lib.ptr(n_keep), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    k = int(n_keep.item())
    if want_mask:
        nb = (n + 63) // 64
        return keep[:k], ws[:n * nb].view(n, nb)
    return keep[:k]


def non_maximum_suppression(bbox, thresh, score=None, limit=None):
    """Greedy NMS.  ``bbox`` (n, 4) float32 (y1, x1, y2, x2).  Boxes are visited in
    the given order, or by descending ``score`` when it is passed; a box is dropped
    when its IoU with an already selected box is ``>= thresh``.  Returns the int32
    indices of the selected boxes (indices into the input order)."""
    bbox, as_np = to_device(bbox, torch.float32)
    if bbox.shape[0] == 0:
        out = torch.zeros((0,), dtype=torch.int32, device=bbox.device)
        return from_device(out, as_np)
    order = None
    if score is not None:
        score, _ = to_device(score)
        order = torch.argsort(score, descending=True, stable=True)
        bbox = bbox[order].contiguous()
    keep = _nms_device(bbox, thresh, limit)
    if order is not None:
        keep = order[keep.long()].to(torch.int32)
    return from_device(keep, as_np)


def nms_suppression_bitmask(bbox, thresh):
    """(keep, mask): the keep list and the (n, ceil(n/64)) uint64 suppression bitmask
    (as int64 bit patterns) the kernel built -- exposed for the bit-exact tests."""
    bbox, as_np = to_device(bbox, torch.float32)
    keep, mask = _nms_device(bbox, thresh, None, want_mask=True)
    if as_np:
        return keep.cpu().numpy(), mask.cpu().numpy().view(np.uint64)
    return keep, mask
