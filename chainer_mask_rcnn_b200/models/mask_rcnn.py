"""Mask R-CNN base class: extractor -> rpn -> head.

Mirrors ``MaskRCNN`` (chainer_mask_rcnn/models/mask_rcnn.py:110-176): ``__call__``
(:142-150) and ``prepare`` (:152-176).  ``predict`` and its CPU post-processing
(:178-337) are "next" rows of the scope table and are not part of
this path yet.
"""
import cv2
import numpy as np
import torch

from . import engine as E
from .region_proposal_network import flatten_proposals


def as_device_f32(x):
    """numpy / torch (any device) -> contiguous float32 CUDA tensor."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('chainer_mask_rcnn_b200 needs a CUDA device (B200); there is no '
                               'CPU fallback')
        x = x.cuda(non_blocking=True)
    if x.dtype != torch.float32:
        raise TypeError('expected float32, got {}'.format(x.dtype))
    return x.contiguous()


class MaskRCNN(object):

    def __init__(self, extractor, rpn, head, mean, min_size=600, max_size=1000,
                 loc_normalize_mean=(0., 0., 0., 0.), loc_normalize_std=(0.1, 0.1, 0.2, 0.2),
                 detections_per_im=100):
        self.extractor = extractor
        self.rpn = rpn
        self.head = head
        self.mean = mean
        self.min_size = min_size
        self.max_size = max_size
        self.loc_normalize_mean = loc_normalize_mean
        self.loc_normalize_std = loc_normalize_std
        self.nms_thresh = 0.5
        self.score_thresh = 0.05
        self._detections_per_im = detections_per_im

    @property
    def n_class(self):
        """Total number of classes including the background."""
        return self.head.n_class

    def __call__(self, x, scales):
        """x (B,3,H,W) float32, scales (B,) -> roi_cls_locs (R',4*n_class),
        roi_scores (R',n_class), rois (R',4), roi_indices (R',), roi_masks
        (R',n_fg,14,14); every proposal goes through the head."""
        ctx = self.ctx
        ctx.prepare(backward=False)
        x = as_device_f32(x)
        img_size = tuple(x.shape[2:])
        feat = self.extractor.forward_nhwc(x)
        _, _, rois, _, cnt, _ = self.rpn.forward_nhwc(feat, img_size, np.asarray(scales))
        rois, roi_indices = flatten_proposals(rois, cnt)
        cls_locs, scores, masks = self.head.forward_nhwc(feat, rois, roi_indices)
        return cls_locs, scores, rois, roi_indices, E.as_nchw_view(masks)

    def prepare(self, imgs):
        """Resize so that the short side is min_size (capped by max_size on the long
        side) and subtract the mean; host-side, as in the reference."""
        prepared, sizes, scales = [], [], []
        for img in imgs:
            _, H, W = img.shape
            scale = 1.
            if self.min_size:
                scale = self.min_size / min(H, W)
            if self.max_size and scale * max(H, W) > self.max_size:
                scale = self.max_size / max(H, W)
            out = cv2.resize(img.transpose(1, 2, 0), None, fx=scale, fy=scale)
            out = (out.transpose(2, 0, 1) - self.mean).astype(np.float32, copy=False)
            prepared.append(out)
            sizes.append((H, W))
            scales.append(scale)
        return prepared, sizes, scales

    # ---- parameters, reference (Chainer npz) naming and layouts ----
    def namedparams(self):
        for name in self.ctx.names():
            yield '/' + name, self.ctx.to_reference(name)

    def state_dict(self):
        return {name: self.ctx.to_reference(name).detach().cpu().numpy().copy()
                for name in self.ctx.names()}

    def load_state_dict(self, params, strict=True):
        names = set(self.ctx.names())
        for name, value in params.items():
            name = name.lstrip('/')
            if name not in names:
                if strict:
                    raise KeyError('unexpected parameter ' + name)
                continue
            self.ctx.set_from_reference(name, value)
            names.discard(name)
        if strict and names:
            raise KeyError('missing parameters: ' + ', '.join(sorted(names)[:5]))

    def save_npz(self, path):
        np.savez(path, **self.state_dict())

    def load_npz(self, path, strict=True):
        with np.load(path) as f:
            self.load_state_dict({k: f[k] for k in f.files}, strict)
