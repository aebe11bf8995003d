"""RPN training targets (host side).

Mirrors ``chainercv.links.model.faster_rcnn.AnchorTargetCreator`` as constructed at
chainer_mask_rcnn/models/mask_rcnn_train_chain.py:61 and called per image at
:151-158.  Host-side NumPy in the reference too; it only depends on the ground
truth and the anchor grid, so the train chain runs it while the GPU is busy with
the backbone.  Sampling draws from ``numpy.random`` (``rng``) in the same order as
the reference: positives first, then negatives.
"""
import numpy as np

from .bbox_tools import bbox2loc, bbox_iou


class AnchorTargetCreator(object):

    def __init__(self, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3, pos_ratio=0.5):
        self.n_sample = n_sample
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh = neg_iou_thresh
        self.pos_ratio = pos_ratio

    def __call__(self, bbox, anchor, img_size, rng=None):
        """bbox (R,4), anchor (S,4) -> loc (S,4) float32, label (S,) int32 in {-1,0,1}."""
        rng = np.random if rng is None else rng
        bbox = np.asarray(bbox, dtype=np.float32)
        anchor = np.asarray(anchor, dtype=np.float32)
        img_h, img_w = img_size
        n_anchor = anchor.shape[0]
        inside = np.flatnonzero((anchor[:, 0] >= 0) & (anchor[:, 1] >= 0) &
                                (anchor[:, 2] <= img_h) & (anchor[:, 3] <= img_w))
        cand = anchor[inside]
        iou = bbox_iou(cand, bbox)
        best_gt = iou.argmax(axis=1)
        best_iou = iou[np.arange(len(cand)), best_gt]
        per_gt_best = iou.max(axis=0)

        lab = np.full((len(cand),), -1, dtype=np.int32)
        lab[best_iou < self.neg_iou_thresh] = 0
        lab[np.where(iou == per_gt_best)[0]] = 1      # every anchor tying a gt's best IoU
        lab[best_iou >= self.pos_iou_thresh] = 1

        max_pos = int(self.pos_ratio * self.n_sample)
        pos = np.flatnonzero(lab == 1)
        if len(pos) > max_pos:
            lab[rng.choice(pos, size=len(pos) - max_pos, replace=False)] = -1
        max_neg = self.n_sample - int(np.sum(lab == 1))
        neg = np.flatnonzero(lab == 0)
        if len(neg) > max_neg:
            lab[rng.choice(neg, size=len(neg) - max_neg, replace=False)] = -1

        label = np.full((n_anchor,), -1, dtype=np.int32)
        label[inside] = lab
        loc = np.zeros((n_anchor, 4), dtype=np.float32)
        loc[inside] = bbox2loc(cand, bbox[best_gt])
        return loc, label
