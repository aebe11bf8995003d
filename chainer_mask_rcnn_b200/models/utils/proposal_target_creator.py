"""RoI-head training targets (host side).

Same interface and sampling behaviour as ``ProposalTargetCreator`` in
chainer_mask_rcnn/models/utils/proposal_target_creator.py:25-184: ground-truth boxes
are appended to the proposals, each RoI is assigned its max-IoU ground truth,
up to ``round(n_sample * pos_ratio)`` foreground RoIs (IoU >= pos_iou_thresh) and
background RoIs (IoU in [neg_iou_thresh_lo, neg_iou_thresh_hi)) are drawn with
``numpy.random.choice`` (foreground first, so a seeded run selects the same RoIs as
the reference), box targets are normalised, and every foreground RoI gets a
``mask_size`` x ``mask_size`` mask target resampled from its instance mask; background
rows are all -1 (ignored by the loss).
"""
import cv2
import numpy as np

from .bbox_tools import bbox2loc, bbox_iou


def _to_host(a):
    if hasattr(a, 'detach'):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


class ProposalTargetCreator(object):

    def __init__(self, n_sample=512, pos_ratio=0.25, pos_iou_thresh=0.5,
                 neg_iou_thresh_hi=0.5, neg_iou_thresh_lo=0.0, mask_size=14,
                 binary_thresh=0.4):
        self.n_sample = n_sample
        self.pos_ratio = pos_ratio
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh_hi = neg_iou_thresh_hi
        self.neg_iou_thresh_lo = neg_iou_thresh_lo
        self.mask_size = mask_size
        self.binary_thresh = binary_thresh   # kept for signature parity; unused upstream too

    def _mask_target(self, instance_mask, box):
        """Label map of `instance_mask` inside the integer-rounded box, resampled to
        mask_size x mask_size by bilinear interpolation of per-label indicator planes
        and an arg-max over labels."""
        y0, x0, y1, x1 = np.round(box).astype(np.int32)
        crop = instance_mask[y0:y1, x0:x1]
        size = (self.mask_size, self.mask_size)
        planes = [cv2.resize((crop == v).astype(np.float32), size)
                  for v in range(int(crop.max()) + 1)]
        return np.argmax(np.stack(planes, axis=2), axis=2).astype(np.int32)

    def __call__(self, roi, bbox, label, mask, loc_normalize_mean=(0., 0., 0., 0.),
                 loc_normalize_std=(0.1, 0.1, 0.2, 0.2)):
        # like the reference (:112-115, :179-183), the work happens on the host and the results
        # go back to where `roi` lives: NumPy in -> NumPy out, device tensor in -> device tensors
        device = roi.device if hasattr(roi, 'detach') and roi.is_cuda else None
        roi, bbox, label = _to_host(roi), _to_host(bbox), _to_host(label)
        if bbox.shape[0] == 0:
            raise ValueError('Empty bbox is not supported.')
        cand = np.concatenate((roi, bbox), axis=0)
        iou = bbox_iou(cand, bbox)
        assigned = iou.argmax(axis=1)
        best = iou.max(axis=1)

        n_pos_max = np.round(self.n_sample * self.pos_ratio)
        pos = np.flatnonzero(best >= self.pos_iou_thresh)
        n_pos = int(min(n_pos_max, pos.size))
        if pos.size > 0:
            pos = np.random.choice(pos, size=n_pos, replace=False)
        neg = np.flatnonzero((best < self.neg_iou_thresh_hi) & (best >= self.neg_iou_thresh_lo))
        n_neg = int(min(self.n_sample - n_pos, neg.size))
        if neg.size > 0:
            neg = np.random.choice(neg, size=n_neg, replace=False)

        chosen = np.append(pos, neg)
        sample_roi = cand[chosen]
        gt_roi_label = label[assigned[chosen]] + 1      # 0 is the background class
        gt_roi_label[n_pos:] = 0
        gt_roi_loc = bbox2loc(sample_roi, bbox[assigned[chosen]])
        gt_roi_loc = (gt_roi_loc - np.array(loc_normalize_mean, np.float32)) / \
            np.array(loc_normalize_std, np.float32)

        gt_roi_mask = np.full((len(sample_roi), self.mask_size, self.mask_size), -1,
                              dtype=np.int32)
        for i, p in enumerate(pos):
            gt_roi_mask[i] = self._mask_target(mask[assigned[p]], sample_roi[i])
        out = (sample_roi, gt_roi_loc, gt_roi_label, gt_roi_mask)
        if device is not None:
            import torch
            out = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in out)
        return out
