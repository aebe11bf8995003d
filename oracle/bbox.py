"""NumPy restatement of the chainercv box utilities on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: chainercv
(requirements.txt:2, ``chainercv>=0.9.0``; last release 0.13.1) is not vendored
under the reference tree and cannot be installed offline, and no reference test
holds golden vectors for these functions.  Each function restates the published
chainercv 0.13 algorithm, anchored on the reference's own
call sites (and cross-checked against torchvision's CPU box_iou / nms / BoxCoder, an
independent implementation, in tests/test_oracle_bbox.py):

  generate_anchor_base      <- models/region_proposal_network.py:20-21,67-68
  enumerate_shifted_anchor  <- models/region_proposal_network.py:148-167 (in-tree)
  loc2bbox                  <- models/mask_rcnn.py:38,232
  bbox2loc, bbox_iou        <- models/utils/proposal_target_creator.py:19-20,124,156
  non_maximum_suppression   <- models/mask_rcnn.py:39,193-194 (CPU algorithm)
  ProposalCreator           <- models/region_proposal_network.py:22-23,70,136-138
  AnchorTargetCreator       <- models/mask_rcnn_train_chain.py:21-22,61,154-155

All arithmetic is fp32 in the operation order written here; boxes are
(y_min, x_min, y_max, x_max).
"""
import numpy as np

f32 = np.float32


def generate_anchor_base(base_size=16, ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32)):
    py = base_size / 2.
    px = base_size / 2.
    anchor_base = np.zeros((len(ratios) * len(anchor_scales), 4), dtype=f32)
    for i in range(len(ratios)):
        for j in range(len(anchor_scales)):
            h = base_size * anchor_scales[j] * np.sqrt(ratios[i])
            w = base_size * anchor_scales[j] * np.sqrt(1. / ratios[i])
            index = i * len(anchor_scales) + j
            anchor_base[index, 0] = py - h / 2.
            anchor_base[index, 1] = px - w / 2.
            anchor_base[index, 2] = py + h / 2.
            anchor_base[index, 3] = px + w / 2.
    return anchor_base


def enumerate_shifted_anchor(anchor_base, feat_stride, height, width):
    """region_proposal_network.py:148-167: (K*A, 4) f32, K row-major over (y, x)."""
    shift_y = np.arange(0, height * feat_stride, feat_stride)
    shift_x = np.arange(0, width * feat_stride, feat_stride)
    shift_x, shift_y = np.meshgrid(shift_x, shift_y)
    shift = np.stack((shift_y.ravel(), shift_x.ravel(),
                      shift_y.ravel(), shift_x.ravel()), axis=1)
    A = anchor_base.shape[0]
    K = shift.shape[0]
    anchor = anchor_base.reshape((1, A, 4)) + \
        shift.reshape((1, K, 4)).transpose((1, 0, 2))
    return anchor.reshape((K * A, 4)).astype(f32)


def loc2bbox(src_bbox, loc):
    if src_bbox.shape[0] == 0:
        return np.zeros((0, 4), dtype=loc.dtype)
    src_bbox = src_bbox.astype(src_bbox.dtype, copy=False)
    src_height = src_bbox[:, 2] - src_bbox[:, 0]
    src_width = src_bbox[:, 3] - src_bbox[:, 1]
    src_ctr_y = src_bbox[:, 0] + f32(0.5) * src_height
    src_ctr_x = src_bbox[:, 1] + f32(0.5) * src_width
    dy, dx, dh, dw = loc[:, 0], loc[:, 1], loc[:, 2], loc[:, 3]
    ctr_y = dy * src_height + src_ctr_y
    ctr_x = dx * src_width + src_ctr_x
    h = np.exp(dh) * src_height
    w = np.exp(dw) * src_width
    dst = np.zeros(loc.shape, dtype=loc.dtype)
    dst[:, 0] = ctr_y - f32(0.5) * h
    dst[:, 1] = ctr_x - f32(0.5) * w
    dst[:, 2] = ctr_y + f32(0.5) * h
    dst[:, 3] = ctr_x + f32(0.5) * w
    return dst


def bbox2loc(src_bbox, dst_bbox):
    height = src_bbox[:, 2] - src_bbox[:, 0]
    width = src_bbox[:, 3] - src_bbox[:, 1]
    ctr_y = src_bbox[:, 0] + f32(0.5) * height
    ctr_x = src_bbox[:, 1] + f32(0.5) * width
    base_height = dst_bbox[:, 2] - dst_bbox[:, 0]
    base_width = dst_bbox[:, 3] - dst_bbox[:, 1]
    base_ctr_y = dst_bbox[:, 0] + f32(0.5) * base_height
    base_ctr_x = dst_bbox[:, 1] + f32(0.5) * base_width
    eps = np.finfo(height.dtype).eps
    height = np.maximum(height, eps)
    width = np.maximum(width, eps)
    dy = (base_ctr_y - ctr_y) / height
    dx = (base_ctr_x - ctr_x) / width
    dh = np.log(base_height / height)
    dw = np.log(base_width / width)
    return np.vstack((dy, dx, dh, dw)).transpose()


def bbox_iou(bbox_a, bbox_b):
    tl = np.maximum(bbox_a[:, None, :2], bbox_b[:, :2])
    br = np.minimum(bbox_a[:, None, 2:], bbox_b[:, 2:])
    area_i = np.prod(br - tl, axis=2) * (tl < br).all(axis=2)
    area_a = np.prod(bbox_a[:, 2:] - bbox_a[:, :2], axis=1)
    area_b = np.prod(bbox_b[:, 2:] - bbox_b[:, :2], axis=1)
    return area_i / (area_a[:, None] + area_b - area_i)


def non_maximum_suppression(bbox, thresh, score=None, limit=None):
    """Greedy NMS, chainercv CPU algorithm.  Returns int32 indices into ``bbox``."""
    if len(bbox) == 0:
        return np.zeros((0,), dtype=np.int32)
    bbox = np.asarray(bbox, dtype=f32)
    if score is not None:
        order = score.argsort()[::-1]
        bbox = bbox[order]
    bbox_area = np.prod(bbox[:, 2:] - bbox[:, :2], axis=1)
    thresh = f32(thresh)
    selec = np.zeros(bbox.shape[0], dtype=bool)
    for i, b in enumerate(bbox):
        tl = np.maximum(b[:2], bbox[selec, :2])
        br = np.minimum(b[2:], bbox[selec, 2:])
        area = np.prod(br - tl, axis=1) * (tl < br).all(axis=1)
        with np.errstate(invalid='ignore', divide='ignore'):
            iou = area / (bbox_area[i] + bbox_area[selec] - area)
        if (iou >= thresh).any():
            continue
        selec[i] = True
        if limit is not None and np.count_nonzero(selec) >= limit:
            break
    selec = np.where(selec)[0]
    if score is not None:
        selec = order[selec]
    return selec.astype(np.int32)


def nms_suppression_bitmask(bbox, thresh):
    """The chainercv GPU kernel's intermediate: mask[i, j//64] bit j%64 set iff
    j > i and IoU(i, j) >= thresh (same fp32 IoU expression as the CPU path)."""
    n = bbox.shape[0]
    nb = (n + 63) // 64
    mask = np.zeros((n, nb), dtype=np.uint64)
    if n == 0:
        return mask
    with np.errstate(invalid='ignore', divide='ignore'):
        iou = bbox_iou(bbox, bbox)
        sup = iou >= f32(thresh)
    sup &= np.triu(np.ones((n, n), dtype=bool), k=1)
    padded = np.zeros((n, nb * 64), dtype=bool)
    padded[:, :n] = sup
    bits = padded.reshape(n, nb, 64).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    return (bits * weights).sum(axis=2).astype(np.uint64)


class ProposalCreator(object):
    """chainercv ProposalCreator (defaults of chainercv 0.13); ``train`` replaces
    the global ``chainer.config.train`` switch."""

    def __init__(self, nms_thresh=0.7, n_train_pre_nms=12000,
                 n_train_post_nms=2000, n_test_pre_nms=6000,
                 n_test_post_nms=300, force_cpu_nms=False, min_size=16):
        self.nms_thresh = nms_thresh
        self.n_train_pre_nms = n_train_pre_nms
        self.n_train_post_nms = n_train_post_nms
        self.n_test_pre_nms = n_test_pre_nms
        self.n_test_post_nms = n_test_post_nms
        self.force_cpu_nms = force_cpu_nms
        self.min_size = min_size

    def __call__(self, loc, score, anchor, img_size, scale=1., train=True,
                 return_index=False):
        if train:
            n_pre_nms, n_post_nms = self.n_train_pre_nms, self.n_train_post_nms
        else:
            n_pre_nms, n_post_nms = self.n_test_pre_nms, self.n_test_post_nms
        loc = np.asarray(loc, dtype=f32)
        score = np.asarray(score, dtype=f32)
        anchor = np.asarray(anchor, dtype=f32)
        roi = loc2bbox(anchor, loc)
        roi[:, slice(0, 4, 2)] = np.clip(roi[:, slice(0, 4, 2)], 0, img_size[0])
        roi[:, slice(1, 4, 2)] = np.clip(roi[:, slice(1, 4, 2)], 0, img_size[1])
        min_size = f32(self.min_size * scale)
        hs = roi[:, 2] - roi[:, 0]
        ws = roi[:, 3] - roi[:, 1]
        keep = np.where((hs >= min_size) & (ws >= min_size))[0]
        roi = roi[keep, :]
        score = score[keep]
        order = score.ravel().argsort()[::-1]
        if n_pre_nms > 0:
            order = order[:n_pre_nms]
        roi = roi[order, :]
        sel = non_maximum_suppression(roi, self.nms_thresh)
        if n_post_nms > 0:
            sel = sel[:n_post_nms]
        roi = roi[sel]
        if return_index:
            return roi, keep[order[sel]].astype(np.int32)
        return roi


def _unmap(data, count, index, fill=0):
    if len(data.shape) == 1:
        ret = np.empty((count,), dtype=data.dtype)
        ret.fill(fill)
        ret[index] = data
    else:
        ret = np.empty((count,) + data.shape[1:], dtype=data.dtype)
        ret.fill(fill)
        ret[index, :] = data
    return ret


class AnchorTargetCreator(object):
    """chainercv AnchorTargetCreator (0.13 defaults)."""

    def __init__(self, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3,
                 pos_ratio=0.5):
        self.n_sample = n_sample
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh = neg_iou_thresh
        self.pos_ratio = pos_ratio

    def __call__(self, bbox, anchor, img_size, rng=np.random):
        img_H, img_W = img_size
        n_anchor = len(anchor)
        inside_index = np.where(
            (anchor[:, 0] >= 0) & (anchor[:, 1] >= 0) &
            (anchor[:, 2] <= img_H) & (anchor[:, 3] <= img_W))[0]
        anchor = anchor[inside_index]
        argmax_ious, label = self._create_label(inside_index, anchor, bbox, rng)
        loc = bbox2loc(anchor, bbox[argmax_ious])
        label = _unmap(label, n_anchor, inside_index, fill=-1)
        loc = _unmap(loc, n_anchor, inside_index, fill=0)
        return loc, label

    def _create_label(self, inside_index, anchor, bbox, rng):
        label = np.empty((len(inside_index),), dtype=np.int32)
        label.fill(-1)
        ious = bbox_iou(anchor, bbox)
        argmax_ious = ious.argmax(axis=1)
        max_ious = ious[np.arange(len(inside_index)), argmax_ious]
        gt_argmax_ious = ious.argmax(axis=0)
        gt_max_ious = ious[gt_argmax_ious, np.arange(ious.shape[1])]
        gt_argmax_ious = np.where(ious == gt_max_ious)[0]
        label[max_ious < self.neg_iou_thresh] = 0
        label[gt_argmax_ious] = 1
        label[max_ious >= self.pos_iou_thresh] = 1
        n_pos = int(self.pos_ratio * self.n_sample)
        pos_index = np.where(label == 1)[0]
        if len(pos_index) > n_pos:
            disable_index = rng.choice(
                pos_index, size=(len(pos_index) - n_pos), replace=False)
            label[disable_index] = -1
        n_neg = self.n_sample - np.sum(label == 1)
        neg_index = np.where(label == 0)[0]
        if len(neg_index) > n_neg:
            disable_index = rng.choice(
                neg_index, size=(len(neg_index) - n_neg), replace=False)
            label[disable_index] = -1
        return argmax_ious, label
