"""CPU restatement of the inference post-processing of ``MaskRCNN.predict``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows chainer_mask_rcnn/models/mask_rcnn.py:
  to_bboxes      <- MaskRCNN._to_bboxes (:203-261) with MaskRCNN._suppress (:178-201)
  segm_results   <- segm_results / expand_boxes (:44-107)
  sigmoid/softmax as chainer.functions (fp32: exp(x - max) / sum; 1 / (1 + exp(-x)))

PARITY UNPINNED for the pieces that live in chainercv (``loc2bbox``,
``non_maximum_suppression``; oracle/bbox.py) -- the reference has no test or golden
vector for predict.  ``cv2.resize`` in ``segm_results`` is restated
by oracle/mask_target.py (pinned bit-exact against cv2 without IPP).

Reference quirk kept on purpose (:255-260): the ``detections_per_im`` cut keeps the
positions ``k`` with ``argsort(score)[k] >= n - D`` -- a positional test on the argsort
*values*, not the D best scores.
"""
import numpy as np

from . import bbox as ob
from . import mask_target as omt

f32 = np.float32


def softmax(x):
    x = np.asarray(x, f32)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(f32)


def sigmoid(x):
    x = np.asarray(x, f32)
    return (f32(1) / (f32(1) + np.exp(-x))).astype(f32)


def suppress(raw_cls_bbox, raw_prob, n_class, score_thresh, nms_thresh):
    """Per-class score threshold + NMS (:178-201).  Classes are visited in order, each
    class contributes its survivors in descending score order."""
    bbox, label, score = [], [], []
    boxes = raw_cls_bbox.reshape((-1, n_class, 4))
    for l in range(1, n_class):
        cls_bbox_l = boxes[:, l, :]
        prob_l = raw_prob[:, l]
        keep = prob_l > score_thresh
        cls_bbox_l = cls_bbox_l[keep]
        prob_l = prob_l[keep]
        keep = ob.non_maximum_suppression(cls_bbox_l, nms_thresh, prob_l)
        bbox.append(cls_bbox_l[keep])
        label.append((l - 1) * np.ones((len(keep),)))
        score.append(prob_l[keep])
    return (np.concatenate(bbox, axis=0).astype(f32), np.concatenate(label, axis=0).astype(np.int32),
            np.concatenate(score, axis=0).astype(f32))


def decode_class_boxes(roi_cls_loc, roi, n_class, size, mean, std):
    """(:223-238) un-normalise the offsets, loc2bbox against the RoI for every class,
    clip to the image.  roi is already divided by the image scale."""
    # the reference tiles the Python-float tuples: float64 arithmetic, rounded once
    mean = np.tile(np.asarray(mean), n_class)
    std = np.tile(np.asarray(std), n_class)
    loc = (roi_cls_loc * std + mean).astype(f32).reshape((-1, n_class, 4))
    roi_cls = np.broadcast_to(roi[:, None], loc.shape)
    cls_bbox = ob.loc2bbox(roi_cls.reshape((-1, 4)), loc.reshape((-1, 4)))
    cls_bbox = cls_bbox.reshape((-1, n_class * 4))
    cls_bbox[:, 0::2] = np.clip(cls_bbox[:, 0::2], 0, size[0])
    cls_bbox[:, 1::2] = np.clip(cls_bbox[:, 1::2], 0, size[1])
    return cls_bbox


def cut_detections(bbox, label, score, detections_per_im):
    """(:245-260) drop boxes whose rounded area is 0, then the positional cut."""
    bbox_int = np.round(bbox).astype(np.int32)
    sizes = (bbox_int[:, 2] - bbox_int[:, 0]) * (bbox_int[:, 3] - bbox_int[:, 1])
    keep = sizes > 0
    bbox, label, score = bbox[keep], label[keep], score[keep]
    if detections_per_im > 0:
        indices = np.argsort(score)
        keep = indices >= (len(indices) - detections_per_im)
        bbox, label, score = bbox[keep], label[keep], score[keep]
    return bbox, label, score


def to_bboxes(roi_cls_locs, roi_scores, rois, roi_indices, sizes, scales, n_class,
              loc_normalize_mean=(0., 0., 0., 0.), loc_normalize_std=(0.1, 0.1, 0.2, 0.2),
              score_thresh=0.05, nms_thresh=0.5, detections_per_im=100):
    probs = softmax(roi_scores)
    bboxes, labels, scores = [], [], []
    for index in range(len(sizes)):
        keep = roi_indices == index
        roi = (rois[keep] / f32(scales[index])).astype(f32)
        cls_bbox = decode_class_boxes(roi_cls_locs[keep], roi, n_class, sizes[index],
                                      loc_normalize_mean, loc_normalize_std)
        bbox, label, score = suppress(cls_bbox, probs[keep], n_class, score_thresh, nms_thresh)
        bbox, label, score = cut_detections(bbox, label, score, detections_per_im)
        bboxes.append(bbox)
        labels.append(label)
        scores.append(score)
    return bboxes, labels, scores


def expand_boxes(boxes, scale):
    """(:44-60) boxes (x1,y1,x2,y2) grown by `scale` around their centres (float64)."""
    w_half = (boxes[:, 2] - boxes[:, 0]) * .5 * scale
    h_half = (boxes[:, 3] - boxes[:, 1]) * .5 * scale
    x_c = (boxes[:, 2] + boxes[:, 0]) * .5
    y_c = (boxes[:, 3] + boxes[:, 1]) * .5
    out = np.zeros(boxes.shape)
    out[:, 0] = x_c - w_half
    out[:, 2] = x_c + w_half
    out[:, 1] = y_c - h_half
    out[:, 3] = y_c + h_half
    return out


def segm_results(bbox, label, roi_mask, im_h, im_w):
    """(:63-107) paste every detection's (sigmoid) mask of its class into the image:
    pad to (M+2)^2, resize to the integer box grown by (M+2)/M, threshold at 0.5."""
    if len(bbox) == 0:
        return np.zeros((0, im_h, im_w), dtype=bool)
    M = roi_mask.shape[2]
    ref_boxes = expand_boxes(bbox[:, [1, 0, 3, 2]], (M + 2.0) / M).astype(np.int32)
    padded = np.zeros((M + 2, M + 2), dtype=f32)
    out = np.zeros((len(bbox), im_h, im_w), dtype=bool)
    for k in range(len(ref_boxes)):
        padded[1:-1, 1:-1] = roi_mask[k, label[k]]
        x0, y0, x1, y1 = ref_boxes[k]
        w = max(x1 - x0 + 1, 1)
        h = max(y1 - y0 + 1, 1)
        mask = omt.resize_linear_f32(padded, h, w) > 0.5
        xa, xb = max(x0, 0), min(x1 + 1, im_w)
        ya, yb = max(y0, 0), min(y1 + 1, im_h)
        if xb > xa and yb > ya:
            out[k, ya:yb, xa:xb] = mask[ya - y0:yb - y0, xa - x0:xb - x0]
    return out
