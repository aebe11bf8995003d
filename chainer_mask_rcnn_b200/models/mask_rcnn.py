"""Mask R-CNN base class: extractor -> rpn -> head, and inference.

Mirrors ``MaskRCNN`` (chainer_mask_rcnn/models/mask_rcnn.py:110-337): ``__call__``
(:142-150), ``prepare`` (:152-176) and ``predict`` (:307-337) with its post-processing
``_to_bboxes`` / ``_suppress`` (:178-261), ``_to_roi_masks`` (:263-287), ``_to_masks``
(:289-305) and the module-level ``segm_results`` / ``expand_boxes`` (:44-107).

The reference post-processes on the host: one NumPy NMS per class and image, one
``cv2.resize`` per detection.  Here the class-wise NMS of a batch is one device call
(``cmr_detections``) and the masks are pasted by ``cmr_paste_masks``; only the final
detections_per_im cut (a few hundred rows, with the reference's positional argsort
rule) runs in NumPy.
"""
import ctypes
import weakref

import cv2
import numpy as np
import torch

from . import engine as E
from .. import _lib
from ..utils import config
from .region_proposal_network import flatten_proposals


def expand_boxes(boxes, scale):
    """Expand an array of (x1, y1, x2, y2) boxes by a given scale (mask_rcnn.py:44-60)."""
    w_half = (boxes[:, 2] - boxes[:, 0]) * .5 * scale
    h_half = (boxes[:, 3] - boxes[:, 1]) * .5 * scale
    x_c = (boxes[:, 2] + boxes[:, 0]) * .5
    y_c = (boxes[:, 3] + boxes[:, 1]) * .5
    out = np.zeros(boxes.shape)
    out[:, 0], out[:, 2] = x_c - w_half, x_c + w_half
    out[:, 1], out[:, 3] = y_c - h_half, y_c + h_half
    return out


def _paste(bbox_dev, label_dev, roi_mask, im_h, im_w, apply_sigmoid):
    """bbox (n,4) f32 / label (n,) i32 device tensors, roi_mask (n, n_fg, ms, ms)-shaped
    device tensor (any strides) -> (n, im_h, im_w) uint8 device tensor."""
    n = bbox_dev.shape[0]
    out = torch.empty((n, im_h, im_w), dtype=torch.uint8, device=bbox_dev.device)
    if n:
        sn, sc, sy, sx = roi_mask.stride()
        _lib.call('cmr_paste_masks', E._p(bbox_dev), E._p(label_dev), E._p(roi_mask), sn, sc, sy,
                  sx, n, roi_mask.shape[2], im_h, im_w, int(apply_sigmoid), E._p(out), E.stream())
    return out


def segm_results(bbox, label, roi_mask, im_h, im_w):
    """Paste per-detection mask probabilities into image-sized boolean masks
    (mask_rcnn.py:63-107).  bbox (n,4) (y1,x1,y2,x2), label (n,), roi_mask
    (n, n_fg, M, M) probabilities -> (n, im_h, im_w) bool."""
    if len(bbox) == 0:
        return np.zeros((0, im_h, im_w), dtype=bool)
    assert roi_mask.shape[3] == roi_mask.shape[2]
    dev = torch.device('cuda', torch.cuda.current_device())
    b = torch.from_numpy(np.ascontiguousarray(bbox, np.float32)).to(dev)
    l = torch.from_numpy(np.ascontiguousarray(label, np.int32)).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(roi_mask, np.float32)).to(dev)
    return _paste(b, l, m, im_h, im_w, False).cpu().numpy().astype(bool)


def prepare_scale(H, W, min_size, max_size):
    """The resize factor of ``prepare`` (mask_rcnn.py:158-165), Python-float arithmetic."""
    scale = 1.
    if min_size:
        scale = min_size / min(H, W)
    if max_size and scale * max(H, W) > max_size:
        scale = max_size / max(H, W)
    return scale


def prepared_size(H, W, scale):
    """(h, w) of cv2.resize(img, None, fx=scale, fy=scale): cvRound of the products."""
    h, w = ctypes.c_int(), ctypes.c_int()
    _lib.call('cmr_prepare_size', int(H), int(W), float(scale), float(scale),
              ctypes.addressof(h), ctypes.addressof(w))
    return h.value, w.value


class _PinnedDownloads(object):
    """Page-locked destination buffers for the (n, H, W) mask stacks predict returns.

    ~100 MB per image leave the device in the reference's output format; into pageable
    memory that copy runs at a fraction of the link rate.  The arrays handed to the caller
    are views of page-locked tensors from torch's caching host allocator (a buffer goes
    back to the cache when the caller drops the array, so a predict loop reuses the same
    few buffers).  At most `budget` bytes are outstanding at a time; beyond that (a caller
    who keeps every result) downloads go to ordinary pageable memory.
    """

    def __init__(self, budget=1 << 30):
        self.budget = budget
        self.outstanding = 0

    def _release(self, nbytes):
        self.outstanding -= nbytes

    def download(self, dev_tensor):
        nbytes = dev_tensor.numel() * dev_tensor.element_size()
        pin = self.outstanding + nbytes <= self.budget
        host = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=pin)
        host.copy_(dev_tensor)
        arr = host.numpy()       # shares (and keeps alive) the tensor's storage
        if pin:
            self.outstanding += nbytes
            weakref.finalize(arr, self._release, nbytes)   # views hold `arr` as their base
        return arr


_downloads = _PinnedDownloads()


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
            scale = prepare_scale(H, W, self.min_size, self.max_size)
            out = cv2.resize(img.transpose(1, 2, 0), None, fx=scale, fy=scale)
            out = (out.transpose(2, 0, 1) - self.mean).astype(np.float32, copy=False)
            prepared.append(out)
            sizes.append((H, W))
            scales.append(scale)
        return prepared, sizes, scales

    def _prepare_device(self, imgs):
        """prepare + concat_examples(padding=0) of predict (mask_rcnn.py:308-311) on the
        device: the raw float32 images are uploaded as they are and cmr_prepare_image
        writes the resized, mean-subtracted, zero-padded (B, 3, Hm, Wm) batch.
        -> x (device), sizes, scales."""
        sizes, scales, shapes = [], [], []
        for img in imgs:
            _, H, W = img.shape
            scale = prepare_scale(H, W, self.min_size, self.max_size)
            sizes.append((H, W))
            scales.append(scale)
            shapes.append(prepared_size(H, W, scale))
        Hm, Wm = max(s[0] for s in shapes), max(s[1] for s in shapes)
        mean = np.asarray(self.mean, np.float32).reshape(-1)
        dev = torch.device('cuda', torch.cuda.current_device())
        x = torch.empty((len(imgs), 3, Hm, Wm), dtype=torch.float32, device=dev)
        for i, img in enumerate(imgs):
            raw = as_device_f32(img)
            _lib.call('cmr_prepare_image', E._p(raw), sizes[i][0], sizes[i][1], float(scales[i]),
                      float(scales[i]), float(mean[0]), float(mean[1]), float(mean[2]),
                      E._p(x[i]), Hm, Wm, E.stream())
            raw.record_stream(torch.cuda.current_stream())
        return x, sizes, scales

    # ---- inference (mask_rcnn.py:178-337) ----
    def _forward_padded(self, x, scales, pred_mask):
        """extractor -> rpn -> head over the padded (B, n_post) proposal table (no host
        sync on the proposal counts; rows past the count are all-zero boxes)."""
        self.ctx.prepare(backward=False)
        x = as_device_f32(x)
        feat = self.extractor.forward_nhwc(x)
        _, _, rois, _, cnt, _ = self.rpn.forward_nhwc(feat, tuple(x.shape[2:]), np.asarray(scales))
        B, n_post, _ = rois.shape
        idx = torch.arange(B, dtype=torch.int32, device=x.device).repeat_interleave(n_post)
        cls_locs, scores, masks = self.head.forward_nhwc(feat, rois.view(-1, 4), idx,
                                                         pred_mask=pred_mask)
        return feat, rois, cnt, cls_locs, scores, masks

    def _detect(self, cls_locs, scores, rois, cnt, sizes, scales):
        """Device half of _to_bboxes: -> per image (bbox, label, score) NumPy arrays in the
        reference's order (class by class, descending score inside a class), before the
        rounded-area filter and the detections_per_im cut."""
        B, max_roi, _ = rois.shape
        C = self.n_class
        dev = rois.device
        max_cand = min(max_roi * (C - 1), 19 * max_roi)     # prob > 0.05 holds for < 20 classes
        if self.score_thresh < 0.05:
            max_cand = max_roi * (C - 1)
        max_cand = (max_cand + 63) // 64 * 64
        info = np.array([[scales[i], sizes[i][0], sizes[i][1]] for i in range(B)], np.float32)
        info = torch.from_numpy(info).to(dev)
        det_bbox = torch.empty((B, max_cand, 4), dtype=torch.float32, device=dev)
        det_label = torch.empty((B, max_cand), dtype=torch.int32, device=dev)
        det_score = torch.empty((B, max_cand), dtype=torch.float32, device=dev)
        n_det = torch.empty((B,), dtype=torch.int32, device=dev)
        lib = _lib.load()
        ws_bytes = lib.cmr_detections_workspace_bytes(B, max_roi, C, max_cand)
        ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
        mean = (ctypes.c_double * 4)(*self.loc_normalize_mean)
        std = (ctypes.c_double * 4)(*self.loc_normalize_std)
        _lib.call('cmr_detections', E._p(cls_locs), cls_locs.stride(0), E._p(scores),
                  scores.stride(0), E._p(rois), E._p(cnt), B, max_roi, C, E._p(info), mean, std,
                  float(self.score_thresh), float(self.nms_thresh), max_cand, E._p(det_bbox),
                  E._p(det_label), E._p(det_score), E._p(n_det), E._p(ws), ws.numel() * 8,
                  E.stream())
        n_det = n_det.cpu().numpy()                        # the one host sync
        out = []
        for i in range(B):
            n = int(n_det[i])
            bbox = det_bbox[i, :n].cpu().numpy()
            label = det_label[i, :n].cpu().numpy()
            score = det_score[i, :n].cpu().numpy()
            out.append((bbox, label, score))
        return out

    def _to_bboxes(self, roi_cls_locs, roi_scores, rois, roi_indices, sizes, scales):
        """Reference signature (mask_rcnn.py:203): concatenated (R', ...) arrays with
        roi_indices sorted by image -> lists of (bbox, label, score) per image."""
        cls_locs = as_device_f32(roi_cls_locs)
        scores = as_device_f32(roi_scores)
        rois = as_device_f32(rois)
        idx = roi_indices.cpu().numpy() if isinstance(roi_indices, torch.Tensor) \
            else np.asarray(roi_indices)
        B = len(sizes)
        counts = np.bincount(idx, minlength=B).astype(np.int32)
        max_roi = max(int(counts.max()), 1)
        dev = rois.device
        pr = torch.zeros((B, max_roi, 4), dtype=torch.float32, device=dev)
        pl = torch.zeros((B, max_roi, cls_locs.shape[1]), dtype=torch.float32, device=dev)
        ps = torch.zeros((B, max_roi, scores.shape[1]), dtype=torch.float32, device=dev)
        start = 0
        for i, c in enumerate(counts):
            pr[i, :c], pl[i, :c], ps[i, :c] = (t[start:start + c] for t in (rois, cls_locs, scores))
            start += c
        cnt = torch.from_numpy(counts).to(dev)
        dets = self._detect(pl.view(B * max_roi, -1), ps.view(B * max_roi, -1), pr, cnt, sizes,
                            scales)
        return self._cut(dets)

    def _cut(self, dets):
        """Host tail of _to_bboxes (mask_rcnn.py:245-260): drop boxes whose rounded area is
        zero, then the detections_per_im cut with the reference's positional rule."""
        bboxes, labels, scores = [], [], []
        for bbox, label, score in dets:
            bbox_int = np.round(bbox).astype(np.int32)
            keep = (bbox_int[:, 2] - bbox_int[:, 0]) * (bbox_int[:, 3] - bbox_int[:, 1]) > 0
            bbox, label, score = bbox[keep], label[keep], score[keep]
            if self._detections_per_im > 0:
                indices = np.argsort(score)
                keep = indices >= (len(indices) - self._detections_per_im)
                bbox, label, score = bbox[keep], label[keep], score[keep]
            bboxes.append(bbox)
            labels.append(label)
            scores.append(score)
        return bboxes, labels, scores

    def _to_roi_masks(self, feat, bboxes, roi_indices, scales):
        """Second head pass on the detected boxes (mask_rcnn.py:263-287) -> list of
        (n_i, n_fg, M, M)-shaped device tensors of mask logits (channels-last views)."""
        B = feat.shape[0]
        n_fg, M = self.n_class - 1, self.head.mask_size
        allb = np.concatenate(bboxes, axis=0)
        if allb.size == 0:
            return [torch.zeros((0, n_fg, M, M), dtype=torch.float32, device=feat.device)
                    for _ in range(B)]
        # float32 boxes times the float64 scales concat_examples returns, rounded once to
        # float32 (mask_rcnn.py:281-282)
        scales = np.asarray(scales, np.float64)
        rois = torch.from_numpy((allb * scales[roi_indices][:, None]).astype(np.float32))
        rois = rois.to(feat.device)
        idx = torch.from_numpy(np.asarray(roi_indices, np.int32)).to(feat.device)
        _, _, masks = self.head.forward_nhwc(feat, rois, idx, pred_bbox=False)
        masks = E.as_nchw_view(masks)
        ends = np.cumsum([len(b) for b in bboxes])       # roi_indices are sorted by image
        return [masks[e - len(b):e] for b, e in zip(bboxes, ends)]

    def _to_masks(self, bboxes, labels, scores, roi_masks, sizes):
        """sigmoid + segm_results per image (mask_rcnn.py:289-305) -> list of (n_i, H, W)
        bool arrays."""
        outs = []
        for bbox, label, roi_mask, size in zip(bboxes, labels, roi_masks, sizes):
            if len(bbox) == 0:
                outs.append(np.zeros((0, size[0], size[1]), dtype=bool))
                continue
            dev = roi_mask.device
            b = torch.from_numpy(np.ascontiguousarray(bbox, np.float32)).to(dev)
            l = torch.from_numpy(np.ascontiguousarray(label, np.int32)).to(dev)
            outs.append(_paste(b, l, roi_mask, size[0], size[1], True))
        # the 0/1 bytes are downloaded straight into the arrays that are returned (a bool
        # view of them, no astype copy): ~100 MB per image at the reference's output format
        return [_downloads.download(o).view(np.bool_) if isinstance(o, torch.Tensor) else o
                for o in outs]

    def predict(self, imgs):
        """imgs: list of (3, H, W) float32 RGB arrays in [0, 255].
        -> bboxes, masks, labels, scores (lists per image; mask_rcnn.py:307-337)."""
        x, sizes, scales = self._prepare_device(imgs)
        scales = np.asarray(scales, np.float64)        # as concat_examples returns them
        with config.using_config('train', False), torch.no_grad():
            feat, rois, cnt, cls_locs, scores, _ = self._forward_padded(x, scales, False)
            bboxes, labels, scores = self._cut(
                self._detect(cls_locs, scores, rois, cnt, sizes, scales))
            roi_indices = np.concatenate(
                [np.full((len(b),), i, np.int32) for i, b in enumerate(bboxes)])
            roi_masks = self._to_roi_masks(feat, bboxes, roi_indices, scales)
            masks = self._to_masks(bboxes, labels, scores, roi_masks, sizes)
        return bboxes, masks, labels, scores

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
