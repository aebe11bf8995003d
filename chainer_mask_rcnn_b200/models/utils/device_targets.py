"""Target creation on the device (csrc/targets.cu): the train step's default.

``AnchorTargetCreator`` / ``ProposalTargetCreator`` (host NumPy, this directory) keep
the reference's exact sampling order under a NumPy seed; the classes here implement the
same assignment rules and sample sizes in CUDA for the whole batch at once, with a
counter-based hash instead of NumPy's global generator, so that the step has no NumPy
pass between the RPN and the RoI head.  The mask-target rasterisation
(proposal_target_creator.py:163-177) runs on the device too (``cmr_mask_targets``) when
the instance masks are given as a device tensor; with host NumPy masks it stays on the
host (cv2, as in the reference) and overlaps with the head's forward pass.
"""
import ctypes

import cv2
import numpy as np
import torch

from ... import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class GroundTruth(object):
    """Per-batch ground truth packed for the device: bbox (B,G,4) f32, label (B,G) i32,
    count (B,) i32, G = max boxes per image."""

    MAX_BOXES = 256

    def __init__(self, bboxes, labels, device, capacity=None):
        B = len(bboxes)
        G = capacity or max(1, max(len(b) for b in bboxes))
        if G > self.MAX_BOXES:
            raise ValueError('at most 256 ground-truth boxes per image are supported')
        self.B, self.G = B, G
        self.nbytes = (B * G * 5 + B) * 4
        # one buffer = one H2D copy; the device copy has a fixed address, so a captured
        # CUDA graph keeps reading it after fill_() has been called with new boxes
        self._host = torch.empty((B * G * 5 + B,), dtype=torch.int32).pin_memory()
        self._buf = torch.empty((B * G * 5 + B,), dtype=torch.int32, device=device)
        self.bbox = self._buf[:B * G * 4].view(torch.float32).view(B, G, 4)
        self.label = self._buf[B * G * 4:B * G * 5].view(B, G)
        self.count = self._buf[B * G * 5:]
        self._copied = None
        self.fill_(bboxes, labels)

    def fill_(self, bboxes, labels):
        B, G = self.B, self.G
        if len(bboxes) != B:
            raise ValueError('expected {} images, got {}'.format(B, len(bboxes)))
        if self._copied is not None:
            self._copied.synchronize()     # the previous H2D copy has read the staging buffer
        packed = self._host.numpy()
        packed[:] = 0
        bb = packed[:B * G * 4].view(np.float32).reshape(B, G, 4)
        ll = packed[B * G * 4:B * G * 5].reshape(B, G)
        cc = packed[B * G * 5:]
        for i, (b, l) in enumerate(zip(bboxes, labels)):
            n = len(b)
            if n == 0:
                raise ValueError('Empty bbox is not supported.')
            if n > G:
                raise ValueError('{} boxes exceed the capacity {}'.format(n, G))
            bb[i, :n] = b
            ll[i, :n] = np.asarray(l)[:n]
            cc[i] = n
        self._buf.copy_(self._host, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        return self


class PackedMasks(object):
    """Binary instance masks packed one bit per pixel: ``data`` is a (B,G,H,ceil(W/8)) uint8
    torch tensor (``numpy.packbits(masks, axis=-1, bitorder='little')``), ``width`` the
    image width W.  8x fewer bytes to move host -> device than uint8 masks."""

    def __init__(self, data, width):
        if data.dtype != torch.uint8 or data.dim() != 4 or data.shape[3] != (width + 7) // 8:
            raise TypeError('PackedMasks needs a (B,G,H,ceil(W/8)) uint8 tensor')
        self.data, self.width = data, int(width)

    @classmethod
    def from_numpy(cls, masks, pin=False):
        """masks: (B,G,H,W) array-like of {0,1}."""
        m = np.asarray(masks)
        t = torch.from_numpy(np.packbits(m.astype(bool), axis=-1, bitorder='little'))
        return cls(t.pin_memory() if pin else t, m.shape[-1])

    def to(self, device, non_blocking=False):
        return PackedMasks(self.data.to(device, non_blocking=non_blocking), self.width)

    is_cuda = property(lambda self: self.data.is_cuda)
    nbytes = property(lambda self: self.data.numel())


class DeviceAnchorTargetCreator(object):

    def __init__(self, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3, pos_ratio=0.5):
        self.n_sample = n_sample
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh = neg_iou_thresh
        self.pos_ratio = pos_ratio
        self._ws = None

    def __call__(self, gt, anchor, img_size, seed, seed_dev=None):
        """gt: GroundTruth; anchor (S,4) CUDA tensor -> gt_loc (B,S,4), gt_label (B,S).
        seed_dev: optional device int64 word mixed into the seed (see cmr_anchor_targets)."""
        S = anchor.shape[0]
        dev = anchor.device
        loc = torch.empty((gt.B, S, 4), dtype=torch.float32, device=dev)
        label = torch.empty((gt.B, S), dtype=torch.int32, device=dev)
        nbytes = _lib.load().cmr_anchor_targets_workspace_bytes(gt.B, S, gt.G)
        if self._ws is None or self._ws.numel() * 8 < nbytes or self._ws.device != dev:
            self._ws = torch.empty(((nbytes + 7) // 8,), dtype=torch.int64, device=dev)
        _lib.call('cmr_anchor_targets', _p(anchor), S, _p(gt.bbox), _p(gt.count), gt.B, gt.G,
                  float(img_size[0]), float(img_size[1]), self.n_sample,
                  float(self.pos_iou_thresh), float(self.neg_iou_thresh), float(self.pos_ratio),
                  int(seed) & (2 ** 64 - 1), None if seed_dev is None else _p(seed_dev), _p(loc),
                  _p(label), _p(self._ws), self._ws.numel() * 8, _stream())
        return loc, label


class DeviceProposalTargetCreator(object):

    def __init__(self, n_sample=512, pos_ratio=0.25, pos_iou_thresh=0.5, neg_iou_thresh_hi=0.5,
                 neg_iou_thresh_lo=0.0, mask_size=14, binary_thresh=0.4):
        self.n_sample = n_sample
        self.pos_ratio = pos_ratio
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh_hi = neg_iou_thresh_hi
        self.neg_iou_thresh_lo = neg_iou_thresh_lo
        self.mask_size = mask_size
        self.binary_thresh = binary_thresh
        self._ws = None

    def sample(self, rois, n_roi, gt, seed, loc_normalize_mean=(0., 0., 0., 0.),
               loc_normalize_std=(0.1, 0.1, 0.2, 0.2), seed_dev=None):
        """rois (B,max_roi,4), n_roi (B,) as returned by ProposalCreator.batch.
        -> sample_roi (B,n,4), gt_roi_loc (B,n,4), gt_roi_label (B,n), gt_assign (B,n),
        n_pos (B,), all on the device, n = n_sample."""
        B, max_roi, _ = rois.shape
        dev = rois.device
        n = self.n_sample
        sample_roi = torch.empty((B, n, 4), dtype=torch.float32, device=dev)
        gt_loc = torch.empty((B, n, 4), dtype=torch.float32, device=dev)
        gt_label = torch.empty((B, n), dtype=torch.int32, device=dev)
        gt_assign = torch.empty((B, n), dtype=torch.int32, device=dev)
        n_pos = torch.empty((B,), dtype=torch.int32, device=dev)
        nbytes = _lib.load().cmr_proposal_targets_workspace_bytes(B, max_roi, gt.G)
        if self._ws is None or self._ws.numel() * 8 < nbytes or self._ws.device != dev:
            self._ws = torch.empty(((nbytes + 7) // 8,), dtype=torch.int64, device=dev)
        mean = (ctypes.c_float * 4)(*loc_normalize_mean)
        std = (ctypes.c_float * 4)(*loc_normalize_std)
        _lib.call('cmr_proposal_targets', _p(rois), _p(n_roi), max_roi, _p(gt.bbox), _p(gt.label),
                  _p(gt.count), B, gt.G, n, float(self.pos_ratio), float(self.pos_iou_thresh),
                  float(self.neg_iou_thresh_hi), float(self.neg_iou_thresh_lo), mean, std,
                  int(seed) & (2 ** 64 - 1), None if seed_dev is None else _p(seed_dev),
                  _p(sample_roi), _p(gt_loc), _p(gt_label), _p(gt_assign), _p(n_pos),
                  _p(self._ws), self._ws.numel() * 8, _stream())
        return sample_roi, gt_loc, gt_label, gt_assign, n_pos

    def mask_targets_device(self, sample_roi, gt_assign, n_pos, masks):
        """Device side: masks (B,G,H,W) uint8 or int32 CUDA tensor of instance masks ->
        (B,n,mask_size,mask_size) int32 targets, -1 on every non-foreground row
        (cmr_mask_targets)."""
        elem = None
        if isinstance(masks, PackedMasks):
            masks, W, elem = masks.data, masks.width, 0
        if masks.dtype not in (torch.uint8, torch.int32) or masks.dim() != 4:
            raise TypeError('masks must be a (B,G,H,W) uint8 or int32 tensor, got {} {}'.format(
                masks.dtype, tuple(masks.shape)))
        B, n, _ = sample_roi.shape
        Bm, G, H, Wt = masks.shape
        if elem is None:
            W, elem = Wt, masks.element_size()
        if Bm != B:
            raise ValueError('masks has {} images, rois {}'.format(Bm, B))
        ms = self.mask_size
        out = torch.empty((B, n, ms, ms), dtype=torch.int32, device=sample_roi.device)
        _lib.call('cmr_mask_targets', _p(masks.contiguous()), elem, B, G, H, W,
                  _p(sample_roi), _p(gt_assign), _p(n_pos), n, ms, _p(out), _stream())
        return out

    def mask_targets(self, sample_roi, gt_assign, n_pos, masks):
        """Host side: (B,n,4) rois, (B,n) assignments, (B,) counts (NumPy) and the per-image
        instance masks -> (B,n,mask_size,mask_size) int32, -1 on background rows.  Same
        arithmetic as the host ProposalTargetCreator."""
        B, n, _ = sample_roi.shape
        ms = self.mask_size
        out = np.full((B, n, ms, ms), -1, dtype=np.int32)
        for b in range(B):
            boxes = np.round(sample_roi[b, :n_pos[b]]).astype(np.int32)
            for i in range(int(n_pos[b])):
                y0, x0, y1, x1 = boxes[i]
                crop = masks[b][gt_assign[b, i]][y0:y1, x0:x1]
                top = int(crop.max()) if crop.size else 0
                if top == 0:
                    out[b, i] = 0
                    continue
                planes = [cv2.resize((crop == v).astype(np.float32), (ms, ms))
                          for v in range(top + 1)]
                out[b, i] = np.argmax(np.stack(planes, axis=2), axis=2)
        return out
