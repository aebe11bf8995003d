"""ProposalCreator with chainercv's interface, one device call per batch.

Replaces ``chainercv.links.model.faster_rcnn.utils.proposal_creator.
ProposalCreator`` as constructed at chainer_mask_rcnn/models/
region_proposal_network.py:70 and called per image at :135-141.  The reference
round-trips loc/score/anchor to the host, sorts with NumPy, runs a GPU NMS kernel
plus a host sweep and copies back; here decode, filter, sort, NMS and top-k all
run in ``cmr_proposals`` (csrc/nms.cu) for the whole batch without a host sync.
"""
import torch

from . import config
from .. import _lib
from .._array import from_device, to_device


class ProposalCreator(object):

    def __init__(self, nms_thresh=0.7, n_train_pre_nms=12000, n_train_post_nms=2000,
                 n_test_pre_nms=6000, n_test_post_nms=300, force_cpu_nms=False,
                 min_size=16):
        self.nms_thresh = nms_thresh
        self.n_train_pre_nms = n_train_pre_nms
        self.n_train_post_nms = n_train_post_nms
        self.n_test_pre_nms = n_test_pre_nms
        self.n_test_post_nms = n_test_post_nms
        self.force_cpu_nms = force_cpu_nms   # accepted for signature parity; no CPU path
        self.min_size = min_size
        self._ws = None

    def budgets(self, train=None):
        if train is None:
            train = config.train
        if train:
            return self.n_train_pre_nms, self.n_train_post_nms
        return self.n_test_pre_nms, self.n_test_post_nms

    def batch(self, locs, scores, anchor, img_size, scale=1., train=None):
        """Batched device entry: locs (B, n, 4), scores (B, n), anchor (n, 4) CUDA
        tensors -> (rois (B, n_post, 4), anchor_index (B, n_post) int32,
        count (B,) int32), all on the device; rows >= count[b] are zero / -1."""
        n_pre, n_post = self.budgets(train)
        B, n_anchor = scores.shape
        if n_post <= 0:
            n_post = n_pre if n_pre > 0 else n_anchor
        n_post = min(n_post, n_anchor)
        n_pre_eff = n_anchor if (n_pre <= 0 or n_pre > n_anchor) else n_pre
        dev = scores.device
        rois = torch.empty((B, n_post, 4), dtype=torch.float32, device=dev)
        idx = torch.empty((B, n_post), dtype=torch.int32, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        ws_bytes = _lib.load().cmr_proposals_workspace_bytes(B, n_anchor, n_pre_eff)
        if self._ws is None or self._ws.numel() * 8 < ws_bytes or self._ws.device != dev:
            self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.int64, device=dev)
        _lib.call('cmr_proposals', _lib.ptr(locs), _lib.ptr(scores), _lib.ptr(anchor), B,
                  n_anchor, float(img_size[0]), float(img_size[1]),
                  float(self.min_size * scale), n_pre_eff, n_post, float(self.nms_thresh),
                  _lib.ptr(rois), _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(self._ws),
                  self._ws.numel() * 8, _lib.stream_ptr())
        return rois, idx, cnt

    def __call__(self, loc, score, anchor, img_size, scale=1., return_index=False):
        """Same contract as chainercv: loc (n, 4), score (n,), anchor (n, 4),
        img_size (H, W) -> (R, 4) float32 proposals (y1, x1, y2, x2)."""
        loc, as_np = to_device(loc, torch.float32)
        score, _ = to_device(score, torch.float32)
        anchor, _ = to_device(anchor, torch.float32)
        rois, idx, cnt = self.batch(loc[None], score.reshape(1, -1), anchor, img_size, scale)
        k = int(cnt[0].item())
        roi = from_device(rois[0, :k], as_np)
        if return_index:
            return roi, from_device(idx[0, :k], as_np)
        return roi
