"""Training step: forward, the five losses, and the hand-scheduled backward pass.

Mirrors ``MaskRCNNTrainChain`` (chainer_mask_rcnn/models/mask_rcnn_train_chain.py:
25-189).  ``__call__(imgs, bboxes, labels, masks, scales)`` returns the scalar loss
(rpn_loc + rpn_cls + roi_loc + roi_cls + roi_mask, unweighted, :180-181) as a
:class:`Loss`; ``loss.backward()`` runs the static backward schedule and fills the
flat gradient buffer.  The two target creators stay on the host, as in the
reference (:126-158); AnchorTargetCreator runs while the GPU computes the backbone.
"""
import numpy as np
import torch

from . import engine as E
from .. import _lib
from ..utils import config
from .mask_rcnn import as_device_f32
from .utils import AnchorTargetCreator
from .utils import ProposalTargetCreator

LOSS_NAMES = ('rpn_loc_loss', 'rpn_cls_loss', 'roi_loc_loss', 'roi_cls_loss', 'roi_mask_loss')


class Loss(object):
    """Scalar loss on the device plus the pending backward schedule."""

    def __init__(self, value, backward_fn):
        self.array = value
        self._backward_fn = backward_fn

    data = property(lambda self: self.array)

    def backward(self):
        if self._backward_fn is None:
            raise RuntimeError('backward() was already called for this loss')
        fn, self._backward_fn = self._backward_fn, None
        fn()

    def item(self):
        return float(self.array.item())

    __float__ = item


def _host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


class MaskRCNNTrainChain(object):

    def __init__(self, mask_rcnn, rpn_sigma=3., roi_sigma=1., anchor_target_creator=None,
                 proposal_target_creator=None):
        self.mask_rcnn = mask_rcnn
        self.ctx = mask_rcnn.ctx
        self.rpn_sigma = rpn_sigma
        self.roi_sigma = roi_sigma
        self.anchor_target_creator = anchor_target_creator or AnchorTargetCreator()
        self.proposal_target_creator = proposal_target_creator or ProposalTargetCreator()
        self.loc_normalize_mean = mask_rcnn.loc_normalize_mean
        self.loc_normalize_std = mask_rcnn.loc_normalize_std
        self.observation = {}
        self.targets = None

    def cleargrads(self):
        self.ctx.grads.zero_()

    def __call__(self, imgs, bboxes, labels, masks, scales):
        m, ctx = self.mask_rcnn, self.ctx
        ctx.prepare(backward=True)
        x = as_device_f32(imgs)
        dev = x.device
        scales = _host(scales)
        batch_size, _, H, W = x.shape
        img_size = (H, W)
        bboxes = [_host(b).astype(np.float32) for b in bboxes]
        labels = [_host(l) for l in labels]

        ctx.recording = True
        try:
            with config.using_config('train', True):
                feat = m.extractor.forward_nhwc(x)
                rpn_locs, rpn_scores, rois, _, cnt, (anchor_np, _) = m.rpn.forward_nhwc(
                    feat, img_size, scales)
            # RPN targets need only the ground truth: computed while the GPU is busy
            gt_rpn = [self.anchor_target_creator(b, anchor_np, img_size) for b in bboxes]
            gt_rpn_locs = np.concatenate([g[0] for g in gt_rpn]).astype(np.float32)
            gt_rpn_labels = np.concatenate([g[1] for g in gt_rpn]).astype(np.int32)

            rois_h = rois.cpu().numpy()           # host sync: proposals are sampled on the host
            counts = cnt.cpu().numpy()
            s_rois, s_idx, gt_locs, gt_labels, gt_masks = [], [], [], [], []
            for i in range(batch_size):
                sr, gl, glab, gm = self.proposal_target_creator(
                    rois_h[i, :counts[i]], bboxes[i], labels[i], masks[i],
                    self.loc_normalize_mean, self.loc_normalize_std)
                s_rois.append(sr)
                s_idx.append(np.full((len(sr),), i, dtype=np.int32))
                gt_locs.append(gl)
                gt_labels.append(glab)
                gt_masks.append(gm)
            up = lambda parts, dt: torch.from_numpy(  # noqa: E731
                np.ascontiguousarray(np.concatenate(parts, axis=0), dtype=dt)).to(dev, non_blocking=True)
            sample_rois = up(s_rois, np.float32)
            sample_idx = up(s_idx, np.int32)
            gt_roi_locs = up(gt_locs, np.float32)
            gt_roi_labels = up(gt_labels, np.int32)
            gt_roi_masks = up(gt_masks, np.int32)
            gt_rpn_locs_d = torch.from_numpy(gt_rpn_locs).to(dev, non_blocking=True)
            gt_rpn_labels_d = torch.from_numpy(gt_rpn_labels).to(dev, non_blocking=True)
            self.targets = dict(sample_rois=sample_rois, sample_roi_indices=sample_idx,
                                gt_roi_locs=gt_roi_locs, gt_roi_labels=gt_roi_labels,
                                gt_roi_masks=gt_roi_masks, gt_rpn_locs=gt_rpn_locs_d,
                                gt_rpn_labels=gt_rpn_labels_d)
            return self.forward_with_targets(feat, rpn_locs, rpn_scores, **self.targets)
        finally:
            ctx.recording = False

    def forward_with_targets(self, feat, rpn_locs, rpn_scores, sample_rois, sample_roi_indices,
                             gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
                             gt_rpn_labels):
        """Head forward + losses for given samples/targets (everything on the device)."""
        m, ctx = self.mask_rcnn, self.ctx
        dev = feat.device
        head, rpn = m.head, m.rpn
        cls_locs, scores, masks = head.forward_nhwc(feat, sample_rois, sample_roi_indices)
        R = cls_locs.shape[0]
        n, hh, ww, _ = feat.shape
        A = rpn.n_anchor
        losses = torch.zeros((8,), dtype=torch.float32, device=dev)
        g_rpn = torch.empty((n, hh, ww, rpn.g_ld), dtype=torch.float32, device=dev)
        g_lin = torch.empty((R, head.lin_ld), dtype=torch.float32, device=dev)
        g_mask = torch.empty((R, 14, 14, head.mask_ld), dtype=torch.float32, device=dev)
        st = E.stream()
        _lib.call('cmr_rpn_loss', E._p(rpn_locs), 4 * A, E._p(rpn_scores), A, E._p(gt_rpn_locs),
                  E._p(gt_rpn_labels), n * hh * ww, A, float(self.rpn_sigma), E._p(g_rpn),
                  rpn.g_ld, E._p(losses), st)
        _lib.call('cmr_roi_loss', E._p(cls_locs), 4 * head.n_class, E._p(scores), head.n_class,
                  E._p(gt_roi_locs), E._p(gt_roi_labels), R, head.n_class, float(self.roi_sigma),
                  E._p(g_lin), head.lin_ld, E._p(losses), st)
        _lib.call('cmr_mask_loss', E._p(masks), head.mask_ld, E._p(gt_roi_labels),
                  E._p(gt_roi_masks), R, 14 * 14, head.n_fg, E._p(g_mask), head.mask_ld,
                  E._p(losses), st)
        loss = losses[:5].sum()
        self.observation = {k: losses[i] for i, k in enumerate(LOSS_NAMES)}
        self.observation['loss'] = loss
        self.outputs = dict(rpn_locs=rpn_locs, rpn_scores=rpn_scores, roi_cls_locs=cls_locs,
                            roi_scores=scores, roi_masks=masks)

        def backward():
            g_feat = head.backward(g_lin, g_mask)
            g_feat = rpn.backward(g_rpn, g_feat)
            m.extractor.backward(g_feat)

        return Loss(loss, backward)
