"""Training step: forward, the five losses, and the hand-scheduled backward pass.

Mirrors ``MaskRCNNTrainChain`` (chainer_mask_rcnn/models/mask_rcnn_train_chain.py:
25-189).  ``__call__(imgs, bboxes, labels, masks, scales)`` returns the scalar loss
(rpn_loc + rpn_cls + roi_loc + roi_cls + roi_mask, unweighted, :180-181) as a
:class:`Loss`; ``loss.backward()`` runs the static backward schedule and fills the
flat gradient buffer.

Targets.  By default anchors and proposals are labelled / sampled on the device
(models/utils/device_targets.py, csrc/targets.cu).  Mask targets are rasterised on the
device as well when ``masks`` is one (B,G,H,W) array -- the reference's own batch format
(a host NumPy int32 array from ``datasets.concat_examples``), a torch tensor (uint8 / int32,
CUDA or pinned host) or a bit-packed ``models.utils.PackedMasks`` -- the step then has no
host synchronisation at all and can be captured in a CUDA graph
(optimizers.GraphedUpdater); a list of per-image mask arrays is rasterised on the host
(cv2, as in the reference), overlapped with the head's forward pass.  Passing the host ``AnchorTargetCreator`` / ``ProposalTargetCreator``
objects instead reproduces the reference's NumPy-seeded sampling exactly (:126-158).
"""
import numpy as np
import torch

from . import engine as E
from .. import _lib
from ..utils import config
from .mask_rcnn import as_device_f32
from .utils import DeviceAnchorTargetCreator
from .utils import DeviceProposalTargetCreator
from .utils import GroundTruth
from .utils import PackedMasks

LOSS_NAMES = ('rpn_loc_loss', 'rpn_cls_loss', 'roi_loc_loss', 'roi_cls_loss', 'roi_mask_loss')


class Loss(object):
    """Scalar loss on the device plus the pending backward schedule."""

    def __init__(self, value, backward_fn):
        self.array = value
        self._backward_fn = backward_fn

    data = property(lambda self: self.array)

    supports_progress = True

    def backward(self, after_head=None, progress=None):
        """Runs the backward schedule.  ``after_head``: optional callable invoked once the
        gradients of the RPN and RoI-head parameters are final (every kernel that writes them
        has been enqueued and joined onto the current stream) and before the backbone's
        backward pass is enqueued -- the data-parallel optimizer starts the all-reduce of
        that bucket there, under the backbone's backward pass.  ``progress(block)``: called
        after each backbone block's backward pass has been enqueued (the optimizer cuts the
        backbone's gradient exchange into pieces there)."""
        if self._backward_fn is None:
            raise RuntimeError('backward() was already called for this loss')
        fn, self._backward_fn = self._backward_fn, None
        fn(after_head, progress)

    def item(self):
        return float(self.array.item())

    __float__ = item


def _host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


class MaskRCNNTrainChain(object):

    def __init__(self, mask_rcnn, rpn_sigma=3., roi_sigma=1., anchor_target_creator=None,
                 proposal_target_creator=None, seed=0, deterministic=None):
        """``deterministic=True`` (or CMR_DETERMINISTIC=1): the reductions whose arrival order
        is not fixed (split weight gradients, ROIAlign backward, bias sums) accumulate in 64-bit
        fixed point, so two runs of the same steps give bit-identical parameters; ~3 % slower."""
        import os
        self.mask_rcnn = mask_rcnn
        self.ctx = mask_rcnn.ctx
        if deterministic is None:
            deterministic = os.environ.get('CMR_DETERMINISTIC', '0') == '1'
        self.ctx.deterministic = bool(deterministic)
        self.rpn_sigma = rpn_sigma
        self.roi_sigma = roi_sigma
        self.anchor_target_creator = anchor_target_creator or DeviceAnchorTargetCreator()
        self.proposal_target_creator = proposal_target_creator or DeviceProposalTargetCreator()
        self.loc_normalize_mean = mask_rcnn.loc_normalize_mean
        self.loc_normalize_std = mask_rcnn.loc_normalize_std
        self.observation = {}
        self.targets = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._seed = int(seed)
        self.seed_dev = None            # optional device int64 word mixed into the seeds
        self._calls = 0
        self._pinned = {}
        self._roi_index = {}
        # The RPN branch (anchor targets -> RPN losses [-> RPN backward]) depends on the RPN
        # convolutions only; the proposal chain that follows them on the main stream (decode,
        # sort, NMS sweep: ~0.5 ms of one-CTA-per-image, latency-bound kernels) leaves the
        # SMs idle.  The branch is forked onto a side stream there.  Its backward half
        # (loc / score / conv1 weight gradients and conv1's data gradient, ~0.5 ms of
        # tensor-core work) runs ahead too when `eager_rpn_backward` is set -- the optimizer
        # entry points set it, because they always call loss.backward() on zeroed gradients.
        self.overlap_rpn_branch = True
        self.eager_rpn_backward = False
        self._rpn_stream = None

    def cleargrads(self):
        self.ctx.grads.zero_()

    # ------------------------------------------------------------------ call --
    def __call__(self, imgs, bboxes, labels, masks, scales):
        m, ctx = self.mask_rcnn, self.ctx
        ctx.prepare(backward=True)
        self.h2d_bytes = self.d2h_bytes = 0
        if not (isinstance(imgs, torch.Tensor) and imgs.is_cuda):
            self.h2d_bytes += int(np.prod(imgs.shape)) * 4
        x = as_device_f32(imgs)
        dev = x.device
        scales = _host(scales)
        batch_size, _, H, W = x.shape
        img_size = (H, W)
        gt = bboxes if isinstance(bboxes, GroundTruth) else None
        if gt is None:
            bboxes = [_host(b).astype(np.float32) for b in bboxes]
            labels = [_host(l) for l in labels]
        self._calls += 1
        dev_anchor = isinstance(self.anchor_target_creator, DeviceAnchorTargetCreator)
        dev_prop = isinstance(self.proposal_target_creator, DeviceProposalTargetCreator)
        if gt is not None and not (dev_anchor and dev_prop):
            raise TypeError('a packed GroundTruth needs the device target creators')
        if gt is None and (dev_anchor or dev_prop):
            gt = GroundTruth(bboxes, labels, dev)
            self.h2d_bytes += gt.nbytes
        seed = (self._seed << 20) + 2 * self._calls
        ctx.recording = True
        try:
            branch = [None]

            def fork_rpn_branch(rpn_locs, rpn_scores, anchor):
                branch[0] = self._start_rpn_branch(feat, rpn_locs, rpn_scores, anchor, gt,
                                                   img_size, seed)

            with config.using_config('train', True):
                feat = m.extractor.forward_nhwc(x)
                rpn_locs, rpn_scores, rois, _, cnt, (anchor_np, anchor) = m.rpn.forward_nhwc(
                    feat, img_size, scales,
                    between=fork_rpn_branch if dev_anchor and self.overlap_rpn_branch else None)
            masks_dev = None
            if isinstance(masks, np.ndarray) and masks.ndim == 4 and dev_prop and \
                    masks.dtype in (np.int32, np.uint8, np.bool_):
                # the reference's batch format (datasets.concat_examples: (B,G,H,W) int32 on
                # the host, models/mask_rcnn_train_chain.py:76-111): uploaded as it is and
                # rasterised on the device -- bit-identical to the host cv2 path
                masks = torch.from_numpy(masks.view(np.uint8) if masks.dtype == np.bool_
                                         else masks)
            if isinstance(masks, (torch.Tensor, PackedMasks)):
                if not dev_prop:
                    raise TypeError('tensor masks need the device proposal target creator')
                if not masks.is_cuda:
                    self.h2d_bytes += masks.numel() * masks.element_size() \
                        if isinstance(masks, torch.Tensor) else masks.nbytes
                masks_dev = masks.to(dev, non_blocking=True)
            # ---- RPN targets
            if branch[0] is not None:
                gt_rpn_locs, gt_rpn_labels = branch[0]['gt_rpn_locs'], branch[0]['gt_rpn_labels']
            elif dev_anchor:
                gt_rpn_locs, gt_rpn_labels = self.anchor_target_creator(
                    gt, anchor, img_size, seed, seed_dev=self.seed_dev)
                gt_rpn_locs = gt_rpn_locs.view(-1, 4)
                gt_rpn_labels = gt_rpn_labels.view(-1)
            else:
                pairs = [self.anchor_target_creator(b, anchor_np, img_size) for b in bboxes]
                gt_rpn_locs = self._upload(np.concatenate([p[0] for p in pairs]), np.float32, dev)
                gt_rpn_labels = self._upload(np.concatenate([p[1] for p in pairs]), np.int32, dev)
            # ---- RoI sampling + head targets
            if dev_prop:
                ptc = self.proposal_target_creator
                sroi, gloc, glab, gasg, npos = ptc.sample(rois, cnt, gt, seed + 1,
                                                          self.loc_normalize_mean,
                                                          self.loc_normalize_std,
                                                          seed_dev=self.seed_dev)
                n = sroi.shape[1]
                max_pos = int(np.round(ptc.n_sample * ptc.pos_ratio))
                key = (batch_size, n, str(dev))
                if key not in self._roi_index:
                    self._roi_index[key] = torch.arange(batch_size, dtype=torch.int32, device=dev) \
                        .repeat_interleave(n)
                sample_rois = sroi.view(-1, 4)
                sample_idx = self._roi_index[key]
                gt_roi_locs, gt_roi_labels = gloc.view(-1, 4), glab.view(-1)
            if dev_prop and masks_dev is not None:
                gt_roi_masks = ptc.mask_targets_device(sroi, gasg, npos, masks_dev) \
                    .view(-1, ptc.mask_size, ptc.mask_size)
            elif dev_prop:
                h_roi = self._pinned_like('roi', (batch_size, max_pos, 4), torch.float32)
                h_asg = self._pinned_like('asg', (batch_size, max_pos), torch.int32)
                h_np = self._pinned_like('np', (batch_size,), torch.int32)
                h_roi.copy_(sroi[:, :max_pos], non_blocking=True)
                h_asg.copy_(gasg[:, :max_pos], non_blocking=True)
                h_np.copy_(npos, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record()
                self.d2h_bytes += h_roi.numel() * 4 + h_asg.numel() * 4 + h_np.numel() * 4

                def gt_roi_masks():
                    # runs after the head's forward pass has been enqueued
                    ready.synchronize()
                    mt = ptc.mask_targets(h_roi.numpy(), h_asg.numpy(), h_np.numpy(), masks)
                    full = np.full((batch_size, n, ptc.mask_size, ptc.mask_size), -1, np.int32)
                    full[:, :max_pos] = mt
                    return self._upload(full.reshape(-1, ptc.mask_size, ptc.mask_size), np.int32,
                                        dev)
            else:
                rois_h = rois.cpu().numpy()       # host sync: proposals are sampled on the host
                counts = cnt.cpu().numpy()
                self.d2h_bytes += rois_h.nbytes + counts.nbytes
                parts = [self.proposal_target_creator(
                    rois_h[i, :counts[i]], bboxes[i], labels[i], masks[i],
                    self.loc_normalize_mean, self.loc_normalize_std) for i in range(batch_size)]
                sample_rois = self._upload(np.concatenate([p[0] for p in parts]), np.float32, dev)
                sample_idx = self._upload(np.concatenate(
                    [np.full((len(p[0]),), i, np.int32) for i, p in enumerate(parts)]), np.int32, dev)
                gt_roi_locs = self._upload(np.concatenate([p[1] for p in parts]), np.float32, dev)
                gt_roi_labels = self._upload(np.concatenate([p[2] for p in parts]), np.int32, dev)
                gt_roi_masks = self._upload(np.concatenate([p[3] for p in parts]), np.int32, dev)
            self.targets = dict(sample_rois=sample_rois, sample_roi_indices=sample_idx,
                                gt_roi_locs=gt_roi_locs, gt_roi_labels=gt_roi_labels,
                                gt_roi_masks=gt_roi_masks, gt_rpn_locs=gt_rpn_locs,
                                gt_rpn_labels=gt_rpn_labels)
            return self.forward_with_targets(feat, rpn_locs, rpn_scores, rpn_branch=branch[0],
                                             **self.targets)
        finally:
            ctx.recording = False

    def _start_rpn_branch(self, feat, rpn_locs, rpn_scores, anchor, gt, img_size, seed):
        """Anchor targets, the two RPN losses and (eager_rpn_backward) the RPN's backward pass
        on a side stream, forked after the RPN convolutions.  Buffers that the main stream
        reads later are allocated here, on the main stream, before the fork."""
        rpn = self.mask_rcnn.rpn
        dev = feat.device
        n, hh, ww, _ = feat.shape
        A = rpn.n_anchor
        eager = bool(self.eager_rpn_backward)
        b = dict(eager=eager,
                 losses=torch.zeros((8,), dtype=torch.float32, device=dev),
                 g_rpn=torch.empty((n, hh, ww, rpn.g_ld), dtype=torch.float32, device=dev),
                 g_feat=torch.empty(tuple(feat.shape), dtype=torch.float32, device=dev)
                 if eager else None)
        if self._rpn_stream is None:
            self._rpn_stream = torch.cuda.Stream()
        side = self._rpn_stream
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(side), E.ws_slot(1):
            side.wait_event(ready)
            gl, glab = self.anchor_target_creator(gt, anchor, img_size, seed,
                                                  seed_dev=self.seed_dev)
            b['gt_rpn_locs'], b['gt_rpn_labels'] = gl.view(-1, 4), glab.view(-1)
            _lib.call('cmr_rpn_loss', E._p(rpn_locs), 4 * A, E._p(rpn_scores), A,
                      E._p(b['gt_rpn_locs']), E._p(b['gt_rpn_labels']), n * hh * ww, A,
                      float(self.rpn_sigma), E._p(b['g_rpn']), rpn.g_ld, E._p(b['losses']),
                      E.stream())
            b['loss_done'] = torch.cuda.Event()
            b['loss_done'].record(side)
            if eager:
                rpn.backward(b['g_rpn'], out=b['g_feat'], masked=False)
            b['done'] = torch.cuda.Event()
            b['done'].record(side)
        b['keep'] = (feat, rpn_locs, rpn_scores)      # alive until the branch has been joined
        return b

    def _upload(self, a, dtype, dev):
        a = np.ascontiguousarray(a, dtype=dtype)
        self.h2d_bytes += a.nbytes
        return torch.from_numpy(a).to(dev, non_blocking=True)

    def _pinned_like(self, key, shape, dtype):
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=dtype).pin_memory()
            self._pinned[key] = buf
        return buf

    def forward_with_targets(self, feat, rpn_locs, rpn_scores, sample_rois, sample_roi_indices,
                             gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
                             gt_rpn_labels, rpn_branch=None):
        """Head forward + losses for given samples/targets (everything on the device).
        ``gt_roi_masks`` may be a callable producing the tensor; it is called after the
        head's forward pass has been enqueued, so host work inside it overlaps."""
        m, ctx = self.mask_rcnn, self.ctx
        dev = feat.device
        head, rpn = m.head, m.rpn
        cls_locs, scores, masks = head.forward_nhwc(feat, sample_rois, sample_roi_indices)
        if callable(gt_roi_masks):
            gt_roi_masks = gt_roi_masks()
            if self.targets is not None:
                self.targets['gt_roi_masks'] = gt_roi_masks
        R = cls_locs.shape[0]
        n, hh, ww, _ = feat.shape
        A = rpn.n_anchor
        g_lin = torch.empty((R, head.lin_ld), dtype=torch.float32, device=dev)
        g_mask = torch.empty((R, 14, 14, head.mask_ld), dtype=torch.float32, device=dev)
        st = E.stream()
        if rpn_branch is not None:       # RPN losses were enqueued on the side stream
            losses, g_rpn = rpn_branch['losses'], rpn_branch['g_rpn']
        else:
            losses = torch.zeros((8,), dtype=torch.float32, device=dev)
            g_rpn = torch.empty((n, hh, ww, rpn.g_ld), dtype=torch.float32, device=dev)
            _lib.call('cmr_rpn_loss', E._p(rpn_locs), 4 * A, E._p(rpn_scores), A,
                      E._p(gt_rpn_locs), E._p(gt_rpn_labels), n * hh * ww, A,
                      float(self.rpn_sigma), E._p(g_rpn), rpn.g_ld, E._p(losses), st)
        _lib.call('cmr_roi_loss', E._p(cls_locs), 4 * head.n_class, E._p(scores), head.n_class,
                  E._p(gt_roi_locs), E._p(gt_roi_labels), R, head.n_class, float(self.roi_sigma),
                  E._p(g_lin), head.lin_ld, E._p(losses), st)
        _lib.call('cmr_mask_loss', E._p(masks), head.mask_ld, E._p(gt_roi_labels),
                  E._p(gt_roi_masks), R, 14 * 14, head.n_fg, E._p(g_mask), head.mask_ld,
                  E._p(losses), st)
        if rpn_branch is not None:
            torch.cuda.current_stream().wait_event(rpn_branch['loss_done'])
        loss = losses[:5].sum()
        self.observation = {k: losses[i] for i, k in enumerate(LOSS_NAMES)}
        self.observation['loss'] = loss
        self.outputs = dict(rpn_locs=rpn_locs, rpn_scores=rpn_scores, roi_cls_locs=cls_locs,
                            roi_scores=scores, roi_masks=masks)

        def backward(after_head=None, progress=None):
            # weight / bias gradients run on a side stream next to the data-gradient chain
            E.grad_side.begin()
            try:
                if rpn_branch is not None and rpn_branch['eager']:
                    # the RPN branch's gradient of the feature map is already there (or
                    # still being produced on the side stream): the head adds to it
                    def rpn_grad():
                        torch.cuda.current_stream().wait_event(rpn_branch['done'])
                        return rpn_branch['g_feat']
                    g_feat = E.relu_mask(head.backward(g_lin, g_mask, accum=rpn_grad), feat)
                else:
                    if rpn_branch is not None:
                        torch.cuda.current_stream().wait_event(rpn_branch['done'])
                    g_feat = head.backward(g_lin, g_mask)
                    g_feat = rpn.backward(g_rpn, g_feat)
                if after_head is not None:
                    E.grad_side.join()
                    after_head()
                    E.grad_side.begin()
                m.extractor.backward(g_feat, progress)
            finally:
                E.grad_side.join()
            if after_head is None:       # (a hook's owner finishes the two parts itself)
                ctx.finish_grads()

        return Loss(loss, backward)
