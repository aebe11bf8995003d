"""Region proposal network on the tensor-core / NMS kernels.

Mirrors ``RegionProposalNetwork`` (chainer_mask_rcnn/models/
region_proposal_network.py:26-145): h = relu(conv1 3x3), loc = 1x1 -> 4A,
score = 1x1 -> A (one logit per anchor), outputs reshaped to (N, H*W*A, 4) and
(N, H*W*A) -- which is exactly the channels-last memory the kernels write -- then
one ``ProposalCreator`` pass.  The reference loops over images and round-trips every
image through the host (:135-141); here the whole batch is one device call and the
only host sync is reading the per-image proposal counts.
"""
import numpy as np
import torch

from . import engine as E
from .layers import Conv
from ..utils import ProposalCreator
from ..utils import generate_anchor_base


def _enumerate_shifted_anchor(anchor_base, feat_stride, height, width):
    """(K*A, 4) float32 anchors, K row-major over (y, x), A inner
    (region_proposal_network.py:148-167)."""
    sy = np.arange(height, dtype=np.float32) * feat_stride
    sx = np.arange(width, dtype=np.float32) * feat_stride
    shift = np.stack(np.broadcast_arrays(sy[:, None], sx[None, :], sy[:, None], sx[None, :]),
                     axis=-1).reshape(-1, 1, 4)
    return (shift + anchor_base[None].astype(np.float32)).reshape(-1, 4).astype(np.float32)


class RegionProposalNetwork(object):

    def __init__(self, ctx, in_channels=512, mid_channels=512, ratios=(0.5, 1, 2),
                 anchor_scales=(8, 16, 32), feat_stride=16, proposal_creator_params=None,
                 root='rpn'):
        self.ctx = ctx
        self.anchor_base = generate_anchor_base(anchor_scales=anchor_scales, ratios=ratios)
        self.feat_stride = feat_stride
        self.proposal_layer = ProposalCreator(**(proposal_creator_params or {}))
        self.n_anchor = A = self.anchor_base.shape[0]
        self.mid = mid_channels
        self.conv1 = Conv(ctx, root + '/conv1', in_channels, mid_channels, 3, 1, 1, bias=True)
        self.score = Conv(ctx, root + '/score', mid_channels, A, 1, bias=True, need_dgrad=False)
        self.loc = Conv(ctx, root + '/loc', mid_channels, 4 * A, 1, bias=True, need_dgrad=False)
        # fused data-gradient bank of loc and score: (mid, [4A | A | 0-pad to 32k])
        self.g_ld = (5 * A + 31) // 32 * 32
        self.w_dgrad = None
        self._anchor_cache = {}
        self.saved = None
        ctx.layers.append(self)

    def prep_frozen(self):
        pass

    def prep_backward(self):
        c, A = self.ctx, self.n_anchor
        if self.w_dgrad is None:
            self.w_dgrad = torch.zeros((self.mid, self.g_ld), dtype=torch.float32, device=c.device)
        E.prep_dgrad_weight(c.param(self.loc.W), 4 * A, 1, self.mid, self.mid, self.mid, None,
                            False, self.w_dgrad, self.g_ld, 0)
        E.prep_dgrad_weight(c.param(self.score.W), A, 1, self.mid, self.mid, self.mid, None,
                            False, self.w_dgrad, self.g_ld, 4 * A)

    def anchors(self, hh, ww, device):
        key = (hh, ww, str(device))
        if key not in self._anchor_cache:
            a = _enumerate_shifted_anchor(self.anchor_base, self.feat_stride, hh, ww)
            self._anchor_cache[key] = (a, torch.from_numpy(a).to(device))
        return self._anchor_cache[key]

    def forward_nhwc(self, feat, img_size, scales, between=None):
        """feat (N,H,W,C) -> rpn_locs (N,HWA,4), rpn_scores (N,HWA), rois (N,n_post,4),
        anchor index (N,n_post), counts (N,), anchor (host, device).
        ``between(rpn_locs, rpn_scores, anchor)``, when given, is called after the three
        convolutions and before the proposal layer is enqueued (the train chain forks the
        RPN loss / backward branch there, next to the latency-bound proposal chain)."""
        n, hh, ww, _ = feat.shape
        anchor_np, anchor = self.anchors(hh, ww, feat.device)
        h = self.conv1.forward(feat, relu=True)
        locs = self.loc.forward(h, round_out=False)
        scores = self.score.forward(h, round_out=False)
        if self.ctx.recording:
            self.saved = (feat, h)
        rpn_locs = locs.view(n, -1, 4)
        rpn_scores = scores.view(n, -1)
        if between is not None:
            between(rpn_locs, rpn_scores, anchor)
        scale = float(np.asarray(scales).ravel()[0]) if np.size(scales) else 1.
        if np.size(scales) > 1 and self.proposal_layer.min_size != 0 and \
                not np.all(np.asarray(scales) == scale):
            # per-image min_size: one call per image, still on the device
            outs = [self.proposal_layer.batch(rpn_locs[i:i + 1], rpn_scores[i:i + 1], anchor,
                                              img_size, float(scales[i])) for i in range(n)]
            rois, idx, cnt = (torch.cat([o[k] for o in outs]) for k in range(3))
        else:
            rois, idx, cnt = self.proposal_layer.batch(rpn_locs, rpn_scores, anchor, img_size,
                                                       scale)
        return rpn_locs, rpn_scores, rois, idx, cnt, (anchor_np, anchor)

    def __call__(self, x, img_size, scales):
        """x: (N, C, H, W) feature map.  Returns (rpn_locs, rpn_scores, rois,
        roi_indices, anchor) like the reference."""
        self.ctx.prepare(backward=False)
        feat = E.to_nhwc(x)
        rpn_locs, rpn_scores, rois, _, cnt, (_, anchor) = self.forward_nhwc(feat, img_size, scales)
        rois, roi_indices = flatten_proposals(rois, cnt)
        return rpn_locs, rpn_scores, rois, roi_indices, anchor

    def backward(self, g, g_feat_other=None, out=None, masked=True):
        """g: (N,H,W,g_ld) gradient of the losses w.r.t. [loc | score] (cmr_rpn_loss);
        g_feat_other: gradient already flowing into the feature map (from the RoI head),
        added in the epilogue.  Returns dL/dfeat * ReLU mask of feat; with ``masked=False``
        the RPN branch's own contribution, not masked and not rounded, written to ``out``
        (the branch then runs ahead of the head's backward pass, which adds to it)."""
        feat, h = self.saved
        self.saved = None
        c, A = self.ctx, self.n_anchor
        hw = h.shape[1:3]
        E.wgrad_tap(g, h, c.grad(self.loc.W), 4 * A, self.mid, hw, self.mid)
        E.wgrad_tap(g, h, c.grad(self.score.W), A, self.mid, hw, self.mid, gy_c0=4 * A)
        E.column_sums(g, 0, 4 * A, c.grad(self.loc.b))
        E.column_sums(g, 4 * A, A, c.grad(self.score.b))
        gh = E.conv_gemm(g, self.w_dgrad, self.mid, mask=h)
        self.conv1.backward_w(gh, feat)
        if not masked:
            return self.conv1.backward_x(gh, feat.shape[1:3], out=out, round_out=False)
        return self.conv1.backward_x(gh, feat.shape[1:3], addend=g_feat_other, mask=feat,
                                     out=out)


def flatten_proposals(rois, cnt):
    """(N, n_post, 4) padded proposals + counts -> concatenated (R', 4), (R',) int32."""
    counts = cnt.tolist()                      # the one host sync of the proposal stage
    parts, idx = [], []
    for i, k in enumerate(counts):
        parts.append(rois[i, :k])
        idx.append(torch.full((k,), i, dtype=torch.int32, device=rois.device))
    return torch.cat(parts, dim=0), torch.cat(idx, dim=0)
