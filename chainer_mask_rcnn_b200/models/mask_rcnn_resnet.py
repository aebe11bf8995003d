"""Mask R-CNN with a ResNet-C4 backbone and a res5 RoI head.

Mirrors ``MaskRCNNResNet`` and ``ResNetRoIHead`` (chainer_mask_rcnn/models/
mask_rcnn_resnet.py:30-196).  Parameters carry the reference's names
('extractor/res4/b3/conv2/W', 'head/deconv6/W', ...) and, through
``namedparams`` / ``state_dict``, the reference's layouts.
"""
import numpy as np
import torch

from . import engine as E
from .. import functions
from .layers import BuildingBlock, Context, Conv
from .mask_rcnn import MaskRCNN
from .region_proposal_network import RegionProposalNetwork
from .resnet_extractor import ResNetExtractorBase


class _Deconv2x2(object):
    """Deconvolution2D(cin, cout, 2, stride=2) + bias + ReLU: ONE GEMM with N = 4 * cout
    (the four filter taps side by side; the input is read once) whose column blocks are
    interleaved by a pixel-shuffle store (models/mask_rcnn_resnet.py:138-139)."""

    def __init__(self, ctx, name, cin, cout):
        self.ctx, self.cin, self.cout = ctx, cin, cout
        self.W = ctx.add_param(name + '/W', (4, cout, cin), 'deconv', (cin, cout, 2, 2), True)
        self.b = ctx.add_param(name + '/b', (cout,), 'vec', (cout,), True)
        self.w_dgrad = None
        ctx.layers.append(self)

    def prep_frozen(self):
        pass

    def prep_backward(self):
        c = self.ctx
        if self.w_dgrad is None:
            self.w_dgrad = torch.empty((self.cin, 2, 2, self.cout), dtype=torch.float32,
                                       device=c.device)
        # gx[y,x,c] = sum_{dy,dx,o} gy[2y+dy, 2x+dx, o] W[c,o,dy,dx]: a 2x2 stride-2
        # correlation, taps not flipped
        E.prep_dgrad_weight(c.param(self.W), self.cout, 4, self.cin, self.cin,
                            self.cout * self.cin, None, False, self.w_dgrad)

    def forward(self, x):
        R, h, w, _ = x.shape
        out = torch.empty((R, 2 * h, 2 * w, self.cout), dtype=torch.float32, device=x.device)
        if self.ctx.precision == 'tf32x3':
            c = self.ctx
            if getattr(self, '_w3_cache', (None, None))[0] != c.version:
                self._w3_cache = (c.version, E.split3(c.param(self.W), order=1))
            E.conv_gemm(E.split3(x), self._w3_cache[1], 4 * self.cout, out=out,
                        bias=c.param(self.b), relu=True, round_out=False, d_stride=2,
                        tap_cols=self.cout)
            return out
        wt = self.ctx.fwd(self.W)                       # (4, cout, cin): tap-major rows
        E.conv_gemm(x, wt, 4 * self.cout, out=out, bias=self.ctx.param(self.b), relu=True,
                    d_stride=2, tap_cols=self.cout)
        return out

    def backward(self, g, x, bcast=None, mask=None):
        """g = dL/dy * [y > 0] (R,2h,2w,cout); returns dL/dx -- not masked, not rounded --
        or, with ``bcast`` (R, cin) and ``mask``: relu_mask(dL/dx + bcast / (h*w)), rounded
        (the average-pooling branch's gradient folded into the same epilogue)."""
        c = self.ctx
        gw = c.grad(self.W)
        hw = x.shape[1:3]
        for t in range(4):
            E.wgrad_tap(g, x, gw[t], self.cout, self.cin, hw, self.cin, gy_stride=2,
                        gy_off=(t // 2, t % 2))
        E.column_sums(g, 0, self.cout, c.grad(self.b))
        if bcast is None:
            return E.conv_gemm(g, self.w_dgrad, self.cin, 2, 2, 2, 0, round_out=False)
        return E.conv_gemm(g, self.w_dgrad, self.cin, 2, 2, 2, 0, mask=mask, bcast=bcast,
                           bcast_group=hw[0] * hw[1], bcast_scale=1.0 / (hw[0] * hw[1]))


class ResNetRoIHead(object):

    mask_size = 14  # Size of the predicted mask.

    def __init__(self, ctx, n_layers, n_class, roi_size, spatial_scale,
                 pooling_func=functions.roi_align_2d, base=64, root='head'):
        b = base
        self.ctx = ctx
        self.n_class = n_class
        self.roi_size = roi_size
        self.spatial_scale = spatial_scale
        self.pooling_func = pooling_func
        self.res5 = BuildingBlock(ctx, root + '/res5', 3, 16 * b, 8 * b, 32 * b, 1)
        # the reference gives res5.a stride roi_size // 7 (2 for a 14x14 pool); the pool
        # is produced at that stride instead (only those bins are ever read)
        self.bin_stride = max(roi_size // 7, 1)
        self.feat = 32 * b
        n_fg = n_class - 1
        self.cls_loc = Conv(ctx, root + '/cls_loc', self.feat, 4 * n_class, 1, bias=True,
                            need_dgrad=False)
        self.score = Conv(ctx, root + '/score', self.feat, n_class, 1, bias=True,
                          need_dgrad=False)
        for l, nm in ((self.cls_loc, 'cls_loc'), (self.score, 'score')):   # Linear: (O, I)
            ctx.kinds[root + '/%s/W' % nm] = ('linear', (l.cout, l.cin), True)
        self.deconv6 = _Deconv2x2(ctx, root + '/deconv6', self.feat, 4 * b)
        self.mask = Conv(ctx, root + '/mask', 4 * b, n_fg, 1, bias=True, need_dgrad=False)
        self.n_fg = n_fg
        self.lin_ld = (5 * n_class + 31) // 32 * 32      # fused [cls_loc | score] gradient
        self.mask_ld = (n_fg + 31) // 32 * 32
        self.w_dgrad_lin = None
        self.w_dgrad_mask = None
        self.saved = None
        ctx.layers.append(self)

    def prep_frozen(self):
        pass

    def prep_backward(self):
        c, nc = self.ctx, self.n_class
        dev = c.device
        if self.w_dgrad_lin is None:
            self.w_dgrad_lin = torch.zeros((self.feat, self.lin_ld), dtype=torch.float32, device=dev)
            self.w_dgrad_mask = torch.zeros((self.mask.cin, self.mask_ld), dtype=torch.float32,
                                            device=dev)
        F = self.feat
        E.prep_dgrad_weight(c.param(self.cls_loc.W), 4 * nc, 1, F, F, F, None, False,
                            self.w_dgrad_lin, self.lin_ld, 0)
        E.prep_dgrad_weight(c.param(self.score.W), nc, 1, F, F, F, None, False,
                            self.w_dgrad_lin, self.lin_ld, 4 * nc)
        M = self.mask.cin
        E.prep_dgrad_weight(c.param(self.mask.W), self.n_fg, 1, M, M, M, None, False,
                            self.w_dgrad_mask, self.mask_ld, 0)

    def _rois_xy(self, rois, roi_indices):
        """[idx, y1, x1, y2, x2] -> the kernel's (idx, x1, y1, x2, y2) rows."""
        r = torch.empty((rois.shape[0], 5), dtype=torch.float32, device=rois.device)
        r[:, 0] = roi_indices.to(torch.float32)
        r[:, 1] = rois[:, 1]
        r[:, 2] = rois[:, 0]
        r[:, 3] = rois[:, 3]
        r[:, 4] = rois[:, 2]
        return r

    def forward_nhwc(self, feat, rois, roi_indices, pred_bbox=True, pred_mask=True):
        """feat (N,H,W,C) NHWC; rois (R,4) yx; -> (R,4*n_class), (R,n_class),
        (R,14,14,n_fg) [a view of a mask_ld-wide buffer]."""
        rois_xy = self._rois_xy(rois, roi_indices)
        rnd = self.ctx.precision == 'tf32'      # operands are TF32-rounded by their producer
        if self.pooling_func is functions.roi_align_2d:
            pool = E.roi_align_nhwc(feat, rois_xy, self.roi_size, self.roi_size, self.bin_stride,
                                    self.spatial_scale, round_out=rnd)
        else:   # a user-supplied pooler works on the reference's NCHW arrays
            p = self.pooling_func(E.as_nchw_view(feat),
                                  torch.cat((roi_indices.to(torch.float32)[:, None], rois), 1),
                                  outh=self.roi_size, outw=self.roi_size,
                                  spatial_scale=self.spatial_scale, axes='yx')
            p = p[:, :, ::self.bin_stride, ::self.bin_stride]
            pool = p.permute(0, 2, 3, 1).contiguous()
            if rnd:
                pool = E.round_tf32(pool)
        res5 = self.res5.forward(pool)
        cls_locs = scores = masks = pool5 = d6 = None
        if pred_bbox:
            pool5 = E.avg_pool(res5, round_out=rnd)
            p4 = pool5.view(-1, 1, 1, self.feat)
            cls_locs = self.cls_loc.forward(p4, round_out=False).view(-1, 4 * self.n_class)
            scores = self.score.forward(p4, round_out=False).view(-1, self.n_class)
        if pred_mask:
            d6 = self.deconv6.forward(res5)
            R = d6.shape[0]
            mbuf = torch.zeros((R, 14, 14, self.mask_ld), dtype=torch.float32, device=feat.device)
            self.mask.forward(d6, round_out=False, out=mbuf)
            masks = mbuf[..., :self.n_fg]
        if self.ctx.recording:
            self.saved = dict(rois_xy=rois_xy, feat_shape=tuple(feat.shape), pool=pool, res5=res5,
                              pool5=pool5, d6=d6)
        return cls_locs, scores, masks

    def __call__(self, x, rois, roi_indices, pred_bbox=True, pred_mask=True):
        """x (N,C,H,W) feature map; rois (R,4) (y1,x1,y2,x2); roi_indices (R,)."""
        from .mask_rcnn import as_device_f32
        self.ctx.prepare(backward=False)
        rois = as_device_f32(rois)
        if isinstance(roi_indices, np.ndarray):
            roi_indices = torch.from_numpy(roi_indices)
        roi_indices = roi_indices.to(rois.device)
        cls_locs, scores, masks = self.forward_nhwc(E.to_nhwc(x), rois, roi_indices, pred_bbox,
                                                    pred_mask)
        if masks is not None:
            masks = E.as_nchw_view(masks)
        return cls_locs, scores, masks

    def backward(self, g_lin, g_mask, accum=None):
        """g_lin (R, lin_ld): [d cls_loc | d score | 0]; g_mask (R,14,14,mask_ld).
        Returns dL/dfeat (N,H,W,C), not masked; ``accum``: a zero-argument callable returning
        the (N,H,W,C) tensor the gradient is added to instead (called right before the
        ROIAlign backward, so that whatever produces it may still be running until then)."""
        s, c, nc = self.saved, self.ctx, self.n_class
        self.saved = None
        res5, d6, pool5 = s['res5'], s['d6'], s['pool5']
        R = res5.shape[0]
        # mask branch
        E.wgrad_tap(g_mask, d6, c.grad(self.mask.W), self.n_fg, self.mask.cin, (14, 14),
                    self.mask.cin)
        E.column_sums(g_mask, 0, self.n_fg, c.grad(self.mask.b))
        gd6 = E.conv_gemm(g_mask, self.w_dgrad_mask, self.mask.cin, mask=d6)
        # box branch
        p4 = pool5.view(R, 1, 1, self.feat)
        g4 = g_lin.view(R, 1, 1, self.lin_ld)
        E.wgrad_tap(g4, p4, c.grad(self.cls_loc.W), 4 * nc, self.feat, (1, 1), self.feat)
        E.wgrad_tap(g4, p4, c.grad(self.score.W), nc, self.feat, (1, 1), self.feat, gy_c0=4 * nc)
        E.column_sums(g_lin, 0, 4 * nc, c.grad(self.cls_loc.b))
        E.column_sums(g_lin, 4 * nc, nc, c.grad(self.score.b))
        g_pool5 = E.conv_gemm(g4, self.w_dgrad_lin, self.feat, round_out=False).view(R, self.feat)
        # both branches meet at res5's output: the mask branch's data gradient is produced
        # with the pooled box-branch gradient (g_pool5 / 49 per pixel), res5's ReLU mask and
        # the tf32 rounding already applied in its epilogue
        g_res5 = self.deconv6.backward(gd6, res5, bcast=g_pool5, mask=res5)
        g_pool = self.res5.backward(g_res5, input_is_relu=False)
        return E.roi_align_nhwc_bwd(g_pool, s['rois_xy'], s['feat_shape'], self.roi_size,
                                    self.roi_size, self.bin_stride, self.spatial_scale,
                                    accum=accum() if accum is not None else None,
                                    deterministic=self.ctx.deterministic)


class MaskRCNNResNet(MaskRCNN):

    feat_stride = 16

    def __init__(self, n_layers, n_fg_class, pretrained_model=None, min_size=600, max_size=1000,
                 ratios=(0.5, 1, 2), anchor_scales=(4, 8, 16, 32),
                 mean=(123.152, 115.903, 103.063), res_initialW=None, rpn_initialW=None,
                 loc_initialW=None, score_initialW=None, mask_initialW=None,
                 proposal_creator_params=dict(min_size=0, n_test_pre_nms=6000,
                                              n_test_post_nms=1000),
                 pooling_func=functions.roi_align_2d, rpn_hidden=1024, roi_size=7,
                 base_channels=64, device=None, seed=0, precision='tf32'):
        if n_layers not in (50, 101):
            raise ValueError('n_layers must be 50 or 101')
        if len(mean) != 3:
            raise ValueError('The mean must be tuple of RGB values.')
        if not torch.cuda.is_available():
            raise RuntimeError('chainer_mask_rcnn_b200 needs a CUDA device (B200); there is no '
                               'CPU fallback')
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.ctx = ctx = Context()
        self._n_layers = n_layers
        b = base_channels
        extractor = ResNetExtractorBase(ctx, n_layers, base=b)
        rpn = RegionProposalNetwork(ctx, 16 * b, rpn_hidden * b // 64, ratios=ratios,
                                    anchor_scales=anchor_scales, feat_stride=self.feat_stride,
                                    proposal_creator_params=proposal_creator_params)
        head = ResNetRoIHead(ctx, n_layers, n_fg_class + 1, roi_size, 1. / self.feat_stride,
                             pooling_func=pooling_func, base=b)
        ctx.finalize(device)
        self.precision = precision
        super(MaskRCNNResNet, self).__init__(
            extractor, rpn, head, mean=np.asarray(mean, dtype=np.float32)[:, None, None],
            min_size=min_size, max_size=max_size)
        self._init_params(seed, res_initialW, rpn_initialW, loc_initialW, score_initialW,
                          mask_initialW)
        if pretrained_model:
            self.load_npz(pretrained_model)

    @property
    def precision(self):
        """'tf32' (default: TF32 tensor-core products of operands rounded by their producer --
        the speed path, <= 1e-3 per operator) or 'tf32x3' (forward only: every GEMM on
        hi/lo-split operands, three times the work, fp32-level results through the whole
        chained model -- the parity mode)."""
        return self.ctx.precision

    @precision.setter
    def precision(self, value):
        if value not in ('tf32', 'tf32x3'):
            raise ValueError("precision must be 'tf32' or 'tf32x3'")
        self.ctx.precision = value

    def load_imagenet_resnet(self, src, bgr_to_rgb=True):
        """What the reference does when no ``pretrained_model`` is given
        (``pretrained_model='auto'``): take Chainer's ImageNet ``ResNet{50,101}Layers``
        snapshot (keys ``conv1/W``, ``bn1/{gamma,beta,avg_mean,avg_var}``,
        ``res2/a/conv1/W`` ... ``res5/b2/bn3/avg_var``; ``fc6`` is ignored), flip conv1 from
        BGR to RGB input (resnet_extractor.py:53-56), fold every BatchNormalization into an
        AffineChannel2D (:16-44, :59) and copy conv1 .. res4 into the extractor and res5 into
        the head (mask_rcnn_resnet.py:68-76, 158-166).  ``src``: path of the ``.npz`` or a
        name -> array mapping."""
        from .resnet_extractor import _convert_bn_to_affine
        if isinstance(src, str):
            with np.load(src) as z:
                src = {k: z[k] for k in z.files}
        src = {k.lstrip('/'): v for k, v in src.items()}
        if bgr_to_rgb and 'conv1/W' in src:
            src['conv1/W'] = np.ascontiguousarray(np.asarray(src['conv1/W'])[:, ::-1])
        params = {}
        for key, value in _convert_bn_to_affine(src).items():
            stage = key.split('/')[0]
            if stage in ('conv1', 'bn1', 'res2', 'res3', 'res4'):
                params['extractor/' + key] = value
            elif stage == 'res5':
                params['head/' + key] = value
        missing = [n for n in self.ctx.names()
                   if ((n.startswith('extractor/') and
                        n.split('/')[1] in ('conv1', 'bn1', 'res2', 'res3', 'res4')) or
                       n.startswith('head/res5/')) and n not in params]
        if missing:
            raise KeyError('ResNet snapshot lacks: ' + ', '.join(sorted(missing)[:5]))
        self.load_state_dict(params, strict=False)

    def _init_params(self, seed, res_std, rpn_std, loc_std, score_std, mask_std):
        """Reference initialisers (mask_rcnn_resnet.py:57-64): Normal(0.01) for the RPN,
        score and mask layers, Normal(0.001) for cls_loc, zero biases.  The ImageNet
        weights the reference downloads are replaced by He-normal residual convolutions
        with identity-like affines (no network here); load real ones with load_npz."""
        g = torch.Generator(device='cpu').manual_seed(seed)
        ctx = self.ctx

        def normal(name, std):
            v = ctx.param(name)
            v.copy_((torch.randn(v.shape, generator=g) * std).to(v.device))

        for name in ctx.names():
            kind = ctx.kinds[name][0]
            v = ctx.param(name)
            leaf = name.split('/')
            if leaf[-1] == 'b' and not leaf[-2].startswith('bn'):
                v.zero_()
            elif leaf[-2].startswith('bn'):
                # a slope below one on the two affines feeding each residual sum keeps the
                # activations O(1) through a randomly initialised net: the variance grows by
                # (1 + slope^2) per block -- 0.5 for the 16 blocks of R50, 0.3 for the 33 of
                # R101 (with 0.5 the R101 losses start at ~30 and the reference's learning
                # rate for 16 images, 0.02, overflows within a few hundred steps)
                slope = 0.5 if self._n_layers == 50 else 0.3
                gain = slope if leaf[-2] in ('bn3', 'bn4') else 1.0
                v.fill_(gain if leaf[-1] == 'W' else 0.0)
            elif name == 'extractor/conv1/W':
                # inputs are mean-subtracted 8-bit pixels (|x| ~ 128)
                normal(name, float(np.sqrt(2. / 147.)) / 64. if res_std is None else res_std)
            elif name.startswith('rpn/'):
                normal(name, 0.01 if rpn_std is None else rpn_std)
            elif name == 'head/cls_loc/W':
                normal(name, 0.001 if loc_std is None else loc_std)
            elif name == 'head/score/W':
                normal(name, 0.01 if score_std is None else score_std)
            elif name in ('head/deconv6/W', 'head/mask/W'):
                normal(name, 0.01 if mask_std is None else mask_std)
            elif kind == 'conv':
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
                normal(name, float(np.sqrt(2. / fan_in)) if res_std is None else res_std)
        ctx.mark_dirty()
