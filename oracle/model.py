"""NumPy restatement of the reference model graph (forward and backward), NCHW fp32.

TEST INFRASTRUCTURE (see oracle/__init__.py); PARITY UNPINNED for the Chainer /
ChainerCV pieces (oracle/nn.py, oracle/bbox.py headers); the forward and the hand-written
backward are cross-checked against torch float64 autograd of the same graph
(tests/test_oracle_model.py).  Follows:

  ResNetExtractorBase.__call__   models/resnet_extractor.py:61-90  (conv1 has a bias;
                                 max_pooling_2d(3, stride=2, pad=1) with Chainer's
                                 default cover_all=True; gradients stop after res2)
  BuildingBlock / BottleneckA/B  chainer.links.model.vision.resnet (stride on the first
                                 1x1 conv; every BN replaced by AffineChannel2D,
                                 models/resnet_extractor.py:16-44)
  RegionProposalNetwork.__call__ models/region_proposal_network.py:82-145
  ResNetRoIHead.__call__         models/mask_rcnn_resnet.py:168-196
  MaskRCNN.__call__              models/mask_rcnn.py:142-150
  MaskRCNNTrainChain.__call__    models/mask_rcnn_train_chain.py:76-189 (losses; the
                                 sampled RoIs / targets are inputs here so that both
                                 sides of a parity test see identical samples)

Parameters live in a flat dict keyed like the reference's npz snapshot
('extractor/res4/b3/conv2/W', 'head/deconv6/b', ...; weights OIHW), the naming
implied by examples/coco/convert_caffe2_to_chainer.py:45-249.
"""
import numpy as np

from . import bbox as ob
from . import nn
from . import roi_align as ora

f32 = np.float32

BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


class Config(object):
    def __init__(self, n_layers=50, n_fg_class=80, ratios=(0.5, 1, 2),
                 anchor_scales=(2, 4, 8, 16, 32), roi_size=14, base=64, feat_stride=16,
                 proposal_creator_params=None, rpn_hidden=None):
        self.n_layers = n_layers
        self.n_fg_class = n_fg_class
        self.n_class = n_fg_class + 1
        self.ratios = tuple(ratios)
        self.anchor_scales = tuple(anchor_scales)
        self.n_anchor = len(ratios) * len(anchor_scales)
        self.roi_size = roi_size
        self.base = base                       # 64 in the reference; tests may shrink it
        self.feat_stride = feat_stride
        self.rpn_hidden = rpn_hidden or 16 * base
        self.proposal_creator_params = proposal_creator_params or dict(
            min_size=0, n_test_pre_nms=6000, n_test_post_nms=1000)   # mask_rcnn_resnet.py:48-52

    def stage_dims(self):
        b = self.base
        n2, n3, n4, n5 = BLOCKS[self.n_layers]
        return {'res2': (n2, b, b, 4 * b, 1), 'res3': (n3, 4 * b, 2 * b, 8 * b, 2),
                'res4': (n4, 8 * b, 4 * b, 16 * b, 2),
                'res5': (n5, 16 * b, 8 * b, 32 * b, self.roi_size // 7)}


def block_names(n_layer):
    return ['a'] + ['b%d' % i for i in range(1, n_layer)]


def make_params(cfg, rs):
    """Synthetic parameters: He-normal convs, affine W ~ U[0.5, 1.5]
    scaled so activations stay O(1), b ~ N(0, 0.1); RPN / head as
    mask_rcnn_resnet.py:57-64."""
    p = {}

    def conv(name, o, c, k, std=None, bias=False):
        std = std if std is not None else np.sqrt(2. / (c * k * k))
        p[name + '/W'] = (rs.standard_normal((o, c, k, k)) * std).astype(f32)
        if bias:
            p[name + '/b'] = np.zeros((o,), f32)

    def affine(name, c, gain=1.0):
        p[name + '/W'] = (rs.uniform(0.5, 1.5, c) * gain).astype(f32)
        p[name + '/b'] = (rs.standard_normal(c) * 0.1).astype(f32)

    b = cfg.base
    conv('extractor/conv1', b, 3, 7, std=np.sqrt(2. / (3 * 49)) / 64., bias=True)
    p['extractor/conv1/b'] = (rs.standard_normal(b) * 0.1).astype(f32)
    affine('extractor/bn1', b)
    dims = cfg.stage_dims()
    for stage in ('res2', 'res3', 'res4', 'res5'):
        n_layer, cin, mid, cout, _ = dims[stage]
        root = ('head/' if stage == 'res5' else 'extractor/') + stage
        for blk in block_names(n_layer):
            c_in = cin if blk == 'a' else cout
            conv('%s/%s/conv1' % (root, blk), mid, c_in, 1)
            affine('%s/%s/bn1' % (root, blk), mid)
            conv('%s/%s/conv2' % (root, blk), mid, mid, 3)
            affine('%s/%s/bn2' % (root, blk), mid)
            conv('%s/%s/conv3' % (root, blk), cout, mid, 1)
            affine('%s/%s/bn3' % (root, blk), cout, gain=0.5)
            if blk == 'a':
                conv('%s/%s/conv4' % (root, blk), cout, c_in, 1)
                affine('%s/%s/bn4' % (root, blk), cout, gain=0.5)
    feat = 16 * b
    A = cfg.n_anchor
    conv('rpn/conv1', cfg.rpn_hidden, feat, 3, std=0.01, bias=True)
    conv('rpn/score', A, cfg.rpn_hidden, 1, std=0.01, bias=True)
    conv('rpn/loc', 4 * A, cfg.rpn_hidden, 1, std=0.01, bias=True)
    p['head/cls_loc/W'] = (rs.standard_normal((4 * cfg.n_class, 32 * b)) * 0.001).astype(f32)
    p['head/cls_loc/b'] = np.zeros((4 * cfg.n_class,), f32)
    p['head/score/W'] = (rs.standard_normal((cfg.n_class, 32 * b)) * 0.01).astype(f32)
    p['head/score/b'] = np.zeros((cfg.n_class,), f32)
    p['head/deconv6/W'] = (rs.standard_normal((32 * b, 4 * b, 2, 2)) * 0.01).astype(f32)
    p['head/deconv6/b'] = np.zeros((4 * b,), f32)
    p['head/mask/W'] = (rs.standard_normal((cfg.n_fg_class, 4 * b, 1, 1)) * 0.01).astype(f32)
    p['head/mask/b'] = np.zeros((cfg.n_fg_class,), f32)
    return p


def is_frozen(name):
    """examples/train_common.py:185-190: conv1, bn1, res2 and every AffineChannel2D are
    excluded from updates."""
    parts = name.split('/')
    if parts[0] == 'extractor' and parts[1] in ('conv1', 'bn1', 'res2'):
        return True
    return any(x.startswith('bn') for x in parts)


# ------------------------------------------------------------- forward ------
def _conv_affine(p, root, i, x, stride, pad, act, tape):
    W = p['%s/conv%d/W' % (root, i)]
    h = nn.conv2d(x, W, None, stride, pad)
    y = nn.affine_channel_2d(h, p['%s/bn%d/W' % (root, i)], p['%s/bn%d/b' % (root, i)])
    if tape is not None:
        tape['%s/conv%d' % (root, i)] = h
    return nn.relu(y) if act else y


def bottleneck(p, root, x, stride, is_a, tape=None):
    h1 = _conv_affine(p, root, 1, x, stride, 0, True, tape)
    h2 = _conv_affine(p, root, 2, h1, 1, 1, True, tape)
    h3 = _conv_affine(p, root, 3, h2, 1, 0, False, tape)
    sc = _conv_affine(p, root, 4, x, stride, 0, False, tape) if is_a else x
    y = nn.relu(h3 + sc)
    cache = (x, h1, h2, y)
    if tape is not None:     # every tensor a per-layer parity test needs (tests/test_gpu_config0.py)
        tape[root] = dict(x=x, h1=h1, h2=h2, h3=h3, sc=sc if is_a else None, y=y)
    return y, cache


def building_block(p, root, x, n_layer, stride, tape=None):
    caches = []
    for blk in block_names(n_layer):
        x, c = bottleneck(p, '%s/%s' % (root, blk), x, stride if blk == 'a' else 1,
                          blk == 'a', tape)
        caches.append(c)
    return x, caches


def extractor(cfg, p, x, tape=None):
    """-> (res4 feature map, caches).  Gradients are not propagated below res3."""
    h = nn.conv2d(x, p['extractor/conv1/W'], p['extractor/conv1/b'].reshape(1, -1, 1, 1)
                  .transpose(0, 2, 3, 1).reshape(-1), 2, 3)
    h = nn.relu(nn.affine_channel_2d(h, p['extractor/bn1/W'], p['extractor/bn1/b']))
    if tape is not None:
        tape['extractor/conv1'] = h
    h = nn.max_pooling_2d(h, 3, 2, 1, cover_all=True)
    if tape is not None:
        tape['extractor/pool1'] = h
    dims = cfg.stage_dims()
    caches = {}
    for stage in ('res2', 'res3', 'res4'):
        n_layer, _, _, _, stride = dims[stage]
        h, caches[stage] = building_block(p, 'extractor/' + stage, h, n_layer, stride, tape)
        if tape is not None:
            tape['extractor/' + stage] = h
    return h, caches


def rpn_forward(cfg, p, feat):
    """-> rpn_locs (n, K*A, 4), rpn_scores (n, K*A), anchor (K*A, 4), hidden h."""
    n, _, hh, ww = feat.shape
    base = ob.generate_anchor_base(cfg.feat_stride, cfg.ratios, cfg.anchor_scales)
    anchor = ob.enumerate_shifted_anchor(base, cfg.feat_stride, hh, ww)
    h = nn.relu(nn.conv2d(feat, p['rpn/conv1/W'], None, 1, 1) +
                p['rpn/conv1/b'].reshape(1, -1, 1, 1))
    locs = nn.conv2d(h, p['rpn/loc/W'], None) + p['rpn/loc/b'].reshape(1, -1, 1, 1)
    scores = nn.conv2d(h, p['rpn/score/W'], None) + p['rpn/score/b'].reshape(1, -1, 1, 1)
    rpn_locs = locs.transpose(0, 2, 3, 1).reshape(n, -1, 4)
    rpn_scores = scores.transpose(0, 2, 3, 1).reshape(n, -1)
    return rpn_locs, rpn_scores, anchor, h


def rpn_proposals(cfg, rpn_locs, rpn_scores, anchor, img_size, scales, train):
    pc = ob.ProposalCreator(**cfg.proposal_creator_params)
    rois, idxs, anchor_idx = [], [], []
    for i in range(len(rpn_locs)):
        roi, ai = pc(rpn_locs[i], rpn_scores[i], anchor, img_size, scale=scales[i],
                     train=train, return_index=True)
        rois.append(roi)
        idxs.append(np.full((len(roi),), i, np.int32))
        anchor_idx.append(ai)
    return np.concatenate(rois), np.concatenate(idxs), anchor_idx


def head_forward(cfg, p, feat, rois, roi_indices, tape=None):
    """ResNetRoIHead.__call__ -> roi_cls_locs, roi_scores, roi_masks, cache."""
    idx_rois = np.concatenate((roi_indices.astype(f32)[:, None], rois), axis=1)
    pool = ora.roi_align_2d(feat, idx_rois, cfg.roi_size, cfg.roi_size,
                            1. / cfg.feat_stride, axes='yx')
    n_layer, _, _, _, stride = cfg.stage_dims()['res5']
    res5, caches = building_block(p, 'head/res5', pool, n_layer, stride, tape)
    pool5 = nn.average_pooling_2d(res5, 7, 7)
    cls_locs = nn.linear(pool5, p['head/cls_loc/W'], p['head/cls_loc/b'])
    scores = nn.linear(pool5, p['head/score/W'], p['head/score/b'])
    d6 = nn.relu(nn.deconv2d(res5, p['head/deconv6/W'], p['head/deconv6/b'], 2))
    masks = nn.conv2d(d6, p['head/mask/W'], None) + p['head/mask/b'].reshape(1, -1, 1, 1)
    if tape is not None:
        tape['head/pool'] = pool
        tape['head/res5'] = res5
        tape['head/deconv6'] = d6
    cache = dict(idx_rois=idx_rois, pool=pool, res5=res5, pool5=pool5, d6=d6, blocks=caches,
                 feat_shape=feat.shape)
    return cls_locs, scores, masks, cache


def mask_rcnn_call(cfg, p, x, scales, train=False):
    """MaskRCNN.__call__ (models/mask_rcnn.py:142-150)."""
    feat, _ = extractor(cfg, p, x)
    rpn_locs, rpn_scores, anchor, _ = rpn_forward(cfg, p, feat)
    rois, roi_indices, _ = rpn_proposals(cfg, rpn_locs, rpn_scores, anchor, x.shape[2:],
                                         scales, train)
    cls_locs, scores, masks, _ = head_forward(cfg, p, feat, rois, roi_indices)
    return cls_locs, scores, rois, roi_indices, masks


# ------------------------------------------------------------- losses -------
def train_losses(cfg, rpn_locs, rpn_scores, gt_rpn_locs, gt_rpn_labels, roi_cls_locs,
                 roi_scores, roi_masks, gt_roi_locs, gt_roi_labels, gt_roi_masks,
                 rpn_sigma=3., roi_sigma=1.):
    """The five losses of MaskRCNNTrainChain.__call__ (:160-181) and their gradients
    with respect to the network outputs."""
    n = len(roi_cls_locs)
    rl = rpn_locs.reshape(-1, 4)
    rs_ = rpn_scores.reshape(-1)
    rpn_loc_loss, g_rl = nn.fast_rcnn_loc_loss(rl, gt_rpn_locs, gt_rpn_labels, rpn_sigma)
    rpn_cls_loss, g_rs = nn.sigmoid_cross_entropy(rs_, gt_rpn_labels)
    cl = roi_cls_locs.reshape(n, -1, 4)
    sel = cl[np.arange(n), gt_roi_labels]
    roi_loc_loss, g_sel = nn.fast_rcnn_loc_loss(sel, gt_roi_locs, gt_roi_labels, roi_sigma)
    g_cl = np.zeros_like(cl)
    g_cl[np.arange(n), gt_roi_labels] = g_sel
    roi_cls_loss, g_sc = nn.softmax_cross_entropy(roi_scores, gt_roi_labels)
    msel = roi_masks[np.arange(n), gt_roi_labels - 1]
    roi_mask_loss, g_msel = nn.sigmoid_cross_entropy(msel, gt_roi_masks)
    g_m = np.zeros_like(roi_masks)
    g_m[np.arange(n), gt_roi_labels - 1] = g_msel
    losses = dict(rpn_loc_loss=rpn_loc_loss, rpn_cls_loss=rpn_cls_loss,
                  roi_loc_loss=roi_loc_loss, roi_cls_loss=roi_cls_loss,
                  roi_mask_loss=roi_mask_loss)
    losses['loss'] = f32(sum(losses.values()))
    grads = dict(rpn_locs=g_rl.reshape(rpn_locs.shape), rpn_scores=g_rs.reshape(rpn_scores.shape),
                 roi_cls_locs=g_cl.reshape(n, -1), roi_scores=g_sc, roi_masks=g_m)
    return losses, grads


# ------------------------------------------------------------- backward -----
def _conv_affine_bwd(p, root, i, x, gy_post_affine, stride, pad, grads, need_gx=True):
    """gy is the gradient at the affine output (before any ReLU mask was applied by the
    caller).  AffineChannel2D is frozen: only gx = W * gy is needed."""
    g = p['%s/bn%d/W' % (root, i)].reshape(1, -1, 1, 1) * gy_post_affine
    gx, gW, _ = nn.conv2d_backward(x, p['%s/conv%d/W' % (root, i)], g, stride, pad, need_gx)
    grads['%s/conv%d/W' % (root, i)] = gW
    return gx


def bottleneck_bwd(p, root, cache, gy, stride, is_a, grads, need_gx=True):
    x, h1, h2, y = cache
    g = gy * (y > 0)
    g2 = _conv_affine_bwd(p, root, 3, h2, g, 1, 0, grads) * (h2 > 0)
    g1 = _conv_affine_bwd(p, root, 2, h1, g2, 1, 1, grads) * (h1 > 0)
    gx = _conv_affine_bwd(p, root, 1, x, g1, stride, 0, grads, need_gx)
    if is_a:
        gs = _conv_affine_bwd(p, root, 4, x, g, stride, 0, grads, need_gx)
        return None if not need_gx else gx + gs
    return gx + g


def building_block_bwd(p, root, caches, gy, n_layer, stride, grads, need_gx=True):
    names = block_names(n_layer)
    for k in range(n_layer - 1, -1, -1):
        blk = names[k]
        gy = bottleneck_bwd(p, '%s/%s' % (root, blk), caches[k], gy,
                            stride if blk == 'a' else 1, blk == 'a', grads,
                            need_gx or k > 0)
    return gy


def head_backward(cfg, p, cache, g_cls_locs, g_scores, g_masks, grads):
    """-> gradient with respect to the feature map."""
    res5, pool5, d6 = cache['res5'], cache['pool5'], cache['d6']
    g5a, grads['head/cls_loc/W'], grads['head/cls_loc/b'] = nn.linear_backward(
        pool5, p['head/cls_loc/W'], g_cls_locs)
    g5b, grads['head/score/W'], grads['head/score/b'] = nn.linear_backward(
        pool5, p['head/score/W'], g_scores)
    g_res5 = nn.average_pooling_2d_backward(res5.shape, (g5a + g5b), 7, 7)
    gd6, grads['head/mask/W'], grads['head/mask/b'] = nn.conv2d_backward(
        d6, p['head/mask/W'], g_masks)
    gd6 = gd6 * (d6 > 0)
    gr, grads['head/deconv6/W'], grads['head/deconv6/b'] = nn.deconv2d_backward(
        res5, p['head/deconv6/W'], gd6, 2)
    g_res5 = g_res5 + gr
    n_layer, _, _, _, stride = cfg.stage_dims()['res5']
    g_pool = building_block_bwd(p, 'head/res5', cache['blocks'], g_res5, n_layer, stride, grads)
    idx_rois = cache['idx_rois'][:, [0, 2, 1, 4, 3]]
    return ora.roi_align_backward(cache['feat_shape'], idx_rois, g_pool, cfg.roi_size,
                                  cfg.roi_size, 1. / cfg.feat_stride, 0)


def rpn_backward(cfg, p, feat, h, g_locs, g_scores, grads):
    n, _, hh, ww = feat.shape
    gl = g_locs.reshape(n, hh, ww, -1).transpose(0, 3, 1, 2)
    gs = g_scores.reshape(n, hh, ww, -1).transpose(0, 3, 1, 2)
    gh1, grads['rpn/loc/W'], grads['rpn/loc/b'] = nn.conv2d_backward(h, p['rpn/loc/W'], gl)
    gh2, grads['rpn/score/W'], grads['rpn/score/b'] = nn.conv2d_backward(h, p['rpn/score/W'], gs)
    gh = (gh1 + gh2) * (h > 0)
    gx, grads['rpn/conv1/W'], grads['rpn/conv1/b'] = nn.conv2d_backward(
        feat, p['rpn/conv1/W'], gh, 1, 1)
    return gx


def extractor_backward(cfg, p, caches, g_feat, grads):
    """res4 and res3 only: unchain_backward at res2 (resnet_extractor.py:86-87)."""
    dims = cfg.stage_dims()
    g = building_block_bwd(p, 'extractor/res4', caches['res4'], g_feat, dims['res4'][0],
                           dims['res4'][4], grads)
    building_block_bwd(p, 'extractor/res3', caches['res3'], g, dims['res3'][0],
                       dims['res3'][4], grads, need_gx=False)


def train_step_grads(cfg, p, x, sample_rois, sample_roi_indices, gt_roi_locs, gt_roi_labels,
                     gt_roi_masks, gt_rpn_locs, gt_rpn_labels):
    """Forward + backward of MaskRCNNTrainChain with the sampled targets given.
    -> (losses dict, grads dict over the trainable parameters)."""
    feat, caches = extractor(cfg, p, x)
    rpn_locs, rpn_scores, anchor, h = rpn_forward(cfg, p, feat)
    cls_locs, scores, masks, hc = head_forward(cfg, p, feat, sample_rois, sample_roi_indices)
    losses, g = train_losses(cfg, rpn_locs, rpn_scores, gt_rpn_locs, gt_rpn_labels, cls_locs,
                             scores, masks, gt_roi_locs, gt_roi_labels, gt_roi_masks)
    grads = {}
    g_feat = head_backward(cfg, p, hc, g['roi_cls_locs'], g['roi_scores'], g['roi_masks'], grads)
    g_feat = g_feat + rpn_backward(cfg, p, feat, h, g['rpn_locs'], g['rpn_scores'], grads)
    extractor_backward(cfg, p, caches, g_feat, grads)
    return losses, grads
