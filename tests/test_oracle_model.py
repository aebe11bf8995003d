"""oracle/model.py -- the hand-written forward AND backward of the whole train graph that
the GPU parity tests compare against -- checked against an independent construction: the
same graph written with torch CPU ops in float64 (torchvision's roi_align(aligned=False)
for the pooler) and differentiated by autograd.  Losses and every trainable gradient must
agree; a mistake in the oracle's backward chain (ReLU gates, residual joins, the
unchain_backward at res2, the frozen affines, loss normalisers) shows up here."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
import torchvision

import synth
from oracle import model as om

D = torch.float64


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(D)


class TorchGraph(object):
    def __init__(self, cfg, params):
        self.cfg = cfg
        self.p = {k: _t(v).requires_grad_(not om.is_frozen(k)) for k, v in params.items()}

    def conv_affine(self, root, i, x, stride, pad, act):
        p = self.p
        h = F.conv2d(x, p['%s/conv%d/W' % (root, i)], None, stride, pad)
        y = p['%s/bn%d/W' % (root, i)].view(1, -1, 1, 1) * h + p['%s/bn%d/b' % (root, i)].view(1, -1, 1, 1)
        return F.relu(y) if act else y

    def bottleneck(self, root, x, stride, is_a):
        h = self.conv_affine(root, 1, x, stride, 0, True)
        h = self.conv_affine(root, 2, h, 1, 1, True)
        h = self.conv_affine(root, 3, h, 1, 0, False)
        sc = self.conv_affine(root, 4, x, stride, 0, False) if is_a else x
        return F.relu(h + sc)

    def stage(self, root, x, n_layer, stride):
        for blk in om.block_names(n_layer):
            x = self.bottleneck('%s/%s' % (root, blk), x, stride if blk == 'a' else 1, blk == 'a')
        return x

    def extractor(self, x):
        p, dims = self.p, self.cfg.stage_dims()
        h = F.conv2d(x, p['extractor/conv1/W'], p['extractor/conv1/b'], 2, 3)
        h = F.relu(p['extractor/bn1/W'].view(1, -1, 1, 1) * h + p['extractor/bn1/b'].view(1, -1, 1, 1))
        h = F.max_pool2d(h, 3, 2, 1, ceil_mode=True)
        h = self.stage('extractor/res2', h, dims['res2'][0], dims['res2'][4]).detach()   # unchain_backward
        h = self.stage('extractor/res3', h, dims['res3'][0], dims['res3'][4])
        return self.stage('extractor/res4', h, dims['res4'][0], dims['res4'][4])

    def losses(self, x, rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
               gt_rpn_labels):
        cfg, p = self.cfg, self.p
        feat = self.extractor(x)
        n = feat.shape[0]
        h = F.relu(F.conv2d(feat, p['rpn/conv1/W'], p['rpn/conv1/b'], 1, 1))
        locs = F.conv2d(h, p['rpn/loc/W'], p['rpn/loc/b'])
        scores = F.conv2d(h, p['rpn/score/W'], p['rpn/score/b'])
        rpn_locs = locs.permute(0, 2, 3, 1).reshape(-1, 4)
        rpn_scores = scores.permute(0, 2, 3, 1).reshape(-1)
        # head; rois are (y1, x1, y2, x2), torchvision wants (idx, x1, y1, x2, y2)
        tv_rois = torch.cat([_t(idx.astype(np.float64))[:, None], _t(rois[:, [1, 0, 3, 2]])], 1)
        pool = torchvision.ops.roi_align(feat, tv_rois, (cfg.roi_size, cfg.roi_size),
                                         1. / cfg.feat_stride, 0, False)
        d5 = cfg.stage_dims()['res5']
        res5 = self.stage('head/res5', pool, d5[0], d5[4])
        pool5 = F.avg_pool2d(res5, 7, 7).flatten(1)
        cls_locs = F.linear(pool5, p['head/cls_loc/W'], p['head/cls_loc/b'])
        roi_scores = F.linear(pool5, p['head/score/W'], p['head/score/b'])
        d6 = F.relu(F.conv_transpose2d(res5, p['head/deconv6/W'], p['head/deconv6/b'], stride=2))
        masks = F.conv2d(d6, p['head/mask/W'], p['head/mask/b'])

        def loc_loss(pred, gt, label, sigma):
            pos = torch.from_numpy(label > 0)
            s = F.smooth_l1_loss(pred[pos], _t(gt)[pos], reduction='sum', beta=1. / sigma ** 2)
            return s / int((label >= 0).sum())

        def bce_ignore(logit, target):
            valid = torch.from_numpy(target != -1)
            s = F.binary_cross_entropy_with_logits(logit[valid], _t(target.astype(np.float64))[valid],
                                                   reduction='sum')
            return s / max(int(valid.sum()), 1)

        R = len(rois)
        ar = torch.arange(R)
        lab = torch.from_numpy(gt_roi_labels.astype(np.int64))
        out = dict(
            rpn_loc_loss=loc_loss(rpn_locs, gt_rpn_locs, gt_rpn_labels, 3.),
            rpn_cls_loss=bce_ignore(rpn_scores, gt_rpn_labels),
            roi_loc_loss=loc_loss(cls_locs.view(R, -1, 4)[ar, lab], gt_roi_locs, gt_roi_labels, 1.),
            roi_cls_loss=F.cross_entropy(roi_scores, lab),
            # background rows index class -1 = the last mask channel; their target is all -1
            roi_mask_loss=bce_ignore(masks[ar, lab - 1], gt_roi_masks),
        )
        out['loss'] = sum(out.values())
        return out


@pytest.mark.parametrize('n_layers_seed', [0, 1])
def test_losses_and_gradients_match_autograd(n_layers_seed):
    rs = np.random.RandomState(50 + n_layers_seed)
    cfg = om.Config(n_layers=50, n_fg_class=3, anchor_scales=(4, 8), roi_size=14, base=4)
    params = om.make_params(cfg, rs)
    # the synthetic conv1 gain assumes [0, 255] pixels; give every layer a visible signal
    x = (rs.uniform(0, 255, (2, 3, 96, 128)) - 120.).astype(np.float32)
    feat, _ = om.extractor(cfg, params, x)
    assert feat.shape[2:] == (7, 9) and np.abs(feat).max() > 1e-3
    n_anchor = feat.shape[2] * feat.shape[3] * cfg.n_anchor * 2
    n_roi = 10
    rois = synth.random_boxes(rs, n_roi, 96, 128, 12., 90.)
    idx = (np.arange(n_roi) % 2).astype(np.int32)
    gt_roi_locs = (rs.standard_normal((n_roi, 4)) * 0.5).astype(np.float32)
    gt_roi_labels = rs.randint(0, cfg.n_class, n_roi).astype(np.int32)
    gt_roi_labels[:2] = 0
    gt_roi_masks = rs.randint(0, 2, (n_roi, 14, 14)).astype(np.int32)
    gt_roi_masks[gt_roi_labels == 0] = -1
    gt_rpn_labels = rs.choice([-1, 0, 1], size=n_anchor, p=[0.6, 0.25, 0.15]).astype(np.int32)
    gt_rpn_locs = (rs.standard_normal((n_anchor, 4)) * 0.3).astype(np.float32)

    want_losses, want_grads = om.train_step_grads(cfg, params, x, rois, idx, gt_roi_locs,
                                                  gt_roi_labels, gt_roi_masks, gt_rpn_locs,
                                                  gt_rpn_labels)
    g = TorchGraph(cfg, params)
    got = g.losses(_t(x), rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
                   gt_rpn_labels)
    for k, v in want_losses.items():
        assert abs(float(v) - got[k].item()) <= 2e-5 * max(abs(got[k].item()), 1e-2), k
    got['loss'].backward()
    trainable = [k for k in params if not om.is_frozen(k)]
    assert sorted(want_grads) == sorted(trainable)
    for k in trainable:
        a, b = want_grads[k].astype(np.float64), g.p[k].grad.numpy()
        assert a.shape == b.shape, k
        denom = max(np.linalg.norm(b), 1e-12)
        assert np.linalg.norm(a - b) / denom <= 2e-4, (k, np.linalg.norm(a - b) / denom)
        assert np.linalg.norm(b) > 0, k          # every trainable parameter receives a gradient
