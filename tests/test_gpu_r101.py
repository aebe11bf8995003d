"""ResNet-101-C4 (BASELINE.json configs[3]'s backbone: res4 has 23 blocks) against the
oracle: every extractor layer on the oracle's own input (north-star bound 1e-3), the
chained extractor, and one train step -- losses, and the backward kernels on identical
activations -- through the same code paths as test_gpu_model.py runs for R50.  Reduced
width (base_channels = 32) and image size so that the NumPy oracle finishes in seconds;
the block count, strides and layer order are the full R101's."""
import numpy as np
import pytest
import torch

import test_gpu_config0 as c0
import test_gpu_model as tm
from chainer_mask_rcnn_b200 import models, optimizers
from chainer_mask_rcnn_b200.models import engine as E
from chainer_mask_rcnn_b200.models.resnet_extractor import N_BLOCKS
from oracle import model as om

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def setup():
    rs = np.random.RandomState(101)
    cfg = om.Config(n_layers=101, n_fg_class=tm.N_FG, anchor_scales=tm.SCALES, roi_size=14,
                    base=tm.BASE)
    params = om.make_params(cfg, rs)
    for k in params:
        if k.endswith('/b') and '/bn' not in k:
            params[k] = (rs.standard_normal(params[k].shape) * 0.05).astype(np.float32)
    model = models.MaskRCNNResNet(101, tm.N_FG, anchor_scales=tm.SCALES, roi_size=14,
                                  base_channels=tm.BASE)
    model.load_state_dict(params)
    x = (rs.uniform(0, 255, (2, 3, 128, 160)) - 115.).astype(np.float32)
    return cfg, params, model, x, rs


def test_structure_and_parameter_names(setup):
    cfg, params, model, _, _ = setup
    assert N_BLOCKS[101] == (3, 4, 23)
    assert len(model.extractor.res4.blocks) == 23 and len(model.head.res5.blocks) == 3
    assert 'extractor/res4/b22/conv3/W' in params
    sd = model.state_dict()
    assert sorted(sd) == sorted(params)
    n_conv = sum(1 for k in sd if k.startswith('extractor/') and '/conv' in k and k.endswith('/W'))
    assert n_conv == 1 + 3 * (3 + 4 + 23) + 3                     # 94 convolutions
    # 54.6 M trainable parameters at full width scale as base^2: structure check at base 32
    frozen = {k for k in params if om.is_frozen(k)}
    assert frozen == set(model.ctx.frozen.names())


def test_extractor_every_layer_on_identical_inputs(setup):
    cfg, params, model, x, _ = setup
    tape = {}
    want, _ = om.extractor(cfg, params, x, tape)
    model.ctx.prepare(backward=False)
    errs = {}
    h = model.extractor.conv1.forward(torch.from_numpy(x).cuda())
    errs['conv1'] = c0.rel(c0.nchw(h), tape['extractor/conv1'])
    for stage in ('res2', 'res3', 'res4'):
        bb = getattr(model.extractor, stage)
        for nm, blk in zip(bb.names, bb.blocks):
            c0._check_block(blk, tape['extractor/%s/%s' % (stage, nm)], errs,
                            '%s/%s' % (stage, nm))
    assert len(errs) == 1 + 30 * 4 + 3
    bad = {k: v for k, v in errs.items() if not v <= 1e-3}
    assert not bad, bad
    got = model.extractor(x)
    assert tuple(got.shape) == want.shape
    assert c0.rel(got, want) <= 1e-2                              # 91 TF32 layers chained


def test_train_step_losses_match_oracle(setup):
    cfg, params, model, x, rs = setup
    feat, _ = om.extractor(cfg, params, x)
    n_anchor = feat.shape[2] * feat.shape[3] * cfg.n_anchor
    (rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
     gt_rpn_labels) = tm._targets(cfg, rs, x, 2 * n_anchor)
    want_losses, _ = om.train_step_grads(cfg, params, x, rois, idx, gt_roi_locs, gt_roi_labels,
                                         gt_roi_masks, gt_rpn_locs, gt_rpn_labels)
    chain = models.MaskRCNNTrainChain(model)
    ctx = model.ctx
    ctx.prepare(backward=True)
    ctx.recording = True
    f = model.extractor.forward_nhwc(torch.from_numpy(x).cuda())
    rpn_locs, rpn_scores, _, _, _, _ = model.rpn.forward_nhwc(f, x.shape[2:], np.ones(2))
    up = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    chain.cleargrads()
    loss = chain.forward_with_targets(f, rpn_locs, rpn_scores, up(rois), up(idx), up(gt_roi_locs),
                                      up(gt_roi_labels), up(gt_roi_masks), up(gt_rpn_locs),
                                      up(gt_rpn_labels))
    ctx.recording = False
    for k in ('rpn_loc_loss', 'rpn_cls_loss', 'roi_loc_loss', 'roi_cls_loss', 'roi_mask_loss'):
        got = float(chain.observation[k].item())
        assert abs(got - float(want_losses[k])) <= 3e-3 * max(abs(float(want_losses[k])), 1e-3), k
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(ctx.grads).all()


def test_backward_on_identical_activations(setup):
    """All 33 residual blocks' data- and weight-gradient GEMMs against the oracle's backward
    evaluated on the CUDA forward's activations: <= 3e-3 per gradient tensor."""
    tm.test_backward_on_identical_activations(setup)


def test_graphed_train_steps_decrease_the_loss():
    rs = np.random.RandomState(9)
    imgs, bboxes, labels, masks, scales = tm._tiny_batch(rs)
    model = models.MaskRCNNResNet(101, tm.N_FG, anchor_scales=tm.SCALES, roi_size=14,
                                  base_channels=tm.BASE, seed=1)
    chain = models.MaskRCNNTrainChain(model, seed=4)
    opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
    opt.add_hook(optimizers.WeightDecay(1e-4))
    up = optimizers.GraphedUpdater(opt, chain, max_boxes=8)
    masks_t = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
    hist = [up(torch.from_numpy(imgs).cuda(), bboxes, labels, masks_t, scales).item()
            for _ in range(5)]
    assert all(np.isfinite(hist)) and hist[-1] < hist[0], hist
    assert up.launches_per_replay > 300          # 33 blocks x (3 fwd + 3 dgrad + 3 wgrad) + ...
