"""Model-level parity: the CUDA graph (MaskRCNNResNet / MaskRCNNTrainChain) against
the NumPy oracle of the reference graph (oracle/model.py) on identical parameters and
inputs, at a width (base_channels=32) and image size the oracle finishes in seconds.

Tolerances.  The north star asks <= 1e-3 (max|delta| / max|ref|) per operator on
identical inputs; that is what the per-kernel tests assert.  Here whole stages are
chained (TF32 tensor-core products, fp32 accumulation, up to ~50 layers deep), so
end-to-end bounds are looser and are written next to each assertion.
"""
import json
import os

import numpy as np
import pytest
import torch

import synth
from chainer_mask_rcnn_b200 import models
from chainer_mask_rcnn_b200.models import engine as E
from oracle import model as om

pytestmark = pytest.mark.gpu

BASE = 32
N_FG = 5
SCALES = (4, 8, 16, 32)


def rel(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def nchw(t_nhwc):
    return t_nhwc.permute(0, 3, 1, 2).cpu().numpy()


@pytest.fixture(scope='module')
def setup():
    rs = np.random.RandomState(0)
    cfg = om.Config(n_layers=50, n_fg_class=N_FG, anchor_scales=SCALES, roi_size=14, base=BASE)
    params = om.make_params(cfg, rs)
    # non-trivial conv1 bias / head biases so that every epilogue term is exercised
    for k in params:
        if k.endswith('/b') and '/bn' not in k:
            params[k] = (rs.standard_normal(params[k].shape) * 0.05).astype(np.float32)
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE)
    model.load_state_dict(params)
    x = (rs.uniform(0, 255, (2, 3, 128, 160)) - 115.).astype(np.float32)
    return cfg, params, model, x, rs


def test_parameter_names_and_layouts_round_trip(setup):
    cfg, params, model, _, _ = setup
    sd = model.state_dict()
    assert sorted(sd) == sorted(params)
    for k in params:
        assert sd[k].shape == params[k].shape, k
        np.testing.assert_array_equal(sd[k], params[k])
    frozen = {k for k in params if om.is_frozen(k)}
    assert frozen == set(model.ctx.frozen.names())


def test_extractor_stages_match_oracle(setup):
    cfg, params, model, x, _ = setup
    tape = {}
    want, _ = om.extractor(cfg, params, x, tape)
    ctx = model.ctx
    ctx.prepare(backward=False)
    xt = torch.from_numpy(x).cuda()
    ex = model.extractor
    h = ex.conv1.forward(xt)
    assert rel(nchw(h), tape['extractor/conv1']) <= 1e-3            # one layer
    h = E.max_pool(h, 3, 2, 1)
    assert nchw(h).shape == tape['extractor/pool1'].shape            # cover_all output size
    assert rel(nchw(h), tape['extractor/pool1']) <= 1e-3
    h = ex.res2.forward(h)
    assert rel(nchw(h), tape['extractor/res2']) <= 2e-3             # 10 layers deep
    # each stage again on the oracle's own input: identical inputs, one stage deep
    for stage, blk in (('res3', ex.res3), ('res4', ex.res4)):
        prev = {'res3': 'res2', 'res4': 'res3'}[stage]
        src = torch.from_numpy(tape['extractor/' + prev]).cuda().permute(0, 2, 3, 1).contiguous()
        got = blk.forward(E.round_tf32(src))
        assert rel(nchw(got), tape['extractor/' + stage]) <= 2e-3
    got = model.extractor(x)
    assert tuple(got.shape) == want.shape
    assert rel(got, want) <= 5e-3                                    # 40 layers end to end


def test_call_matches_oracle_given_same_proposals(setup):
    """MaskRCNN.__call__: proposals are integer-exact given the same RPN outputs
    (tests/test_gpu_nms.py); here the RPN outputs differ at the 1e-3 level, so the head
    is compared on the model's own proposals."""
    cfg, params, model, x, _ = setup
    from chainer_mask_rcnn_b200.utils import config
    with config.using_config('train', False):
        cls_locs, scores, rois, roi_indices, masks = model(x, np.array([1., 1.], np.float32))
    assert cls_locs.shape[1] == 4 * (N_FG + 1) and scores.shape[1] == N_FG + 1
    assert masks.shape[1:] == (N_FG, 14, 14)
    R = rois.shape[0]
    assert roi_indices.shape == (R,) and roi_indices.dtype == torch.int32
    feat, _ = om.extractor(cfg, params, x)
    sel = np.linspace(0, R - 1, 24).astype(np.int64)
    w_cl, w_sc, w_m, _ = om.head_forward(cfg, params, feat, rois.cpu().numpy()[sel],
                                         roi_indices.cpu().numpy()[sel])
    assert rel(cls_locs[sel], w_cl) <= 1e-2
    assert rel(scores[sel], w_sc) <= 1e-2
    assert rel(masks[sel], w_m) <= 1e-2


def _targets(cfg, rs, x, n_anchor_total, n_roi=24):
    H, W = x.shape[2:]
    rois = synth.random_boxes(rs, n_roi, H, W, 12., 120.)
    idx = (np.arange(n_roi) % 2).astype(np.int32)
    gt_roi_locs = (rs.standard_normal((n_roi, 4)) * 0.5).astype(np.float32)
    gt_roi_labels = rs.randint(0, cfg.n_class, n_roi).astype(np.int32)
    gt_roi_labels[:4] = 0
    gt_roi_masks = rs.randint(0, 2, (n_roi, 14, 14)).astype(np.int32)
    gt_roi_masks[gt_roi_labels == 0] = -1
    gt_rpn_labels = rs.choice([-1, 0, 1], size=n_anchor_total, p=[0.8, 0.12, 0.08]).astype(np.int32)
    gt_rpn_locs = (rs.standard_normal((n_anchor_total, 4)) * 0.3).astype(np.float32)
    return rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs, gt_rpn_labels


def test_train_step_losses_and_gradients_match_oracle(setup):
    cfg, params, model, x, rs = setup
    feat, _ = om.extractor(cfg, params, x)
    n_anchor = feat.shape[2] * feat.shape[3] * cfg.n_anchor
    (rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
     gt_rpn_labels) = _targets(cfg, rs, x, 2 * n_anchor)
    want_losses, want_grads = om.train_step_grads(cfg, params, x, rois, idx, gt_roi_locs,
                                                  gt_roi_labels, gt_roi_masks, gt_rpn_locs,
                                                  gt_rpn_labels)
    chain = models.MaskRCNNTrainChain(model)
    ctx = model.ctx
    ctx.prepare(backward=True)
    ctx.recording = True
    xt = torch.from_numpy(x).cuda()
    f = model.extractor.forward_nhwc(xt)
    rpn_locs, rpn_scores, _, _, _, _ = model.rpn.forward_nhwc(f, x.shape[2:], np.ones(2))
    up = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    chain.cleargrads()
    loss = chain.forward_with_targets(f, rpn_locs, rpn_scores, up(rois), up(idx), up(gt_roi_locs),
                                      up(gt_roi_labels), up(gt_roi_masks), up(gt_rpn_locs),
                                      up(gt_rpn_labels))
    ctx.recording = False
    for k in ('rpn_loc_loss', 'rpn_cls_loss', 'roi_loc_loss', 'roi_cls_loss', 'roi_mask_loss'):
        got = float(chain.observation[k].item())
        assert abs(got - float(want_losses[k])) <= 2e-3 * max(abs(float(want_losses[k])), 1e-3), k
    assert abs(loss.item() - float(want_losses['loss'])) <= 2e-3 * float(want_losses['loss'])
    loss.backward()
    torch.cuda.synchronize()
    sd_names = set(ctx.train.names())
    assert set(want_grads) == sd_names
    worst, l2 = {}, {}
    for name in sorted(want_grads):
        kind = ctx.kinds[name][0]
        g = ctx.grad(name)
        if kind == 'conv':
            g = g.permute(0, 3, 1, 2)
        elif kind == 'deconv':
            g = g.permute(2, 1, 0).reshape(ctx.kinds[name][1])
        else:
            g = g.reshape(ctx.kinds[name][1])
        worst[name] = rel(g, want_grads[name])
        gn = g.detach().cpu().numpy().astype(np.float64)
        l2[name] = float(np.linalg.norm(gn - want_grads[name]) /
                         max(np.linalg.norm(want_grads[name]), 1e-30))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           'gpurun_out')
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, 'grad_errors.json'), 'w') as fjs:
            json.dump({'max_rel': worst, 'l2_rel': l2}, fjs, indent=1, sort_keys=True)
    # The backward pass chains ~35 TF32 GEMMs behind ~50 forward ones, and a forward
    # activation that differs by 1e-3 flips a few ReLU gates, which moves single
    # gradient entries by O(1) of their size: bound the relative L2 error at 3e-2 and
    # the worst entry at 5e-2 of max|grad| per tensor (measured: growing smoothly from
    # 6e-4 next to the losses to 2e-2 at res3).  The next test removes the gate flips
    # and holds the backward kernels to 3e-3.
    bad = {k: (worst[k], l2[k]) for k in worst if not (worst[k] <= 5e-2 and l2[k] <= 3e-2)}
    assert not bad, bad
    # the layers next to the losses are one or two GEMMs deep: hold them tighter
    for name in ('head/mask/W', 'head/cls_loc/W', 'head/score/W', 'rpn/loc/W', 'rpn/score/W',
                 'head/mask/b', 'head/cls_loc/b', 'rpn/loc/b'):
        assert worst[name] <= 5e-3, (name, worst[name])


def _np_nchw(t):
    return np.ascontiguousarray(t.detach().permute(0, 3, 1, 2).cpu().numpy())


def test_backward_on_identical_activations(setup):
    """Backward kernels alone: the oracle's backward is evaluated on the activations the
    CUDA forward produced (same ReLU gates, same GEMM inputs), so only the TF32 products
    and the accumulation order of the backward GEMMs differ."""
    from oracle import nn as onn
    cfg, params, model, x, rs = setup
    chain = models.MaskRCNNTrainChain(model)
    ctx = model.ctx
    ctx.prepare(backward=True)
    ctx.recording = True
    f = model.extractor.forward_nhwc(torch.from_numpy(x).cuda())
    rpn_locs, rpn_scores, _, _, _, _ = model.rpn.forward_nhwc(f, x.shape[2:], np.ones(2))
    n_anchor = f.shape[1] * f.shape[2] * cfg.n_anchor
    (rois, idx, gt_roi_locs, gt_roi_labels, gt_roi_masks, gt_rpn_locs,
     gt_rpn_labels) = _targets(cfg, np.random.RandomState(11), x, 2 * n_anchor)
    up = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    chain.cleargrads()
    loss = chain.forward_with_targets(f, rpn_locs, rpn_scores, up(rois), up(idx), up(gt_roi_locs),
                                      up(gt_roi_labels), up(gt_roi_masks), up(gt_rpn_locs),
                                      up(gt_rpn_labels))
    ctx.recording = False
    # ---- harvest the CUDA forward's activations into the oracle's cache format
    def block_caches(bb):
        return [tuple(_np_nchw(t) for t in b.saved) for b in bb.blocks]
    caches = {'res3': block_caches(model.extractor.res3), 'res4': block_caches(model.extractor.res4)}
    hs = model.head.saved
    res5_caches = block_caches(model.head.res5)
    pool7 = res5_caches[0][0]
    pool14 = np.zeros(pool7.shape[:2] + (14, 14), np.float32)
    pool14[:, :, ::2, ::2] = pool7
    res5_caches[0] = (pool14,) + res5_caches[0][1:]
    idx_rois = np.concatenate((idx.astype(np.float32)[:, None], rois), axis=1)
    hc = dict(idx_rois=idx_rois, res5=_np_nchw(hs['res5']),
              pool5=hs['pool5'].cpu().numpy()[:, :, None, None], d6=_np_nchw(hs['d6']),
              blocks=res5_caches, feat_shape=_np_nchw(f).shape)
    feat_np, h_np = (_np_nchw(t) for t in model.rpn.saved)
    o = chain.outputs
    _, g = om.train_losses(cfg, o['rpn_locs'].cpu().numpy(), o['rpn_scores'].cpu().numpy(),
                           gt_rpn_locs, gt_rpn_labels, o['roi_cls_locs'].cpu().numpy(),
                           o['roi_scores'].cpu().numpy(), _np_nchw(o['roi_masks']), gt_roi_locs,
                           gt_roi_labels, gt_roi_masks)
    want = {}
    g_feat = om.head_backward(cfg, params, hc, g['roi_cls_locs'], g['roi_scores'], g['roi_masks'],
                              want)
    g_feat = g_feat + om.rpn_backward(cfg, params, feat_np, h_np, g['rpn_locs'], g['rpn_scores'],
                                      want)
    om.extractor_backward(cfg, params, caches, g_feat, want)
    loss.backward()
    torch.cuda.synchronize()
    errs = {}
    for name in sorted(want):
        kind = ctx.kinds[name][0]
        gt = ctx.grad(name)
        if kind == 'conv':
            gt = gt.permute(0, 3, 1, 2)
        elif kind == 'deconv':
            gt = gt.permute(2, 1, 0).reshape(ctx.kinds[name][1])
        else:
            gt = gt.reshape(ctx.kinds[name][1])
        errs[name] = rel(gt, want[name])
    bad = {k: v for k, v in errs.items() if not v <= 3e-3}
    assert set(want) == set(ctx.train.names())
    assert not bad, bad


def test_sgd_update_matches_numpy(setup):
    cfg, params, model, x, rs = setup
    from chainer_mask_rcnn_b200 import optimizers
    ctx = model.ctx
    opt = optimizers.MomentumSGD(lr=0.01, momentum=0.9)
    opt.setup(models.MaskRCNNTrainChain(model))
    opt.add_hook(optimizers.WeightDecay(1e-4))
    g = torch.Generator(device='cuda').manual_seed(1)
    ctx.grads.copy_(torch.randn(ctx.grads.shape, device='cuda', generator=g))
    p0 = ctx.train.data.clone()
    frozen0 = ctx.frozen.data.clone()
    v = torch.zeros_like(p0)
    for _ in range(2):
        gg = ctx.grads + 1e-4 * p0
        v = 0.9 * v - 0.01 * gg
        p0 = p0 + v
        opt.update()
    assert float((ctx.train.data - p0).abs().max()) <= 1e-6
    assert torch.equal(ctx.frozen.data, frozen0)
    # the update kernel also refreshed the tf32 copy the forward GEMMs read
    assert torch.equal(ctx.rounded, E.round_tf32(ctx.train.data.clone()))
    model.load_state_dict(params)     # restore for other tests


@pytest.mark.parametrize('targets', ['device', 'host'])
def test_full_train_chain_runs_and_decreases_loss(targets):
    """End-to-end __call__ (tiny image, random targets) with the device target creators
    (default) and with the reference-order host ones: the loss is finite and a few SGD
    steps on a fixed batch reduce it."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(3)
    np.random.seed(3)
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE)
    if targets == 'host':
        chain = models.MaskRCNNTrainChain(
            model, anchor_target_creator=models.utils.AnchorTargetCreator(),
            proposal_target_creator=models.utils.ProposalTargetCreator())
    else:
        chain = models.MaskRCNNTrainChain(model)
    opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
    opt.add_hook(optimizers.WeightDecay(1e-4))
    H, W = 160, 192
    imgs = (rs.uniform(0, 255, (2, 3, H, W)) - 115.).astype(np.float32)
    bboxes, labels, masks = [], [], []
    for _ in range(2):
        b = synth.random_boxes(rs, 3, H, W, 40., 120.)
        b = b[(b[:, 2] - b[:, 0] > 8) & (b[:, 3] - b[:, 1] > 8)]
        m = np.zeros((len(b), H, W), np.int32)
        for i, (y1, x1, y2, x2) in enumerate(b.astype(int)):
            m[i, y1:y2, x1:x2] = 1
        bboxes.append(b); labels.append(rs.randint(0, N_FG, len(b)).astype(np.int32))
        masks.append(m)
    scales = np.ones(2, np.float32)
    hist = []
    for _ in range(6):
        np.random.seed(7)
        loss = opt.update(chain, imgs, bboxes, labels, masks, scales)
        hist.append(loss.item())
    assert all(np.isfinite(hist)), hist
    assert hist[-1] < hist[0], hist


def _tiny_batch(rs, H=160, W=192, G=4):
    imgs = (rs.uniform(0, 255, (2, 3, H, W)) - 115.).astype(np.float32)
    bboxes, labels, masks = [], [], []
    yy, xx = np.mgrid[:H, :W]
    for _ in range(2):
        b = synth.random_boxes(rs, 3, H, W, 40., 120.)
        b = b[(b[:, 2] - b[:, 0] > 8) & (b[:, 3] - b[:, 1] > 8)]
        m = np.zeros((G, H, W), np.int32)
        for i, (y1, x1, y2, x2) in enumerate(b):
            m[i] = (((yy - (y1 + y2) / 2) / ((y2 - y1) / 2)) ** 2 +
                    ((xx - (x1 + x2) / 2) / ((x2 - x1) / 2)) ** 2 <= 1)
        bboxes.append(b); labels.append(rs.randint(0, N_FG, len(b)).astype(np.int32))
        masks.append(m)
    return imgs, bboxes, labels, masks, np.ones(2, np.float32)


def test_device_mask_targets_equal_host_rasterisation_in_the_chain():
    """Same model, same seed: the chain fed tensor masks (cmr_mask_targets) and the chain
    fed the reference's host NumPy masks (cv2, IPP off = OpenCV's own bilinear code)
    produce identical mask targets and therefore the same loss."""
    import cv2
    rs = np.random.RandomState(5)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE)
    old = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        out = []
        for form in ('host', 'device'):
            chain = models.MaskRCNNTrainChain(model, seed=11)
            m = masks if form == 'host' else torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
            loss = chain(imgs, bboxes, labels, m, scales)
            out.append((loss.item(), chain.targets['gt_roi_masks'].cpu().numpy(),
                        chain.targets['gt_roi_labels'].cpu().numpy()))
    finally:
        cv2.ipp.setUseIPP(old)
    assert (out[0][2] > 0).sum() > 0                       # there are foreground rows
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])


def test_graphed_updater_replays_the_eager_step():
    """optimizers.GraphedUpdater: the captured step computes what the eager step computes.
    Run-to-run, the atomic weight-gradient reductions perturb the parameters by ~1e-7
    and the discrete RoI sampling of the next iteration amplifies that to ~2 % in the
    RoI losses (measured eager against eager, tools/graph_debug.py); so iteration 1 is
    compared tightly and the replayed iterations within that run-to-run spread."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(9)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    imgs_t = torch.from_numpy(imgs).cuda()
    masks_t = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
    results = []
    for use_graph in (False, True):
        model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                      base_channels=BASE, seed=1)
        chain = models.MaskRCNNTrainChain(model, seed=4)
        opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
        opt.add_hook(optimizers.WeightDecay(1e-4))
        up = optimizers.GraphedUpdater(opt, chain, max_boxes=8, use_graph=use_graph)
        hist = [up(imgs_t, bboxes, labels, masks_t, scales).item() for _ in range(4)]
        if use_graph:
            assert up.launches_per_replay > 100
        # a changed batch goes through the same graph (fixed input buffers are refilled)
        hist.append(up(imgs_t.flip(0), bboxes[::-1], labels[::-1], masks_t.flip(0), scales).item())
        results.append((hist, model.ctx.train.data.clone()))
    (h0, p0), (h1, p1) = results
    assert all(np.isfinite(h0 + h1))
    np.testing.assert_allclose(h1[0], h0[0], rtol=1e-5)
    np.testing.assert_allclose(h1, h0, rtol=5e-2)
    assert float((p1 - p0).abs().max()) <= 2e-3 * float(p0.abs().max())
    assert h0[3] < h0[0] and h1[3] < h1[0]


def test_prefetch_step_matches_direct_call():
    """GraphedUpdater.prefetch() + step() (inputs copied on a side stream into staging
    buffers, then device-to-device into the graph's inputs) runs the same iterations as
    calling the updater with the inputs directly -- from pinned host memory too."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(12)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    imgs_p = torch.from_numpy(imgs).pin_memory()
    masks_p = models.utils.PackedMasks.from_numpy(np.stack(masks), pin=True)
    hists = []
    for mode in ('direct', 'prefetch'):
        model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                      base_channels=BASE, seed=2)
        chain = models.MaskRCNNTrainChain(model, seed=6)
        opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
        up = optimizers.GraphedUpdater(opt, chain, max_boxes=8)
        hist = []
        if mode == 'prefetch':
            up.prefetch(imgs_p, bboxes, labels, masks_p, scales)
        for _ in range(4):
            if mode == 'direct':
                loss = up(imgs_p, bboxes, labels, masks_p, scales)
            else:
                loss = up.step()
                up.prefetch(imgs_p, bboxes, labels, masks_p, scales)
            hist.append(loss.item())
        hists.append(hist)
    assert all(np.isfinite(hists[0] + hists[1]))
    np.testing.assert_allclose(hists[1][0], hists[0][0], rtol=1e-5)
    np.testing.assert_allclose(hists[1], hists[0], rtol=5e-2)
    with pytest.raises(RuntimeError):
        up.step(); up.step()


def test_ragged_image_size_train_and_predict():
    """Image sizes that are not multiples of the feature stride (cover_all pooling, ragged
    tiles in every GEMM, im2col TMA boxes crossing image borders): a train step runs, its
    loss is finite and decreases, and predict() returns masks of the original sizes."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(21)
    H, W = 150, 205
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE, min_size=150, max_size=400)
    chain = models.MaskRCNNTrainChain(model, seed=3)
    opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
    imgs = (rs.uniform(0, 255, (3, 3, H, W)) - 115.).astype(np.float32)
    bboxes, labels, masks = [], [], []
    for _ in range(3):
        b = synth.random_boxes(rs, 4, H, W, 30., 120.)
        b = b[(b[:, 2] - b[:, 0] > 8) & (b[:, 3] - b[:, 1] > 8)]
        m = np.zeros((4, H, W), np.uint8)
        for i, (y1, x1, y2, x2) in enumerate(b.astype(int)):
            m[i, y1:y2, x1:x2] = 1
        bboxes.append(b); labels.append(rs.randint(0, N_FG, len(b)).astype(np.int32))
        masks.append(m)
    masks_t = torch.from_numpy(np.stack(masks)).cuda()
    hist = [opt.update(chain, imgs, bboxes, labels, masks_t, np.ones(3, np.float32)).item()
            for _ in range(5)]
    assert all(np.isfinite(hist)) and hist[-1] < hist[0], hist
    model.score_thresh = 1e-3
    raw = [rs.uniform(0, 255, (3, 97, 131)).astype(np.float32),
           rs.uniform(0, 255, (3, 150, 101)).astype(np.float32)]
    bb, mm, ll, ss = model.predict(raw)
    for i, im in enumerate(raw):
        assert mm[i].shape == (len(bb[i]),) + im.shape[1:]
        assert np.isfinite(bb[i]).all() and np.isfinite(ss[i]).all()


def test_reference_format_numpy_batch_replays_the_graph():
    """The reference's own batch (datasets.concat_examples: imgs float32, masks (B,G,H,W)
    int32 NumPy arrays on the host, page-locked or not) goes through GraphedUpdater's graph
    replay and gives the losses of the same batch passed as device tensors / packed bits;
    MaskRCNNTrainChain.__call__ takes the same arrays on the device-target fast path."""
    from chainer_mask_rcnn_b200 import datasets, optimizers
    rs = np.random.RandomState(31)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    examples = [(imgs[i], bboxes[i], labels[i], masks[i][:len(bboxes[i])], 1.0) for i in range(2)]
    hists = {}
    for form in ('tensor', 'numpy', 'pinned'):
        model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                      base_channels=BASE, seed=3)
        chain = models.MaskRCNNTrainChain(model, seed=8)
        opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
        up = optimizers.GraphedUpdater(opt, chain, max_boxes=8)
        if form == 'tensor':
            a_i = torch.from_numpy(imgs).cuda()
            a_m = torch.from_numpy(np.stack(masks).astype(np.uint8)[:, :3]).cuda()
        else:
            a_i, _, _, a_m, _ = datasets.concat_examples(
                examples, padding=0, indices_concat=[0, 2, 3, 4], indices_to_device=[],
                pinned=form == 'pinned')
            assert isinstance(a_m, np.ndarray) and a_m.dtype == np.int32 and a_m.ndim == 4
            assert torch.from_numpy(a_m).is_pinned() == (form == 'pinned')
        hists[form] = [up(a_i, bboxes, labels, a_m, scales).item() for _ in range(4)]
        assert up.launches_per_replay > 100
        if form == 'numpy':         # the chain itself on the reference-format arrays
            chain2 = models.MaskRCNNTrainChain(model, seed=8)
            loss = chain2(a_i, bboxes, labels, a_m, scales)
            assert np.isfinite(loss.item()) and chain2.d2h_bytes == 0
    assert all(np.isfinite(sum(hists.values(), [])))
    for form in ('numpy', 'pinned'):
        np.testing.assert_allclose(hists[form][0], hists['tensor'][0], rtol=1e-5)
        np.testing.assert_allclose(hists[form], hists['tensor'], rtol=5e-2)


def test_graphed_updater_bounds_its_states_and_sees_reloaded_weights():
    """ At most max_states (batch geometry) keys are kept -- least recently used
    first out, its graph and buffers freed; the random draws continue through ONE seed word
    across keys; weights loaded after a graph was captured reach the next replay."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(41)
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE, seed=5)
    chain = models.MaskRCNNTrainChain(model, seed=2)
    opt = optimizers.MomentumSGD(lr=0.0, momentum=0.0).setup(chain)     # parameters stay put
    up = optimizers.GraphedUpdater(opt, chain, max_boxes=8, max_states=2)
    batches = {}
    for k, (H, W) in enumerate(((160, 192), (128, 160), (144, 176))):
        imgs, bboxes, labels, masks, scales = _tiny_batch(np.random.RandomState(50 + k), H, W)
        batches[k] = (torch.from_numpy(imgs).cuda(), bboxes, labels,
                      torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda(), scales)
    for k in (0, 1, 0, 1):
        up(*batches[k])
    assert len(up._states) == 2 and up.evictions == 0
    assert int(up._seed_word.item()) == 4                   # one word across both keys
    torch.cuda.synchronize()
    up(*batches[2])                                          # third geometry: evicts key 0's
    assert len(up._states) == 2 and up.evictions == 1
    assert [k[0][2:] for k in up._states] == [(128, 160), (144, 176)]
    # scales are not part of the key when the proposal layer ignores them (min_size = 0)
    n = len(up._states)
    up(batches[2][0], batches[2][1], batches[2][2], batches[2][3], np.array([1.3, 0.7]))
    assert len(up._states) == n
    # reload: zero the RPN score layer -> the rpn_cls loss becomes log(2) on the next REPLAY
    st = list(up._states.values())[-1]
    assert st.graph is not None
    sd = model.state_dict()
    sd['rpn/score/W'][:] = 0
    sd['rpn/score/b'][:] = 0
    model.load_state_dict(sd)
    up(*batches[2])
    assert abs(float(chain.observation['rpn_cls_loss'].item()) - np.log(2.)) < 1e-4


def test_rpn_branch_on_side_stream_gives_the_same_step():
    """The RPN branch forked onto a side stream next to the proposal chain -- losses only
    (default), and with its backward pass run ahead (what optimizer.update / GraphedUpdater
    do) -- computes the same losses and gradients as the single-stream schedule."""
    rs = np.random.RandomState(17)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    imgs_t = torch.from_numpy(imgs).cuda()
    masks_t = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                  base_channels=BASE, seed=4)
    out = {}
    for mode in ('serial', 'side_losses', 'side_backward'):
        chain = models.MaskRCNNTrainChain(model, seed=9)
        chain.overlap_rpn_branch = mode != 'serial'
        chain.eager_rpn_backward = mode == 'side_backward'
        chain.cleargrads()
        loss = chain(imgs_t, bboxes, labels, masks_t, scales)
        obs = {k: float(v.item()) for k, v in chain.observation.items()}
        loss.backward()
        torch.cuda.synchronize()
        out[mode] = (obs, model.ctx.grads.clone())
    for mode in ('side_losses', 'side_backward'):
        for k, v in out['serial'][0].items():
            assert abs(out[mode][0][k] - v) <= 1e-6 * max(abs(v), 1e-3), (mode, k)
        g0, g1 = out['serial'][1], out[mode][1]
        # atomic reductions (weight gradients; the RoI head's gradient added to the RPN's
        # or the other way round) -> equal up to summation order, and the tf32 rounding of the
        # feature-map gradient turns a last-bit difference into 2^-11 of single entries
        assert float((g1 - g0).abs().max()) <= 2e-5 * float(g0.abs().max()), mode
        for name in ('rpn/conv1/W', 'rpn/loc/W', 'extractor/res4/b5/conv3/W',
                     'extractor/res3/a/conv1/W'):
            a, b = model.ctx.train.view(name, g0), model.ctx.train.view(name, g1)
            assert float((a - b).abs().max()) <= 1e-3 * float(a.abs().max()), (mode, name)


def test_deterministic_mode_is_bit_reproducible():
    """MaskRCNNTrainChain(deterministic=True): split weight gradients, bias sums and the
    ROIAlign backward accumulate in 64-bit fixed point, so two runs of the same graph-replayed
    steps from the same initial state end with BIT-IDENTICAL parameters (the default mode's
    atomic fp32 reductions differ from run to run in the last bits); and the deterministic
    gradients are the default mode's up to that rounding."""
    from chainer_mask_rcnn_b200 import optimizers
    rs = np.random.RandomState(23)
    imgs, bboxes, labels, masks, scales = _tiny_batch(rs)
    imgs_t = torch.from_numpy(imgs).cuda()
    masks_t = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()

    def run(deterministic, steps=4):
        model = models.MaskRCNNResNet(50, N_FG, anchor_scales=SCALES, roi_size=14,
                                      base_channels=BASE, seed=6)
        chain = models.MaskRCNNTrainChain(model, seed=10, deterministic=deterministic)
        opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
        opt.add_hook(optimizers.WeightDecay(1e-4))
        up = optimizers.GraphedUpdater(opt, chain, max_boxes=8)
        hist = [up(imgs_t, bboxes, labels, masks_t, scales).item() for _ in range(steps)]
        torch.cuda.synchronize()
        return hist, model.ctx.train.data.clone(), model.ctx.grads.clone()

    h1, p1, g1 = run(True)
    h2, p2, g2 = run(True)
    assert all(np.isfinite(h1)) and h1[-1] < h1[0]
    assert torch.equal(p1, p2) and torch.equal(g1, g2)          # bit-identical
    h0, p0, g0 = run(False, steps=1)
    hd, pd, gd = run(True, steps=1)
    np.testing.assert_allclose(hd[0], h0[0], rtol=1e-5)
    assert float((gd - g0).abs().max()) <= 2e-5 * float(g0.abs().max())
    assert float((pd - p0).abs().max()) <= 1e-6 * float(p0.abs().max())
