"""Eager vs eager vs graph: per-iteration losses and parameter differences on the tiny model."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa
from chainer_mask_rcnn_b200 import models, optimizers
import test_gpu_model as T

rs = np.random.RandomState(9)
imgs, bboxes, labels, masks, scales = T._tiny_batch(rs)
imgs_t = torch.from_numpy(imgs).cuda()
masks_t = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
res = []
for mode in ('eager', 'eager', 'graph'):
    model = models.MaskRCNNResNet(50, T.N_FG, anchor_scales=T.SCALES, roi_size=14, base_channels=T.BASE, seed=1)
    chain = models.MaskRCNNTrainChain(model, seed=4)
    opt = optimizers.MomentumSGD(lr=0.002, momentum=0.9).setup(chain)
    opt.add_hook(optimizers.WeightDecay(1e-4))
    up = optimizers.GraphedUpdater(opt, chain, max_boxes=8, use_graph=(mode == 'graph'))
    hist, ps, gs, obs = [], [], [], []
    for i in range(4):
        hist.append(up(imgs_t, bboxes, labels, masks_t, scales).item())
        ps.append(model.ctx.train.data.clone()); gs.append(model.ctx.grads.clone())
        obs.append([float(chain.observation[k]) for k in ('rpn_loc_loss', 'rpn_cls_loss', 'roi_loc_loss', 'roi_cls_loss', 'roi_mask_loss')])
    res.append((hist, ps, gs, obs))
    print(mode, hist)
for a, b, name in ((0, 1, 'eager-eager'), (0, 2, 'eager-graph')):
    for i in range(4):
        pa, pb = res[a][1][i], res[b][1][i]
        ga, gb = res[a][2][i], res[b][2][i]
        print(name, 'iter', i, 'param rel', float((pa - pb).abs().max() / pa.abs().max()),
              'grad rel', float((ga - gb).abs().max() / ga.abs().max()), 'losses', res[a][3][i], res[b][3][i])
