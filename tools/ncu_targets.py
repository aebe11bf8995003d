"""Small driver for ncu captures: runs one kernel of the train step on its real shape a
few times.
  conv      res5 3x3 forward (50176 x 512 x 4608), fused affine + ReLU   [<256,4>, 8 epi warps]
  conv1x1   res5 conv3 (50176 x 2048 x 512), affine + residual + ReLU    [<256,3>, 12 epi warps]
  small     res4 3x3 (8568 x 256 x 2304)
  wgrad     res5 3x3 weight gradient, all 9 taps (512 x 512 over 50176 pixels)
  wgrad1x1  res5 conv3 weight gradient (2048 x 512 over 50176 pixels)
  roi       ROIAlign NHWC forward + backward, train shape (1024 RoIs, 2x51x84x1024, 7x7 bins)
  roi_cl    the drop-in operator's kernels (functions.roi_align_2d forward + backward):
            1000 RoIs on a 1x1024x50x68 map, 14x14 bins (BASELINE.json configs[4])"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from chainer_mask_rcnn_b200.models import engine as E  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'conv'
if what not in ('conv1x1', 'wgrad1x1'):
    what = what.rstrip('0123456789').rstrip('_') if what[-1].isdigit() else what      # 'roi2' = a second capture of target 'roi'
dev = 'cuda'
x = E.round_tf32(torch.randn((1024, 7, 7, 512), device=dev))
w = E.round_tf32(torch.randn((512, 3, 3, 512), device=dev) / 68.)
g = E.round_tf32(torch.randn((1024, 7, 7, 512), device=dev))
w3 = E.round_tf32(torch.randn((2048, 1, 1, 512), device=dev) / 22.)
res = torch.randn((1024, 7, 7, 2048), device=dev)
g3 = E.round_tf32(torch.randn((1024, 7, 7, 2048), device=dev))
xs = E.round_tf32(torch.randn((2, 51, 84, 256), device=dev))
ws = E.round_tf32(torch.randn((256, 3, 3, 256), device=dev) / 48.)
scale = torch.ones(2048, device=dev)
bias = torch.zeros(2048, device=dev)
gw = torch.zeros((512, 3, 3, 512), device=dev)
gw3 = torch.zeros((2048, 512), device=dev)
if what == 'roi':
    import synth
    rs = np.random.RandomState(0)
    feat = torch.randn((2, 51, 84, 1024), device=dev)
    rois = torch.from_numpy(synth.rois_xy(rs, 1024, 2, 800, 1333)).cuda()
if what == 'roi_cl':
    import synth
    from chainer_mask_rcnn_b200 import functions
    rs = np.random.RandomState(0)
    xm = torch.from_numpy(rs.standard_normal((1, 1024, 50, 68)).astype(np.float32)).cuda()
    xm = xm.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
    rois = torch.from_numpy(synth.rois_xy(rs, 1000, 1, 800, 1088)).cuda()
if what == 'roi_full':       # the channels-last kernels at the same shape as roi_cl
    import synth
    rs = np.random.RandomState(0)
    featf = torch.from_numpy(rs.standard_normal((1, 50, 68, 1024)).astype(np.float32)).cuda()
    rois = torch.from_numpy(synth.rois_xy(rs, 1000, 1, 800, 1088)).cuda()
for _ in range(4):
    if what == 'roi_full':
        y = E.roi_align_nhwc(featf, rois, 14, 14, 1, 1. / 16, round_out=False)
        E.roi_align_nhwc_bwd(y, rois, tuple(featf.shape), 14, 14, 1, 1. / 16)
    if what == 'roi_cl':
        y = functions.roi_align_2d(xm, rois, 14, 14, 1. / 16)
        y.backward(torch.ones_like(y))
        xm.grad = None
    if what == 'conv':
        E.conv_gemm(x, w, 512, 3, 3, 1, 1, scale=scale, bias=bias, relu=True)
    if what == 'conv1x1':
        E.conv_gemm(x, w3, 2048, scale=scale, bias=bias, addend=res, relu=True)
    if what == 'small':
        E.conv_gemm(xs, ws, 256, 3, 3, 1, 1, scale=scale, bias=bias, relu=True)
    if what == 'wgrad':
        E.wgrad_tap(g, x, gw, 512, 512, (7, 7), 9 * 512, x_off=(-1, -1), taps=(3, 3))
    if what == 'wgrad1x1':
        E.wgrad_tap(g3, x, gw3, 2048, 512, (7, 7), 512)
    if what == 'roi':
        y = E.roi_align_nhwc(feat, rois, 14, 14, 2, 1. / 16)
        E.roi_align_nhwc_bwd(y, rois, tuple(feat.shape), 14, 14, 2, 1. / 16)
torch.cuda.synchronize()
