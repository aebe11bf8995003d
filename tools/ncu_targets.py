"""Small driver for ncu captures: runs the two dominant tensor-core kernels on the
res5 3x3 shapes of the train step (50176 x 512 x 4608) a few times."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chainer_mask_rcnn_b200.models import engine as E  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'both'
x = E.round_tf32(torch.randn((1024, 7, 7, 512), device='cuda'))
w = E.round_tf32(torch.randn((512, 3, 3, 512), device='cuda') / 68.)
g = E.round_tf32(torch.randn((1024, 7, 7, 512), device='cuda'))
scale = torch.ones(512, device='cuda')
bias = torch.zeros(512, device='cuda')
gw = torch.zeros((512, 3, 3, 512), device='cuda')
for _ in range(4):
    if what in ('conv', 'both'):
        E.conv_gemm(x, w, 512, 3, 3, 1, 1, scale=scale, bias=bias, relu=True)
    if what in ('wgrad', 'both'):
        E.wgrad_tap(g, x, gw, 512, 512, (7, 7), 9 * 512, gw_col0=4 * 512, x_off=(0, 0))
torch.cuda.synchronize()
