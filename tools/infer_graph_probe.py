"""Is the inference box pass launch-bound?  Times MaskRCNN._forward_padded eagerly (wall
clock around a synchronise) and as a CUDA-graph replay, batch 1 and 2, 800x1333."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chainer_mask_rcnn_b200 import models  # noqa: E402
from chainer_mask_rcnn_b200.utils import config  # noqa: E402

model = models.MaskRCNNResNet(50, 80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                              min_size=800, max_size=1333)
for bs in (1, 2):
    x = torch.randn((bs, 3, 800, 1333), device='cuda') * 60
    scales = np.ones(bs, np.float32)

    def fwd():
        with config.using_config('train', False), torch.no_grad():
            return model._forward_padded(x, scales, False)
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        fwd()
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t) * 100
    t = time.perf_counter()
    for _ in range(10):
        fwd()
    cpu_only = (time.perf_counter() - t) * 100
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fwd()
        with torch.cuda.graph(g, stream=s):
            out = fwd()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    graph = (time.perf_counter() - t) * 100
    print('batch %d: eager %.3f ms (host launch time alone %.3f ms), graph replay %.3f ms' % (
        bs, eager, cpu_only, graph))
