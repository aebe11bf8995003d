"""Timing probes of the weight-gradient kernel on the res5 3x3 shape (all 9 taps):
CMR_WGRAD_PROBE=0 normal, 1 = no loads after the first ring fill (MMA-only rate),
2 = no MMAs (load-only rate).  Prints ms per launch inside a CUDA graph of 8 launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chainer_mask_rcnn_b200.models import engine as E  # noqa: E402

dev = 'cuda'
x = E.round_tf32(torch.randn((1024, 7, 7, 512), device=dev))
g = E.round_tf32(torch.randn((1024, 7, 7, 512), device=dev))
g3 = E.round_tf32(torch.randn((1024, 7, 7, 2048), device=dev))
gw = torch.zeros((512, 3, 3, 512), device=dev)
gw3 = torch.zeros((2048, 512), device=dev)
cases = {
    '3x3 taps9 512x512': (lambda: E.wgrad_tap(g, x, gw, 512, 512, (7, 7), 9 * 512, x_off=(-1, -1),
                                              taps=(3, 3)), 2.0 * 50176 * 512 * 512 * 9),
    '1x1 2048x512': (lambda: E.wgrad_tap(g3, x, gw3, 2048, 512, (7, 7), 512),
                     2.0 * 50176 * 2048 * 512),
}
for probe in sys.argv[1:] or ['0', '1', '2']:
    os.environ['CMR_WGRAD_PROBE'] = probe
    for name, (fn, flops) in cases.items():
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(8):
                fn()
        gr.replay()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 8)
        print('probe %s  %-20s %.3f ms  %.1f TF/s-equivalent' % (probe, name, best,
                                                                 flops / best / 1e9))
