"""Per-launch timing of one full-size train step: every tensor-core launch is bracketed
by CUDA events (Python-level hook), then grouped by GEMM shape.  With --graph every
distinct shape is additionally re-launched 8 times inside a CUDA graph (no launch gaps,
operands L2-warm as in the real step) and the replay time / 8 is reported as `g_ms`.
Usage: python tools/layer_bench.py [--out FILE] [--graph]"""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chainer_mask_rcnn_b200 import models, optimizers  # noqa: E402
from chainer_mask_rcnn_b200.models import engine as E  # noqa: E402

records = []
replays = {}
enabled = [False]
_conv, _wgrad = E.conv_gemm, E.wgrad_tap


def conv_hook(x, w, n, kh=1, kw=1, stride=1, pad=0, **kw_):
    if not enabled[0]:
        return _conv(x, w, n, kh, kw, stride, pad, **kw_)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = _conv(x, w, n, kh, kw, stride, pad, **kw_)
    b.record()
    B = x.shape[0]
    oh, ow = kw_.get('out_hw') or (E.conv_out(x.shape[1], kh, stride, pad),
                                   E.conv_out(x.shape[2], kw, stride, pad))
    K = kh * kw * (kw_.get('in_c') or x.shape[3])
    flags = ''.join(c for c, k in (('s', 'scale'), ('b', 'bias'), ('a', 'addend'), ('m', 'mask'))
                    if kw_.get(k) is not None) + ('r' if kw_.get('relu') else '')
    key = ('gemm', B * oh * ow, n, K, '%dx%d s%d d%d %s' % (kh, kw, stride,
                                                           kw_.get('d_stride', 1), flags))
    records.append((key[0], key[1:], a, b, 2.0 * B * oh * ow * n * K))
    if key not in replays:
        kw2 = dict(kw_)
        kw2['out'] = out
        replays[key] = lambda: _conv(x, w, n, kh, kw, stride, pad, **kw2)
    return out


def wgrad_hook(gy, x, gw, rows, cols, loop_hw, gw_ld, **kw_):
    if not enabled[0]:
        return _wgrad(gy, x, gw, rows, cols, loop_hw, gw_ld, **kw_)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _wgrad(gy, x, gw, rows, cols, loop_hw, gw_ld, **kw_)
    b.record()
    M = gy.shape[0] * loop_hw[0] * loop_hw[1]
    t = kw_.get('taps', (1, 1))
    key = ('wgrad', M, rows, cols, 'taps%d' % (t[0] * t[1]))
    records.append((key[0], key[1:], a, b, 2.0 * M * rows * cols * t[0] * t[1]))
    if key not in replays:
        scratch = torch.zeros_like(gw)
        replays[key] = lambda: _wgrad(gy, x, scratch, rows, cols, loop_hw, gw_ld, **kw_)


E.conv_gemm, E.wgrad_tap = conv_hook, wgrad_hook


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--graph', action='store_true')
    args = ap.parse_args()
    model = models.MaskRCNNResNet(50, 80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333)
    chain = models.MaskRCNNTrainChain(model)
    opt = optimizers.MomentumSGD(lr=0.0025).setup(chain)
    imgs, bboxes, labels, masks, scales = bench.synth_batch(0)
    masks = torch.from_numpy(np.stack(masks).astype(np.uint8)).cuda()
    x = torch.from_numpy(imgs).cuda()
    for _ in range(3):
        opt.update(chain, x, bboxes, labels, masks, scales)
    torch.cuda.synchronize()
    enabled[0] = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.update(chain, x, bboxes, labels, masks, scales)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1)
    agg = collections.OrderedDict()
    for kind, shape, a, b, flops in records:
        ms = a.elapsed_time(b)
        r = agg.setdefault((kind,) + shape, [0, 0.0, 0.0])
        r[0] += 1; r[1] += ms; r[2] += flops
    enabled[0] = False
    g_ms = {}
    if args.graph:
        reps = 8
        for key, fn in replays.items():
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    fn()
            g.replay()
            best = 1e9
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / reps)
            g_ms[key] = best
            del g
    rows = []
    for k, (n, ms, fl) in agg.items():
        r = dict(kind=k[0], M=k[1], N=k[2], K=k[3], tag=k[4], launches=n, ms=ms,
                 tflops=fl / ms / 1e9)
        if k in g_ms:
            r['g_ms'] = g_ms[k] * n
            r['g_tflops'] = fl / (g_ms[k] * n) / 1e9
        rows.append(r)
    rows.sort(key=lambda r: -r.get('g_ms', r['ms']))
    tot = {kind: sum(r['ms'] for r in rows if r['kind'] == kind) for kind in ('gemm', 'wgrad')}
    print('step %.2f ms (instrumented); gemm %.2f ms, wgrad %.2f ms' % (step_ms, tot['gemm'],
                                                                     tot['wgrad']))
    if g_ms:
        gt = {kind: sum(r['g_ms'] for r in rows if r['kind'] == kind) for kind in ('gemm', 'wgrad')}
        print('in-graph: gemm %.2f ms, wgrad %.2f ms' % (gt['gemm'], gt['wgrad']))
        tot['gemm_graph'], tot['wgrad_graph'] = gt['gemm'], gt['wgrad']
    for r in rows:
        print('%-5s M=%-7d N=%-5d K=%-5d %-16s x%-3d %7.3f ms %7.1f TF/s | graph %7.3f ms %7.1f TF/s' % (
            r['kind'], r['M'], r['N'], r['K'], r['tag'], r['launches'], r['ms'], r['tflops'],
            r.get('g_ms', 0.), r.get('g_tflops', 0.)))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(dict(step_ms=step_ms, totals=tot, rows=rows), f, indent=1)


if __name__ == '__main__':
    main()
