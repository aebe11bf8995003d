"""Same-box A/B timing of single conv_gemm_tc launches of the train step under the tile
configurations of cmr_set_conv_variant.  Each (shape, variant) is replayed as a CUDA graph of
`reps` launches over rotating operand sets (together larger than L2), timed with CUDA events.
Usage: python tools/conv_shape_bench.py [--out FILE] [--variants 0,1,2,3] [--shapes a,b,...]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chainer_mask_rcnn_b200 import _lib  # noqa: E402
from chainer_mask_rcnn_b200.models import engine as E  # noqa: E402

# name: (B, H, W, C_in, N, kh, pad, flags)   flags: s scale, b bias, a addend, m mask, r relu
SHAPES = {
    'res5_conv3_fwd': (1024, 7, 7, 512, 2048, 1, 0, 'sbar'),
    'res5_conv1_dgrad': (1024, 7, 7, 512, 2048, 1, 0, 'am'),
    'res5_conv3_dgrad': (1024, 7, 7, 2048, 512, 1, 0, 'm'),
    'res5_conv1_fwd': (1024, 7, 7, 2048, 512, 1, 0, 'sbr'),
    'res5_3x3_fwd': (1024, 7, 7, 512, 512, 3, 1, 'sbr'),
    'res5_3x3_dgrad': (1024, 7, 7, 512, 512, 3, 1, 'm'),
    'res4_conv3_fwd': (2, 51, 84, 256, 1024, 1, 0, 'sbar'),
    'res4_conv1_fwd': (2, 51, 84, 1024, 256, 1, 0, 'sbr'),
    'res4_3x3_fwd': (2, 51, 84, 256, 256, 3, 1, 'sbr'),
    'res3_conv3_fwd': (2, 101, 167, 128, 512, 1, 0, 'sbar'),
    'res3_conv1_fwd': (2, 101, 167, 512, 128, 1, 0, 'sbr'),
    'res3_3x3_fwd': (2, 101, 167, 128, 128, 3, 1, 'sbr'),
    'res2_conv3_fwd': (2, 201, 334, 64, 256, 1, 0, 'sbar'),
    'res2_3x3_fwd': (2, 201, 334, 64, 64, 3, 1, 'sbr'),
}


def run_shape(name, variant, reps=8, iters=5):
    B, H, W, C, N, k, pad, flags = SHAPES[name]
    M = B * H * W
    K = k * k * C
    per_set = 4 * (M * C + M * N * (1 + ('a' in flags) + ('m' in flags)))
    n_sets = max(2, min(reps, -(-(300 << 20) // per_set)))
    dev = 'cuda'
    sets = []
    for _ in range(n_sets):
        x = torch.randn((B, H, W, C), device=dev)
        E.round_tf32(x, x)
        out = torch.empty((B, H, W, N), device=dev)
        addend = torch.randn((B, H, W, N), device=dev) if 'a' in flags else None
        mask = torch.randn((B, H, W, N), device=dev) if 'm' in flags else None
        sets.append((x, out, addend, mask))
    w = torch.randn((N, k, k, C), device=dev) * 0.05
    E.round_tf32(w, w)
    scale = torch.rand((N,), device=dev) + 0.5 if 's' in flags else None
    bias = torch.randn((N,), device=dev) if 'b' in flags else None
    _lib.load().cmr_set_conv_variant(variant)

    def launch(i):
        x, out, addend, mask = sets[i % n_sets]
        E.conv_gemm(x, w, N, k, k, 1, pad, out=out, scale=scale, bias=bias, addend=addend,
                    mask=mask, relu='r' in flags)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n_sets):
            launch(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / reps
        best = t if best is None else min(best, t)
    dbg = None
    if os.environ.get('CONV_DEBUG'):
        buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        _lib.call('cmr_set_conv_debug', _lib.ptr(buf))
        launch(0)
        torch.cuda.synchronize()
        _lib.call('cmr_set_conv_debug', None)
        d = buf.view(148, 8).cpu().double()
        lead = d[d[:, 6] > 0]
        dbg = {'mma_wait_acc': float(lead[:, 0].mean()), 'mma_wait_ops': float(lead[:, 1].mean()),
               'mma_total': float(lead[:, 2].mean()), 'tma_wait_free': float(d[:, 3].mean()),
               'epi_wait_acc': float(d[:, 4].mean()), 'epi_total': float(d[:, 5].mean()),
               'tiles': float(lead[:, 6].mean())}
    _lib.load().cmr_set_conv_variant(0)
    flops = 2.0 * M * N * K
    nbytes = per_set + 4 * N * K
    return {'shape': name, 'variant': variant, 'M': M, 'N': N, 'K': K, 'flags': flags,
            'dbg': dbg, 'us': best * 1e3, 'tflops': flops / best / 1e9, 'gbs': nbytes / best / 1e6}


# name: (B, H, W, cout(rows), cin(cols), k, pad)   weight gradient of a stride-1 convolution
WGRAD_SHAPES = {
    'w_res5_3x3': (1024, 7, 7, 512, 512, 3, 1),
    'w_res5_conv3': (1024, 7, 7, 2048, 512, 1, 0),
    'w_res5_conv1': (1024, 7, 7, 512, 2048, 1, 0),
    'w_rpn_3x3': (2, 51, 84, 1024, 1024, 3, 1),
    'w_res4_3x3': (2, 51, 84, 256, 256, 3, 1),
    'w_res4_conv3': (2, 51, 84, 1024, 256, 1, 0),
    'w_res4_conv1': (2, 51, 84, 256, 1024, 1, 0),
    'w_res3_3x3': (2, 101, 167, 128, 128, 3, 1),
    'w_res3_conv3': (2, 101, 167, 512, 128, 1, 0),
    'w_res3_conv1': (2, 101, 167, 128, 512, 1, 0),
    'w_mask_head': (1024, 14, 14, 80, 256, 1, 0),
}


def run_wgrad(name, reps=8, iters=5):
    B, H, W, rows, cols, k, pad = WGRAD_SHAPES[name]
    M = B * H * W
    per_set = 4 * M * (rows + cols)
    n_sets = max(2, min(reps, -(-(300 << 20) // per_set)))
    sets = []
    for _ in range(n_sets):
        gy = torch.randn((B, H, W, rows), device='cuda')
        x = torch.randn((B, H, W, cols), device='cuda')
        E.round_tf32(gy, gy)
        E.round_tf32(x, x)
        sets.append((gy, x))
    gw = torch.zeros((rows, k * k * cols), device='cuda')

    def launch(i):
        gy, x = sets[i % n_sets]
        E.wgrad_tap(gy, x, gw, rows, cols, (H, W), k * k * cols, x_off=(-pad, -pad), taps=(k, k))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n_sets):
            launch(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / reps
        best = t if best is None else min(best, t)
    flops = 2.0 * M * rows * cols * k * k
    nbytes = per_set + 4 * rows * cols * k * k
    return {'shape': name, 'variant': 0, 'M': M, 'N': rows, 'K': cols * k * k, 'flags': 'wgrad',
            'dbg': None, 'us': best * 1e3, 'tflops': flops / best / 1e9, 'gbs': nbytes / best / 1e6}


def stream_rates():
    """DRAM rates of plain element-wise passes over res5-sized tensors (the read/write mixes of
    the conv epilogues), for comparison."""
    n = 50176 * 2048
    a = [torch.randn(n, device='cuda') for _ in range(3)]
    b = [torch.randn(n, device='cuda') for _ in range(3)]
    c = [torch.empty(n, device='cuda') for _ in range(3)]
    out = {}
    for name, fn, nb in (('copy_1r1w', lambda i: c[i].copy_(a[i]), 8 * n),
                         ('add_2r1w', lambda i: torch.add(a[i], b[i], out=c[i]), 12 * n),
                         ('fill_0r1w', lambda i: c[i].zero_(), 4 * n)):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(9):
            fn(i % 3)
        e1.record()
        torch.cuda.synchronize()
        out[name] = nb * 9 / e0.elapsed_time(e1) / 1e6
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--variants', default='0,1,2,3')
    ap.add_argument('--shapes', default=','.join(list(SHAPES) + list(WGRAD_SHAPES)))
    ap.add_argument('--reps', type=int, default=8)
    args = ap.parse_args()
    rows = []
    print('stream GB/s:', stream_rates(), flush=True)
    for name in args.shapes.split(','):
        if name in WGRAD_SHAPES:
            r = run_wgrad(name, reps=args.reps)
            rows.append(r)
            print('%-18s  %7.1f us %5.0f TF %5.0f GB/s' % (name, r['us'], r['tflops'], r['gbs']),
                  flush=True)
            continue
        line = '%-18s' % name
        for v in [int(t) for t in args.variants.split(',')]:
            r = run_shape(name, v, reps=args.reps)
            rows.append(r)
            line += '  v%d %7.1f us %5.0f TF %5.0f GB/s |' % (v, r['us'], r['tflops'], r['gbs'])
        print(line, flush=True)
        for r in rows[-len(args.variants.split(',')):]:
            if r['dbg']:
                print('    v%-2d ' % r['variant'] + ' '.join('%s %.0f' % (k, v / 1e3) for k, v in
                                                          r['dbg'].items()) + '  (kcycles)')
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
