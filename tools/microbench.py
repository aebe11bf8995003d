"""Kernel microbenchmarks (BASELINE config 5 and conv GEMM shapes) and the inference
throughput of BASELINE config 3, CUDA-event timed.
Usage: python tools/microbench.py [roi|nms|conv|peaks|infer|all] [--out FILE]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from chainer_mask_rcnn_b200 import _lib, functions, utils  # noqa: E402

PEAKS = {'hbm_gbs': 6543.7, 'bf16_tflops': 1606.6}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))))
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    _flush.zero_()


def time_ms(fn, iters=20, warmup=5, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def bench_roi(out):
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((1, 1024, 50, 68)).astype(np.float32)).cuda()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    for R in (300, 1000, 2000, 6000):
        rois = torch.from_numpy(synth.rois_xy(rs, R, 1, 800, 1088)).cuda()
        for (oh, ratio) in ((14, 0), (7, 0), (14, 2)):
            y = torch.empty((R, 1024, oh, oh), device='cuda')
            gx = torch.empty_like(x)
            nbytes = 4 * (R * 1024 * oh * oh + x.numel() + 5 * R)

            def fwd():
                _lib.call('cmr_roi_align_fwd', _lib.ptr(x), 1, 1024, 50, 68, _lib.ptr(rois), R,
                          oh, oh, 1. / 16, ratio, _lib.ptr(y), _lib.stream_ptr())

            def bwd():
                _lib.call('cmr_roi_align_bwd', _lib.ptr(y), _lib.ptr(rois), R, 1, 1024, 50, 68,
                          oh, oh, 1. / 16, ratio, _lib.ptr(gx), _lib.stream_ptr())

            yn = torch.empty((R, oh, oh, 1024), device='cuda')

            def fwd_nhwc():
                _lib.call('cmr_roi_align_nhwc_fwd', _lib.ptr(x_nhwc), 1, 50, 68, 1024,
                          _lib.ptr(rois), R, oh, oh, 1, 1. / 16, ratio, 0, _lib.ptr(yn),
                          _lib.stream_ptr())

            def bwd_nhwc():
                _lib.call('cmr_roi_align_nhwc_bwd', _lib.ptr(yn), _lib.ptr(rois), R, 1, 50, 68,
                          1024, oh, oh, 1, 1. / 16, ratio, _lib.ptr(gx), _lib.stream_ptr())

            # the drop-in operator (functions.roi_align_2d, NCHW in and out) incl. its layout
            # conversions; the backward is timed through autograd on a retained graph
            xg = x.detach().clone().requires_grad_(True)
            yg = functions.roi_align_2d(xg, rois, oh, oh, 1. / 16, ratio)
            gyg = torch.ones_like(yg)

            def fwd_api():
                with torch.no_grad():
                    functions.roi_align_2d(x, rois, oh, oh, 1. / 16, ratio)

            def bwd_api():
                xg.grad = None
                yg.backward(gyg, retain_graph=True)

            for name, fn in (('roi_align_fwd_nchw', fwd), ('roi_align_bwd_nchw', bwd),
                             ('roi_align_fwd_nhwc', fwd_nhwc), ('roi_align_bwd_nhwc', bwd_nhwc),
                             ('roi_align_2d_api_fwd', fwd_api), ('roi_align_2d_api_bwd', bwd_api)):
                if R == 6000 and 'bwd' in name and oh == 14 and ratio == 2:
                    continue
                med, best = time_ms(fn, iters=10, warmup=3)
                rec = dict(kernel=name, R=R, out=oh, sampling_ratio=ratio, ms=med, ms_min=best,
                           algo_bytes=nbytes, gbs=nbytes / med / 1e6,
                           frac_hbm=nbytes / med / 1e6 / PEAKS['hbm_gbs'])
                print(json.dumps(rec)); out.append(rec)


def bench_nms(out):
    rs = np.random.RandomState(1)
    for n in (300, 1000, 2000, 6000, 12000):
        boxes = torch.from_numpy(synth.clustered_boxes(rs, n, 800, 1088, max(2, n // 12))).cuda()
        keep = torch.empty((n,), dtype=torch.int32, device='cuda')
        nk = torch.zeros((1,), dtype=torch.int32, device='cuda')
        wsb = _lib.load().cmr_nms_workspace_bytes(n)
        ws = torch.empty((wsb // 8,), dtype=torch.int64, device='cuda')
        for thresh, limit in ((0.7, 0), (0.7, 2000), (0.5, 0)):
            def fn():
                _lib.call('cmr_nms', _lib.ptr(boxes), n, thresh, limit, _lib.ptr(keep),
                          _lib.ptr(nk), _lib.ptr(ws), wsb, _lib.stream_ptr())
            med, best = time_ms(fn, iters=10, warmup=3)
            k = int(nk.item())
            nbytes = 16 * n + 8 * n * ((n + 63) // 64) + 4 * k
            rec = dict(kernel='nms', n=n, thresh=thresh, limit=limit, kept=k, ms=med, ms_min=best,
                       algo_bytes=nbytes, gbs=nbytes / med / 1e6,
                       iou_pairs_per_s=n * (n - 1) / 2 / (med * 1e-3))
            print(json.dumps(rec)); out.append(rec)
    # full proposal pipeline, COCO train size, batch 2
    from oracle_free_anchor import anchors  # noqa
    anchor = torch.from_numpy(anchors(51, 84, (2, 4, 8, 16, 32))).cuda()
    n_anchor = anchor.shape[0]
    loc = torch.from_numpy(np.stack([synth.rpn_outputs(rs, n_anchor)[0] for _ in range(2)])).cuda()
    score = torch.from_numpy(np.stack([synth.rpn_outputs(rs, n_anchor)[1] for _ in range(2)])).cuda()
    pc = utils.ProposalCreator(min_size=0, n_test_pre_nms=6000, n_test_post_nms=1000)
    for train in (True, False):
        def fn():
            pc.batch(loc, score, anchor, (800, 1333), 1.6, train=train)
        med, best = time_ms(fn, iters=10, warmup=3)
        rec = dict(kernel='proposals_batch2', train=train, n_anchor=n_anchor, ms=med, ms_min=best)
        print(json.dumps(rec)); out.append(rec)


def bench_conv(out):
    shapes = [
        # name, B, H, W, C, N, k, s, p
        ('res5.conv2 3x3', 1024, 7, 7, 512, 512, 3, 1, 1),
        ('res5.conv3 1x1', 1024, 7, 7, 512, 2048, 1, 1, 0),
        ('res5.b.conv1 1x1', 1024, 7, 7, 2048, 512, 1, 1, 0),
        ('rpn.conv1 3x3', 2, 51, 84, 1024, 1024, 3, 1, 1),
        ('res4.conv2 3x3', 2, 51, 84, 256, 256, 3, 1, 1),
        ('res4.conv3 1x1', 2, 51, 84, 256, 1024, 1, 1, 0),
        ('res3.conv2 3x3', 2, 101, 167, 128, 128, 3, 1, 1),
        ('res2.conv1 1x1', 2, 201, 334, 64, 64, 1, 1, 0),
    ]
    for name, B, H, W, C, N, k, s, p in shapes:
        x = torch.randn((B, H, W, C), device='cuda')
        w = torch.randn((N, k, k, C), device='cuda') / (k * C ** 0.5)
        oh = (H + 2 * p - k) // s + 1
        ow = (W + 2 * p - k) // s + 1
        d = torch.empty((B, oh, ow, N), device='cuda')
        scale = torch.ones((N,), device='cuda')
        bias = torch.zeros((N,), device='cuda')
        flops = 2.0 * B * oh * ow * N * k * k * C
        for tile in (64, 128, 256):
            if tile > 64 and N <= tile // 2:
                continue
            desc = _lib.ConvDesc(B, H, W, C, C, oh, ow, k, k, s, p, N, oh, ow, N, 1, 0, 0, 1, 1, tile)

            def fn():
                _lib.call('cmr_conv_gemm_tc', ctypes.byref(desc), _lib.ptr(x), _lib.ptr(w),
                          _lib.ptr(d), _lib.ptr(scale), _lib.ptr(bias), None, None,
                          _lib.stream_ptr())
            med, best = time_ms(fn, iters=10, warmup=3)
            rec = dict(kernel='conv_gemm_tc', layer=name, M=B * oh * ow, N=N, K=k * k * C, tile_n=tile,
                       ms=med, ms_min=best, tflops=flops / med / 1e9,
                       frac_bf16_peak=flops / med / 1e9 / PEAKS['bf16_tflops'])
            print(json.dumps(rec)); out.append(rec)
        # cuDNN/cuBLAS TF32 library baseline for context
        torch.backends.cudnn.allow_tf32 = True
        xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        med, best = time_ms(lambda: torch.nn.functional.conv2d(xc, wc, stride=s, padding=p), iters=10, warmup=3)
        rec = dict(kernel='cudnn_tf32_reference', layer=name, ms=med, ms_min=best, tflops=flops / med / 1e9)
        print(json.dumps(rec)); out.append(rec)


def bench_peaks(out):
    """cuBLAS TF32 / bf16 GEMM throughput on this box (roofline denominators)."""
    n = 8192
    a = torch.randn((n, n), device='cuda')
    b = torch.randn((n, n), device='cuda')
    torch.backends.cuda.matmul.allow_tf32 = True
    med, best = time_ms(lambda: torch.matmul(a, b), iters=10, warmup=3, flush=False)
    rec = dict(kernel='cublas_tf32_8192', ms=med, ms_min=best, tflops=2.0 * n ** 3 / best / 1e9,
               tflops_median=2.0 * n ** 3 / med / 1e9)
    print(json.dumps(rec)); out.append(rec)
    # sustained: back to back for ~3 s
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    iters = max(10, int(3000 / best))
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rec = dict(kernel='cublas_tf32_8192_sustained', ms=ms, tflops=2.0 * n ** 3 / ms / 1e9)
    print(json.dumps(rec)); out.append(rec)
    torch.backends.cuda.matmul.allow_tf32 = False
    ah, bh = a.bfloat16(), b.bfloat16()
    med, best = time_ms(lambda: torch.matmul(ah, bh), iters=10, warmup=3, flush=False)
    rec = dict(kernel='cublas_bf16_8192', ms=med, ms_min=best, tflops=2.0 * n ** 3 / best / 1e9)
    print(json.dumps(rec)); out.append(rec)


def bench_hbm(out):
    """HBM stream rates on this box, to put the write-dominated kernels (ROIAlign forward,
    the residual epilogues) in context: write-only (fill), read-only (sum) and copy."""
    n = 1 << 29                                   # 2 GiB of float32
    a = torch.empty((n,), device='cuda')
    b = torch.empty((n,), device='cuda')
    a.normal_()
    for name, fn, nbytes in (('hbm_write_only_fill', lambda: b.fill_(1.0), 4 * n),
                             ('hbm_read_only_sum', lambda: a.sum(), 4 * n),
                             ('hbm_copy', lambda: b.copy_(a), 8 * n)):
        med, best = time_ms(fn, iters=10, warmup=3, flush=False)
        rec = dict(kernel=name, ms=med, ms_min=best, gbs=nbytes / best / 1e6,
                   gbs_median=nbytes / med / 1e6)
        print(json.dumps(rec)); out.append(rec)


def bench_infer(out):
    """BASELINE config 3: R50-C4 inference on a 1333x800 image, 6000 -> 1000 proposals ->
    100 detections with masks (MaskRCNN.predict: both head passes, per-class NMS, mask
    paste), images/s at batch 1 and 2.  Weights are random (flat class probabilities), so
    score_thresh is lowered until ~100 detections survive, like a trained model's output."""
    from chainer_mask_rcnn_b200 import models
    rs = np.random.RandomState(0)
    model = models.MaskRCNNResNet(50, 80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333)
    model.score_thresh = 1. / 81. * 1.02
    for bs in (1, 2):
        imgs = [rs.uniform(0, 255, (3, 800, 1333)).astype(np.float32) for _ in range(bs)]
        for _ in range(3):       # warm-up: also fills the page-locked download buffers
            bboxes, masks, labels, scores = model.predict(imgs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = 10
        for _ in range(iters):
            bboxes, masks, labels, scores = model.predict(imgs)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / iters
        # where the end-to-end time goes: each stage of predict bracketed by a device
        # synchronise (so the sum is a little above the pipelined e2e time)
        stages = {}

        def timed(obj, name):
            fn = getattr(obj, name)

            def wrapper(*a, **k):
                torch.cuda.synchronize()
                t = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                stages[name] = stages.get(name, 0.) + (time.perf_counter() - t) * 1e3 / 3
                return r
            setattr(obj, name, wrapper)
        names = ('_prepare_device', '_forward_padded', '_detect', '_cut', '_to_roi_masks',
                 '_to_masks')
        for n in names:
            timed(model, n)
        for _ in range(3):
            model.predict(imgs)
        for n in names:
            delattr(model, n)
        # device part only (network + detections), without the host mask download
        from chainer_mask_rcnn_b200.utils import config
        x = torch.from_numpy(np.stack(model.prepare(imgs)[0])).cuda()
        sizes = [(800, 1333)] * bs
        scales = np.ones(bs, np.float32)

        def dev_only():
            with config.using_config('train', False):
                feat, rois, cnt, cl, sc, _ = model._forward_padded(x, scales, False)
                model._detect(cl, sc, rois, cnt, sizes, scales)
        med, best = time_ms(dev_only, iters=10, warmup=2, flush=False)
        rec = dict(kernel='predict_r50_c4', batch=bs, n_det=[len(b) for b in bboxes],
                   ms_per_batch_e2e=ms, images_per_s_e2e=bs / ms * 1e3,
                   ms_per_batch_box_pass=med, images_per_s_box_pass=bs / med * 1e3,
                   stages_ms={k: round(v, 3) for k, v in stages.items()})
        print(json.dumps(rec)); out.append(rec)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='?', default='all')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    out = []
    if args.what in ('roi', 'all'):
        bench_roi(out)
    if args.what in ('nms', 'all'):
        bench_nms(out)
    if args.what in ('conv', 'all'):
        bench_conv(out)
    if args.what in ('peaks', 'all'):
        bench_peaks(out)
    if args.what in ('hbm', 'all'):
        bench_hbm(out)
    if args.what in ('infer', 'all'):
        bench_infer(out)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, 'w') as f:
            for r in out:
                f.write(json.dumps(r) + '\n')
