#!/usr/bin/env python
"""bench.py -- images/sec of the Mask R-CNN R50-C4 train step on B200.

  python bench.py --gpus N --steps K --warmup W            (this framework)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU path)

One "step" = one full training iteration of BASELINE.json configs[1]: R50-C4, COCO
shapes (80 classes, 15 anchors), batch 2 per GPU, 3x800x1333 synthetic images with 40
instances each: forward, target creation (anchors, RoI sampling, mask rasterisation),
five losses, backward, gradient all-reduce (N > 1) and the MomentumSGD update, run
through the package's public per-iteration call (optimizers.GraphedUpdater: the step is
a CUDA-graph replay).  Prints ONE JSON line (rank 0).

  value     inputs (images, boxes, instance masks) already resident in HBM
  e2e       the same call with images / bit-packed masks in pinned host memory and boxes as NumPy
            arrays: H2D copies and the loss read-back are inside the timed region
  roofline  per-launch CUDA-event times of the tensor-core kernels, taken in a second
            pass of the same K steps run eagerly (events cannot be read inside a graph)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'images/sec Mask R-CNN R50-C4 train step (3x800x1333)'
H, W, BS, N_INST, N_FG = 800, 1333, 2, 40, 80
MEAN = (123.152, 115.903, 103.063)
FLOPS_PER_IMAGE = {50: 3.157e12, 101: 3.644e12}   # GEMM FLOPs of one train step / 2


def peaks():
    p = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
         'source': 'fallback'}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p.update(json.load(f))
        p['source'] = 'measured'
    except Exception:
        pass
    return p


def synth_batch(seed, bs=BS):
    """SURVEY.md 8d synthetic inputs: U[0,255) - mean images; 40 instances per image with
    log-uniform sizes, uniform labels and filled-ellipse masks."""
    rs = np.random.RandomState(seed)
    imgs = (rs.uniform(0, 255, (bs, 3, H, W)).astype(np.float32) -
            np.asarray(MEAN, np.float32)[None, :, None, None])
    bboxes, labels, masks = [], [], []
    for _ in range(bs):
        hh = np.exp(rs.uniform(np.log(24), np.log(480), N_INST))
        ww = np.exp(rs.uniform(np.log(24), np.log(480), N_INST))
        cy, cx = rs.uniform(0, H, N_INST), rs.uniform(0, W, N_INST)
        b = np.stack([np.clip(cy - hh / 2, 0, H), np.clip(cx - ww / 2, 0, W),
                      np.clip(cy + hh / 2, 0, H), np.clip(cx + ww / 2, 0, W)], 1).astype(np.float32)
        m = np.zeros((N_INST, H, W), np.int32)
        for i, (y1, x1, y2, x2) in enumerate(b):
            ys, xs = int(np.floor(y1)), int(np.floor(x1))
            ye, xe = int(np.ceil(y2)), int(np.ceil(x2))
            yy, xx = np.mgrid[ys:ye, xs:xe]
            ry, rx = max((y2 - y1) / 2, 1.), max((x2 - x1) / 2, 1.)
            m[i, ys:ye, xs:xe] = ((((yy + .5 - (y1 + y2) / 2) / ry) ** 2 +
                                   ((xx + .5 - (x1 + x2) / 2) / rx) ** 2) <= 1.)
        bboxes.append(b)
        labels.append(rs.randint(0, N_FG, N_INST).astype(np.int32))
        masks.append(m)
    return imgs, bboxes, labels, masks, np.full((bs,), 1.6, np.float32)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.path = tempfile.mktemp(prefix='clocks_', suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                    'sw_power_cap'), f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out['reasons'] = sorted(reasons)
        return out


class all_host_threads(object):
    """Gives NumPy's BLAS every host core for the CPU legs (torchrun exports
    OMP_NUM_THREADS=1 to its workers); `.threads` is what the BLAS pool then reports."""

    def __enter__(self):
        from threadpoolctl import threadpool_info, threadpool_limits
        self._limits = threadpool_limits(limits=os.cpu_count(), user_api='blas')
        blas = [d['num_threads'] for d in threadpool_info() if d.get('user_api') == 'blas']
        self.threads = max(blas) if blas else 1
        return self

    def __exit__(self, *exc):
        self._limits.restore_original_limits()
        return False


def cpu_baseline(budget_s=12.0, n_layers=50):
    """The oracle port of the reference CPU path on a bounded sample (oracle/cpu_step.py)."""
    from oracle import cpu_step
    with all_host_threads() as pool:
        f = cpu_step.calibrate_fraction(budget_s)
        s = cpu_step.CpuStepSample(f, n_layers=n_layers)
        s.step()                                   # warm the BLAS threads / page in buffers
        t = s.step()
    return {'value': s.images_per_second(t), 'unit': 'images/s', 'cores': pool.threads,
            'kind': 'port',
            'sample': '%.4f of one image: %dx%d crop through backbone+RPN and %d RoIs through '
                      'the res5 head, forward+backward, %.1f s; NumPy im2col + BLAS sgemm '
                      'restatement of the Chainer CPU path (oracle/model.py), linearly '
                      'extrapolated' % (s.fraction, s.h, s.w, s.n_roi, t)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port; chainer cannot
    be installed offline) on the host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle import cpu_step
    total = max(args.steps + args.warmup, 1)
    budget = min(15.0, 150.0 / total)
    with all_host_threads() as pool:
        f = cpu_step.calibrate_fraction(budget)
        s = cpu_step.CpuStepSample(f, n_layers=args.layers)
        for _ in range(args.warmup):
            s.step()
        ts = [s.step() for _ in range(args.steps)]
    sec = float(np.mean(ts))
    v = s.images_per_second(sec)
    sample = ('%.4f of one image per step (%dx%d crop, %d RoIs), forward+backward, linearly '
              'extrapolated to images/s' % (s.fraction, s.h, s.w, s.n_roi))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus, args.layers),
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': pool.threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus, layers=50):
    return {'workload': 'R%d-C4 COCO train step, bs=2 per GPU, 3x800x1333 synthetic images + ' % layers +
                        '40 instances/image, 12000->2000 proposals, 512 sampled RoIs/image, '
                        'roi_size 14 (BASELINE.json configs[%d])' % (1 if layers == 50 else 3),
            'global_batch': BS * n_gpus, 'parallelism': 'dp%d' % n_gpus,
            'l2': 'per-step working set (>4 GB of activations) exceeds the 126 MB L2; no flush'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--layers', type=int, default=50, choices=[50, 101],
                    help='backbone depth: 50 = BASELINE configs[1] (default), 101 = configs[3]')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if rank == 0:
        entry.build()            # a no-op when the in-tree library is current
    if world > 1:
        dist.barrier()           # nobody loads the library before rank 0 is done with it
    from chainer_mask_rcnn_b200 import _lib, models, optimizers
    lib = _lib.load()
    warmup = max(args.warmup, 3)

    model = models.MaskRCNNResNet(args.layers, N_FG, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333, seed=0)
    chain = models.MaskRCNNTrainChain(model)
    opt = optimizers.MomentumSGD(lr=0.00125 * BS * world, momentum=0.9)
    if world > 1:
        opt = optimizers.create_multi_node_optimizer(opt, optimizers.create_communicator())
    opt.setup(chain)
    opt.add_hook(optimizers.WeightDecay(1e-4))

    imgs, bboxes, labels, masks, scales = synth_batch(rank)
    np.random.seed(1000 + rank)
    imgs_pinned = torch.from_numpy(imgs).pin_memory()
    imgs_dev = imgs_pinned.cuda()
    # instance masks packed one bit per pixel (B,G,H,ceil(W/8)): the device-side mask-target
    # path reads the bits; 8x fewer bytes per step than uint8 masks
    masks_pinned = models.utils.PackedMasks.from_numpy(np.stack(masks), pin=True)
    masks_dev = masks_pinned.to('cuda')
    updater = optimizers.GraphedUpdater(opt, chain, max_boxes=64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    losses = []

    def step_resident():
        loss = updater(imgs_dev, bboxes, labels, masks_dev, scales)
        losses.append(loss.array.clone())

    for _ in range(warmup):      # call 1 runs eagerly, call 2 captures the graph, then replays
        step_resident()
    barrier()

    clocks = ClockSampler(local) if rank == 0 else None
    n0 = lib.cmr_launch_count()
    ms_total, wall_total = timed(step_resident, args.steps)
    n1 = lib.cmr_launch_count()
    clk = clocks.stop() if clocks else None
    n_launch = updater.launches_per_replay * args.steps + (n1 - n0)
    launches = torch.tensor([n_launch], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(launches)
    ms_per_step = ms_total / args.steps
    value = BS * world / (ms_per_step * 1e-3)

    # roofline pass: the same K steps launched eagerly, every tensor-core launch bracketed
    # by CUDA events on its stream (cmr_prof_enable)
    def step_eager():
        loss = opt.update(chain, imgs_dev, bboxes, labels, masks_dev, scales)
        losses.append(loss.array)

    # (weight gradients on the main stream for this pass: a kernel's own rate is what the
    # roofline is about, not the rate it gets while sharing the SMs with another stream)
    os.environ['CMR_GRAD_SIDE'] = '0'
    step_eager()
    lib.cmr_prof_enable(1)
    ms_eager, _ = timed(step_eager, args.steps)
    lib.cmr_prof_enable(0)
    os.environ.pop('CMR_GRAD_SIDE')
    import ctypes
    prof = {}
    # the two secondary views of the conv_gemm launches are collected before kind 0
    for kind, name in ((4, 'conv_tensor_bound'), (5, 'conv_hbm_bound'), (0, 'conv_gemm_tc'),
                       (1, 'conv_wgrad_tc'), (2, 'roi_align'), (3, 'roi_align_bwd')):
        ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.cmr_prof_collect(kind, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt))
        prof[name] = (ms.value, work.value, cnt.value)

    # end to end: images start in pinned host memory every step, the loss is read back
    e2e = None
    if not args.no_e2e:
        h2d, d2h = [0], [0]

        def step_e2e():
            # iteration i runs on the inputs prefetched during iteration i-1; the copies of
            # iteration i+1's inputs (issued right after the replay is enqueued, on a side
            # stream) overlap it.  Every timed step contains one full set of H2D copies
            # from pinned host memory and one loss read-back.
            loss = updater.step()
            updater.prefetch(imgs_pinned, bboxes, labels, masks_pinned, scales)
            h2d[0] = updater.h2d_bytes                  # images + instance masks + boxes/labels
            v = loss.item()
            d2h[0] = updater.d2h_bytes                  # the loss
            return v

        updater.prefetch(imgs_pinned, bboxes, labels, masks_pinned, scales)
        step_e2e()
        n_e2e = args.steps
        ms_e2e, wall_e2e = timed(step_e2e, n_e2e)
        per = max(ms_e2e, wall_e2e) / n_e2e
        e2e = {'value': BS * world / (per * 1e-3), 'unit': 'images/s',
               'h2d_bytes_per_step': int(h2d[0]), 'd2h_bytes_per_step': int(d2h[0]),
               'ms_per_step': per}

    if rank == 0:
        pk = peaks()
        ms_k, work_k, cnt_k = prof['conv_gemm_tc']
        tf32_peak = pk['bf16_tflops_sustained'] / 2.0
        achieved = work_k / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
        ms_w, work_w, cnt_w = prof['conv_wgrad_tc']
        roi = {}
        for name in ('roi_align', 'roi_align_bwd'):
            ms_r, work_r, cnt_r = prof[name]
            gbs = work_r / (ms_r * 1e-3) / 1e9 if ms_r > 0 else 0.0
            roi[name] = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                         'frac': gbs / pk['hbm_gbs'], 'launches_per_step': cnt_r / args.steps,
                         'us_per_launch': 1e3 * ms_r / cnt_r if cnt_r else None}
        # conv_gemm launches split by what bounds them (arithmetic intensity against the
        # machine balance, csrc/conv_tc.cu): long reductions against the tensor peak, short
        # reductions with residual / mask operands against the HBM peak
        ms_t, work_t, cnt_t = prof['conv_tensor_bound']
        ms_h, work_h, cnt_h = prof['conv_hbm_bound']
        tf_t = work_t / (ms_t * 1e-3) / 1e12 if ms_t > 0 else 0.0
        gb_h = work_h / (ms_h * 1e-3) / 1e9 if ms_h > 0 else 0.0
        split = {
            'tensor_bound_launches': {'achieved': tf_t, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                                      'frac': tf_t / tf32_peak if tf32_peak else None,
                                      'launches_per_step': cnt_t / args.steps,
                                      'ms_per_step': ms_t / args.steps},
            'hbm_bound_launches': {'achieved': gb_h, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                   'frac': gb_h / pk['hbm_gbs'],
                                   'launches_per_step': cnt_h / args.steps,
                                   'ms_per_step': ms_h / args.steps},
        }
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32',
            'data': 'synthetic', 'config': workload_config(world, args.layers),
            'roofline': {
                'bound': 'tensor', 'kernel': 'conv_gemm_tc_kernel (fprop + dgrad implicit GEMM)',
                'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                'frac': achieved / tf32_peak if tf32_peak else None,
                # dram__bytes_read + write of ONE representative launch (res5 3x3 forward,
                # 236.8 GFLOP, CTA-pair kernel) from the ncu --set full capture
                # profiles/r1_ncu_conv_v12_raw.csv; `achieved` sums all launches of the step
                'traffic': 184.4e6,
                'peak_source': '%s bf16_tflops_sustained / 2: kind::tf32 issues at half the '
                               'bf16 rate (cuBLAS TF32 8192^3 on this pool: 763 burst / 622 '
                               'sustained TFLOP/s)' % pk['source'],
                'launches_per_step': cnt_k / args.steps,
                'share_of_step': ms_k / ms_total if ms_total else None,
                'measured_in': 'second pass of the same %d steps launched eagerly with '
                               'per-launch CUDA events on one stream (%.2f ms/step; the timed '
                               'region replays a CUDA graph with the weight gradients on a '
                               'side stream)' % (args.steps, ms_eager / args.steps),
                'wgrad': {'achieved': work_w / (ms_w * 1e-3) / 1e12 if ms_w > 0 else 0.0,
                          'launches_per_step': cnt_w / args.steps,
                          'share_of_step': ms_w / ms_total if ms_total else None},
                # the HBM-bound kernel north_star names: ROIAlign launches of the same pass,
                # algorithmic bytes 4*(R*C*oh*ow + N*C*H*W + 5R) per launch (oh*ow = the 7x7
                # bins res5's stride-2 convolutions read)
                'roi_align': roi['roi_align'], 'roi_align_bwd': roi['roi_align_bwd'],
                'split': split,
            },
            'step_tflops': FLOPS_PER_IMAGE[args.layers] * BS * world / (ms_per_step * 1e-3) / 1e12,
            'gpu_launches': int(launches.item()),
            'clocks': clk,
            'loss_first': float(losses[0].item()), 'loss_last': float(losses[-1].item()),
            'host_wall_ms_per_step': wall_total / args.steps,
        }
        if e2e:
            line['e2e'] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(n_layers=args.layers)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
