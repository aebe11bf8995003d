#!/usr/bin/env python
"""bench.py -- images/sec of the Mask R-CNN R50-C4 train step on B200.

  python bench.py --gpus N --steps K --warmup W            (this framework)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU path)
  python bench.py --config {train,r101,infer,roi_nms} ...  (BASELINE.json configs[1] (default),
                                                            [3], [2], [4]; one JSON line each)

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


def measure_tf32_peak(seconds=1.0):
    """cuBLAS TF32 8192^3 on THIS box, burst (best single call) and sustained (back to back
    for `seconds`): a measured stand-in for the TF32 tensor peak, which MEASURED_PEAKS.json
    does not hold (a library GEMM used as a yardstick only -- nothing on the hot path)."""
    import torch
    n = 8192
    a = torch.randn((n, n), device='cuda')
    b = torch.randn((n, n), device='cuda')
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            torch.matmul(a, b)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        iters = max(5, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / iters
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    fl = 2.0 * n ** 3
    return {'cublas_tf32_burst': fl / best / 1e9, 'cublas_tf32_sustained': fl / sus / 1e9}


def ncu_traffic(pattern, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, read at run time from the
    newest committed `ncu --set full` raw page under profiles/ matching `pattern` (CSV of
    `ncu -i ... --page raw --csv`).  -> (bytes or None, file name or None)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', pattern)))
    for path in reversed(files):
        try:
            with open(path, newline='') as f:
                rows = list(csv.reader(f))
            hdr = rows[0]
            units = rows[1]
            ik = hdr.index('Kernel Name')
            ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
            mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            for row in rows[2:]:
                if kernel_substr in row[ik]:
                    rd = float(row[ir].replace(',', '')) * mult.get(units[ir], 1.0)
                    wr = float(row[iw].replace(',', '')) * mult.get(units[iw], 1.0)
                    return rd + wr, os.path.basename(path)
        except Exception:
            continue
    return None, None


def ncu_l2_delivery(pattern, kernel_substr):
    """L2 -> SM bytes per L2 cycle of one launch (l1tex__m_xbar2l1tex_read_bytes.sum /
    lts__cycles_active.avg) from the newest committed ncu raw page matching `pattern`, next to
    the chip-wide ceiling /opt/skills/guides/B300_MICROARCH.md measures ("LTS throughput cap
    ~6300 B/cycle, path-independent").  None when no page is there."""
    import csv
    import glob
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for path in reversed(sorted(glob.glob(os.path.join(ROOT, 'profiles', pattern)))):
        try:
            with open(path, newline='') as f:
                rows = list(csv.reader(f))
            hdr, units = rows[0], rows[1]
            ik = hdr.index('Kernel Name')
            ib = hdr.index('l1tex__m_xbar2l1tex_read_bytes.sum')
            ic = hdr.index('lts__cycles_active.avg')
            it = hdr.index('gpu__time_duration.sum')
            for row in rows[2:]:
                if kernel_substr in row[ik]:
                    nbytes = float(row[ib].replace(',', '')) * unit.get(units[ib], 1.0)
                    cycles = float(row[ic].replace(',', ''))
                    us = float(row[it].replace(',', ''))
                    return {'bytes_per_l2_cycle': nbytes / cycles, 'cap_bytes_per_l2_cycle': 6300.0,
                            'frac': nbytes / cycles / 6300.0, 'tb_per_s': nbytes / us / 1e6,
                            'source': os.path.basename(path)}
        except Exception:
            continue
    return None


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
        extra = s.per_image_host_stages()
    return {'value': s.images_per_second(t, extra), 'unit': 'images/s', 'cores': pool.threads,
            'kind': 'port', 'sample': s.describe(t, extra)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port; chainer cannot
    be installed offline) on the host cores, rank 0 only."""
    if rank != 0:
        return
    if args.config == 'roi_nms':
        return run_reference_roi_nms(args)
    if args.config == 'infer':
        return run_reference_infer(args)
    from oracle import cpu_step
    total = max(args.steps + args.warmup, 1)
    budget = min(15.0, 150.0 / total)
    with all_host_threads() as pool:
        f = cpu_step.calibrate_fraction(budget)
        s = cpu_step.CpuStepSample(f, n_layers=args.layers)
        extra = s.per_image_host_stages()
        for _ in range(args.warmup):
            s.step()
        ts = [s.step() for _ in range(args.steps)]
    sec = float(np.mean(ts))
    v = s.images_per_second(sec, extra)
    sample = s.describe(sec, extra)
    line = {
        'impl': 'reference', 'metric': METRIC if args.layers == 50 else METRIC_R101,
        'value': v, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': BS / v * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus, args.layers),
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': pool.threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


METRIC_R101 = 'images/sec Mask R-CNN R101-C4 train step (3x800x1333)'
METRIC_INFER = 'images/sec Mask R-CNN R50-C4 inference (1333x800, 6000->1000 proposals -> 100 detections)'
METRIC_ROI = 'GB/s ROIAlign forward, 1000 proposals on a 1024x50x68 map, 14x14 bins (algorithmic bytes)'


def workload_config(n_gpus, layers=50):
    return {'workload': 'R%d-C4 COCO train step, bs=2 per GPU, 3x800x1333 synthetic images + ' % layers +
                        '40 instances/image, 12000->2000 proposals, 512 sampled RoIs/image, '
                        'roi_size 14 (BASELINE.json configs[%d])' % (1 if layers == 50 else 3),
            'global_batch': BS * n_gpus, 'parallelism': 'dp%d' % n_gpus,
            'l2': 'per-step working set (>4 GB of activations) exceeds the 126 MB L2; no flush'}


# ------------------------------------------------------------------ shared --
class Timer(object):
    """K calls of fn bracketed by a barrier + device synchronise on both sides, timed with
    CUDA events on the current stream, max over ranks."""

    def __init__(self, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def __call__(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], dtype=torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])


def collect_prof(lib, kinds):
    import ctypes
    out = {}
    for kind, name in kinds:
        ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.cmr_prof_collect(kind, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt))
        out[name] = (ms.value, work.value, cnt.value)
    return out


def tensor_roofline(prof, pk, tf32, steps, ms_total, measured_in):
    """roofline object of the dominant kernel (conv_gemm_tc_kernel) + the secondary views."""
    ms_k, work_k, cnt_k = prof['conv_gemm_tc']
    sustained_half = pk['bf16_tflops_sustained'] / 2.0
    burst_half = pk['bf16_tflops'] / 2.0
    achieved = work_k / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
    traffic, traffic_src = ncu_traffic('r*_ncu_conv_*raw.csv', 'conv_gemm_tc_kernel')
    rl = {
        'bound': 'tensor', 'kernel': 'conv_gemm_tc_kernel (fprop + dgrad implicit GEMM)',
        'achieved': achieved, 'peak': sustained_half, 'unit': 'TFLOP/s',
        'frac': achieved / sustained_half if sustained_half else None,
        'peak_source': '%s bf16_tflops_sustained / 2 (kind::tf32 issues at half the bf16 rate; '
                       'the kernel is timed inside a long step)' % pk['source'],
        # the same achieved rate against the other candidate denominators, all measured
        'frac_of_bf16_burst_half': achieved / burst_half if burst_half else None,
        'peak_bf16_burst_half': burst_half,
        'tf32_measured_this_box': tf32,
        'frac_of_cublas_tf32_sustained': (achieved / tf32['cublas_tf32_sustained']
                                          if tf32 and tf32.get('cublas_tf32_sustained') else None),
        # dram__bytes_read + write of ONE representative launch (res5 3x3 forward, 236.8 GFLOP,
        # CTA-pair kernel; algorithmic 105.5 MB x 2 operands + 102.8 MB output) parsed at run time
        # from the newest committed `ncu --set full` raw page; `achieved` sums all launches
        'traffic': traffic, 'traffic_source': traffic_src,
        'launches_per_step': cnt_k / steps,
        'share_of_step': ms_k / ms_total if ms_total else None,
        'measured_in': measured_in,
    }
    if 'conv_wgrad_tc' in prof:
        ms_w, work_w, cnt_w = prof['conv_wgrad_tc']
        rl['wgrad'] = {'achieved': work_w / (ms_w * 1e-3) / 1e12 if ms_w > 0 else 0.0,
                       'launches_per_step': cnt_w / steps,
                       'share_of_step': ms_w / ms_total if ms_total else None}
    for name in ('roi_align', 'roi_align_bwd'):
        if name not in prof:
            continue
        ms_r, work_r, cnt_r = prof[name]
        gbs = work_r / (ms_r * 1e-3) / 1e9 if ms_r > 0 else 0.0
        rl[name] = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                    'frac': gbs / pk['hbm_gbs'], 'launches_per_step': cnt_r / steps,
                    'us_per_launch': 1e3 * ms_r / cnt_r if cnt_r else None}
    ms_t, work_t, cnt_t = prof['conv_tensor_bound']
    ms_h, work_h, cnt_h = prof['conv_hbm_bound']
    tf_t = work_t / (ms_t * 1e-3) / 1e12 if ms_t > 0 else 0.0
    gb_h = work_h / (ms_h * 1e-3) / 1e9 if ms_h > 0 else 0.0
    rl['split'] = {
        'tensor_bound_launches': {'achieved': tf_t, 'peak': sustained_half, 'unit': 'TFLOP/s',
                                  'frac': tf_t / sustained_half if sustained_half else None,
                                  'frac_of_bf16_burst_half': tf_t / burst_half if burst_half else None,
                                  'launches_per_step': cnt_t / steps,
                                  'ms_per_step': ms_t / steps},
        'hbm_bound_launches': {'achieved': gb_h, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                               'frac': gb_h / pk['hbm_gbs'],
                               'launches_per_step': cnt_h / steps,
                               'ms_per_step': ms_h / steps,
                               # what binds them after the epilogue rewrite:
                               # operand delivery from L2, from the committed ncu page of the
                               # largest such launch (res5 conv3 + residual)
                               'l2_to_sm': ncu_l2_delivery('r*_ncu_conv1x1_raw.csv',
                                                           'conv_gemm_tc_kernel')},
    }
    return rl


PROF_KINDS = ((4, 'conv_tensor_bound'), (5, 'conv_hbm_bound'), (0, 'conv_gemm_tc'),
              (1, 'conv_wgrad_tc'), (2, 'roi_align'), (3, 'roi_align_bwd'))


# ------------------------------------------------------------------- train --
def run_train(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from chainer_mask_rcnn_b200 import _lib, datasets, models, optimizers
    lib = _lib.load()
    warmup = max(args.warmup, 3)
    timed = Timer(world)
    barrier = timed.barrier

    model = models.MaskRCNNResNet(args.layers, N_FG, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333, seed=0)
    chain = models.MaskRCNNTrainChain(model)
    # the reference's rule, examples/train_common.py:124-125: lr = 0.00125 x global batch
    # (CMR_BENCH_LR overrides it: a knob for checking one GPU at the 8-GPU learning rate)
    lr = float(os.environ.get('CMR_BENCH_LR', 0.00125 * BS * world))
    opt = optimizers.MomentumSGD(lr=lr, momentum=0.9)
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

    # sustained: the same resident step back to back for >= args.sustain seconds (power /
    # thermal steady state), clocks sampled over the whole stretch
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(np.ceil(args.sustain * 1e3 / ms_per_step)))
        del losses[:]
        clocks_s = ClockSampler(local) if rank == 0 else None
        ms_sus, _ = timed(step_resident, n_sus)
        clk_s = clocks_s.stop() if clocks_s else None
        sustained = {'value': BS * world / (ms_sus / n_sus * 1e-3), 'unit': 'images/s',
                     'steps': n_sus, 'seconds': ms_sus * 1e-3, 'ms_per_step': ms_sus / n_sus,
                     'clocks': clk_s}
        loss_sus = float(losses[-1].item())
        del losses[:-1]

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
    prof = collect_prof(lib, PROF_KINDS)

    # end to end: images start in pinned host memory every step, the loss is read back
    e2e = e2e_ref = None
    if not args.no_e2e:
        def e2e_leg(a_imgs, a_masks, n_steps):
            h2d, d2h = [0], [0]

            def step_e2e():
                # iteration i runs on the inputs prefetched during iteration i-1; the copies of
                # iteration i+1's inputs (issued right after the replay is enqueued, on a side
                # stream) overlap it.  Every timed step contains one full set of H2D copies
                # from host memory and one loss read-back.
                loss = updater.step()
                updater.prefetch(a_imgs, bboxes, labels, a_masks, scales)
                h2d[0] = updater.h2d_bytes              # images + instance masks + boxes/labels
                v = loss.item()
                d2h[0] = updater.d2h_bytes              # the loss
                return v

            updater.prefetch(a_imgs, bboxes, labels, a_masks, scales)
            for _ in range(3):                           # eager call, capture, first replay
                step_e2e()
            ms_e, wall_e = timed(step_e2e, n_steps)
            per = max(ms_e, wall_e) / n_steps
            return {'value': BS * world / (per * 1e-3), 'unit': 'images/s',
                    'h2d_bytes_per_step': int(h2d[0]), 'd2h_bytes_per_step': int(d2h[0]),
                    'ms_per_step': per}

        e2e = e2e_leg(imgs_pinned, masks_pinned, args.steps)
        e2e['inputs'] = ('images float32 + instance masks bit-packed (PackedMasks, packed by the '
                         'caller outside the timed region) in pinned host memory')
        # the reference's own batch format: datasets.concat_examples -> imgs (B,3,H,W) float32
        # and masks (B,40,800,1333) int32 NumPy arrays on the host (341 MB per step).  No host
        # pre-processing at all: the int32 block is uploaded as it is and read by the device
        # mask-target kernel.  Page-locked arrays (concat_examples(pinned=True)) upload
        # asynchronously under the previous iteration; pageable ones are staged by the driver.
        examples = [(imgs[i], bboxes[i], labels[i], masks[i], float(scales[i])) for i in range(BS)]
        e2e_ref = {}
        for form in ('pinned', 'pageable'):
            r_imgs, _, _, r_masks, _ = datasets.concat_examples(
                examples, padding=0, indices_concat=[0, 2, 3, 4], indices_to_device=[],
                pinned=form == 'pinned')
            assert r_masks.dtype == np.int32 and r_masks.shape == (BS, N_INST, H, W)
            e2e_ref[form] = e2e_leg(r_imgs, r_masks, max(3, args.steps // 2))
            updater._states.clear()                      # free the 341 MB input / staging buffers
            updater._prefetched = None
            del r_imgs, r_masks
            torch.cuda.empty_cache()
        e2e_ref['inputs'] = ('NumPy arrays exactly as the reference\'s concat_examples returns '
                             'them: imgs float32, masks (2,40,800,1333) int32; nothing packed or '
                             'converted on the host')

    tf32 = measure_tf32_peak() if rank == 0 else None
    if rank == 0:
        pk = peaks()
        line = {
            'metric': METRIC if args.layers == 50 else METRIC_R101,
            'value': value, 'unit': 'images/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32',
            'data': 'synthetic', 'config': workload_config(world, args.layers),
            'roofline': tensor_roofline(
                prof, pk, tf32, args.steps, ms_total,
                'second pass of the same %d steps launched eagerly with per-launch CUDA events '
                'on one stream (%.2f ms/step; the timed region replays a CUDA graph with the '
                'weight gradients on a side stream)' % (args.steps, ms_eager / args.steps)),
            'step_tflops': FLOPS_PER_IMAGE[args.layers] * BS * world / (ms_per_step * 1e-3) / 1e12,
            'gpu_launches': int(launches.item()),
            'clocks': clk,
            'loss_first': float(losses[0].item()), 'loss_last': float(losses[-1].item()),
            'loss_finite': bool(np.isfinite(float(losses[0].item())) and
                                np.isfinite(float(losses[-1].item()))),
            'host_wall_ms_per_step': wall_total / args.steps,
        }
        if sustained:
            sustained['loss_last'] = loss_sus
            line['sustained'] = sustained
        if e2e:
            line['e2e'] = e2e
            line['e2e_reference_format'] = e2e_ref
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(n_layers=args.layers)
        print(json.dumps(line))


# ------------------------------------------------------------------- infer --
def infer_inputs(rank, bs):
    rs = np.random.RandomState(100 + rank)
    return [rs.uniform(0, 255, (3, H, W)).astype(np.float32) for _ in range(bs)]


def run_infer(args, rank, world, local):
    """BASELINE.json configs[2]: MaskRCNN.predict on 1333x800 images -- 6000 -> 1000 proposals,
    box head on 1000 RoIs, per-class NMS -> 100 detections, mask head on those, mask paste.
    Random weights give flat class probabilities, so score_thresh is lowered until ~100
    detections per image survive, like a trained model's output."""
    import torch
    import torch.distributed as dist
    from chainer_mask_rcnn_b200 import _lib, models
    from chainer_mask_rcnn_b200.utils import config as cfg
    lib = _lib.load()
    warmup = max(args.warmup, 3)
    timed = Timer(world)
    bs = 1
    model = models.MaskRCNNResNet(50, N_FG, anchor_scales=(2, 4, 8, 16, 32), roi_size=14,
                                  min_size=800, max_size=1333, seed=0)
    model.score_thresh = 1. / 81. * 1.02
    imgs = infer_inputs(rank, bs)
    imgs_dev = [torch.from_numpy(a).cuda() for a in imgs]
    n_det = [0]

    def step_e2e():                                   # the public call, host arrays in and out
        bboxes, masks, labels, scores = model.predict(imgs)
        n_det[0] = int(np.mean([len(b) for b in bboxes]))
        step_e2e.out_bytes = int(sum(m.nbytes for m in masks) + sum(b.nbytes for b in bboxes) +
                                 sum(l.nbytes for l in labels) + sum(s.nbytes for s in scores))

    def step_resident():                              # raw images resident, masks left on device
        x, sizes, scales = model._prepare_device(imgs_dev)
        scales = np.asarray(scales, np.float64)
        with cfg.using_config('train', False), torch.no_grad():
            feat, rois, cnt, cl, sc, _ = model._forward_padded(x, scales, False)
            bboxes, labels, scores = model._cut(model._detect(cl, sc, rois, cnt, sizes, scales))
            idx = np.concatenate([np.full((len(b),), i, np.int32) for i, b in enumerate(bboxes)])
            model._to_roi_masks(feat, bboxes, idx, scales)

    for _ in range(warmup):
        step_e2e()
        step_resident()
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = lib.cmr_launch_count()
    ms_total, _ = timed(step_resident, args.steps)
    n1 = lib.cmr_launch_count()
    clk = clocks.stop() if clocks else None
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)
    lib.cmr_prof_enable(1)
    ms_prof, _ = timed(step_resident, args.steps)
    lib.cmr_prof_enable(0)
    prof = collect_prof(lib, PROF_KINDS)
    sustained = None
    if args.sustain > 0:
        n_sus = int(np.ceil(args.sustain * 1e3 / (ms_total / args.steps)))
        ms_sus, _ = timed(step_resident, n_sus)
        sustained = {'value': bs * world / (ms_sus / n_sus * 1e-3), 'unit': 'images/s',
                     'steps': n_sus, 'seconds': ms_sus * 1e-3}
    launches = torch.tensor([n1 - n0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(launches)
    tf32 = measure_tf32_peak() if rank == 0 else None
    if rank == 0:
        pk = peaks()
        per = ms_total / args.steps
        per_e = max(ms_e2e, wall_e2e) / args.steps
        line = {
            'metric': METRIC_INFER, 'value': bs * world / (per * 1e-3), 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': warmup, 'ms_per_step': per,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32',
            'data': 'synthetic',
            'config': {'workload': 'R50-C4 inference, batch 1 per GPU, 3x800x1333 raw image, '
                                   '6000->1000 proposals, box head on 1000 RoIs, 80-class NMS '
                                   '-> 100 detections, mask head + paste (BASELINE.json '
                                   'configs[2])', 'detections_per_image': n_det[0],
                       'global_batch': bs * world, 'parallelism': 'replicas x%d' % world,
                       'l2': 'working set (~1.5 GB of activations per image) exceeds the L2; '
                             'no flush'},
            'roofline': tensor_roofline(prof, pk, tf32, args.steps, ms_prof,
                                        'a third pass of the same %d steps with per-launch CUDA '
                                        'events (forward GEMMs only)' % args.steps),
            'e2e': {'value': bs * world / (per_e * 1e-3), 'unit': 'images/s',
                    'h2d_bytes_per_step': int(sum(a.nbytes for a in imgs)),
                    'd2h_bytes_per_step': step_e2e.out_bytes, 'ms_per_step': per_e,
                    'note': 'MaskRCNN.predict(imgs): raw float32 images from host memory in, '
                            'boxes / labels / scores / full-resolution boolean masks (the '
                            'reference\'s output format, ~1 MB per detection) back on the host'},
            'gpu_launches': int(launches.item()), 'clocks': clk,
        }
        if sustained:
            line['sustained'] = sustained
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_infer()
        print(json.dumps(line))


def cpu_baseline_infer(budget_s=10.0):
    from oracle import cpu_step
    with all_host_threads() as pool:
        f = cpu_step.calibrate_fraction(budget_s * 3)       # forward only: ~1/3 of a train sample
        s = cpu_step.CpuStepSample(f, n_layers=50)
        s.forward_only()
        t = s.forward_only()
    # per image: the backbone + RPN on the whole image and the box head on 1000 RoIs
    # (= 1000/512 of the train sample's RoIs per image); the mask pass (100 RoIs) is left out
    v = s.fraction / t
    return {'value': v, 'unit': 'images/s', 'cores': pool.threads, 'kind': 'port',
            'sample': '%.4f of one image, forward only (%dx%d crop, %d RoIs ~ the 1000-RoI box '
                      'pass scaled), %.1f s, NumPy + BLAS restatement (oracle/model.py), linearly '
                      'extrapolated' % (s.fraction, s.h, s.w, s.n_roi_infer, t)}


def run_reference_infer(args):
    cb = cpu_baseline_infer(min(10.0, 100.0 / max(args.steps + args.warmup, 1)))
    v = cb['value']
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC_INFER, 'value': v, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / v,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': 'R50-C4 inference 1333x800 (configs[2])'},
        'cpu_baseline': cb, 'gpu_launches': 0,
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


# ----------------------------------------------------------------- roi_nms --
def roi_nms_inputs(seed=0):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import synth
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((1, 1024, 50, 68)).astype(np.float32)
    rois = {R: synth.rois_xy(rs, R, 1, 800, 1088) for R in (300, 1000, 2000, 6000)}
    boxes = {n: synth.clustered_boxes(rs, n, 800, 1088, max(2, n // 12))
             for n in (300, 1000, 2000, 6000, 12000)}
    return x, rois, boxes


def run_roi_nms(args, rank, world, local):
    """BASELINE.json configs[4]: the drop-in ROIAlign operator (functions.roi_align_2d: NCHW
    map in, (R,C,14,14) out, forward and backward through autograd) and NMS at 300..6000
    (12000) proposals on a 1024x50x68 map, GB/s of ALGORITHMIC bytes
    (4*(R*C*oh*ow + N*C*H*W + 5R) each way; NMS 16n + 8n*ceil(n/64) + 4k).  L2 is flushed
    (a 256 MB buffer is overwritten) before every timed launch; kernel time = CUDA events
    recorded inside the library around the kernel + zero fill (cmr_prof)."""
    import torch
    from chainer_mask_rcnn_b200 import _lib, functions
    lib = _lib.load()
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 5)
    x_np, rois_np, boxes_np = roi_nms_inputs(rank)
    x = torch.from_numpy(x_np).cuda()
    x_cl = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)      # channels-last map
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    pk = peaks()
    import ctypes

    def kernel_ms(kind):
        ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.cmr_prof_collect(kind, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt))
        return ms.value / max(cnt.value, 1), work.value / max(cnt.value, 1), cnt.value

    def time_call(fn, n):
        """median wall-to-wall (CUDA events around the whole Python call)"""
        ts = []
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    cases = []
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = lib.cmr_launch_count()
    for R in (300, 1000, 2000, 6000):
        rois = torch.from_numpy(rois_np[R]).cuda()
        for xin, layout in ((x, 'nchw'), (x_cl, 'channels_last')):
            xg = xin.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
            y = functions.roi_align_2d(xg, rois, 14, 14, 1. / 16)
            gy = torch.ones_like(y)

            def fwd():
                with torch.no_grad():
                    functions.roi_align_2d(xin, rois, 14, 14, 1. / 16)

            def bwd():
                xg.grad = None
                y.backward(gy, retain_graph=True)

            for _ in range(warmup):
                fwd(); bwd()
            torch.cuda.synchronize()
            lib.cmr_prof_enable(1)
            api_f = time_call(fwd, steps)
            k_f, bytes_f, _ = kernel_ms(6)
            api_b = time_call(bwd, steps)
            k_b, bytes_b, _ = kernel_ms(7)
            lib.cmr_prof_enable(0)
            nbytes = 4.0 * (R * 1024 * 196 + x.numel() + 5 * R)
            cases.append({'op': 'roi_align_2d', 'R': R, 'map': layout, 'algo_bytes': nbytes,
                          'fwd_kernel_ms': k_f, 'fwd_kernel_gbs': nbytes / k_f / 1e6 if k_f else None,
                          'fwd_api_ms': api_f, 'fwd_api_gbs': nbytes / api_f / 1e6,
                          'bwd_kernel_ms': k_b, 'bwd_kernel_gbs': nbytes / k_b / 1e6 if k_b else None,
                          'bwd_api_ms': api_b, 'bwd_api_gbs': nbytes / api_b / 1e6})
            del xg, y, gy
    for n in (300, 1000, 2000, 6000, 12000):
        boxes = torch.from_numpy(boxes_np[n]).cuda()
        keep = torch.empty((n,), dtype=torch.int32, device='cuda')
        nk = torch.zeros((1,), dtype=torch.int32, device='cuda')
        wsb = lib.cmr_nms_workspace_bytes(n)
        ws = torch.empty((wsb // 8,), dtype=torch.int64, device='cuda')

        def nms():
            _lib.call('cmr_nms', _lib.ptr(boxes), n, 0.7, 0, _lib.ptr(keep), _lib.ptr(nk),
                      _lib.ptr(ws), wsb, _lib.stream_ptr())
        for _ in range(warmup):
            nms()
        ms = time_call(nms, steps)
        k = int(nk.item())
        nb = 16.0 * n + 8.0 * n * ((n + 63) // 64) + 4.0 * k
        cases.append({'op': 'nms', 'n': n, 'thresh': 0.7, 'kept': k, 'ms': ms, 'algo_bytes': nb,
                      'gbs': nb / ms / 1e6, 'iou_pairs_per_s': n * (n - 1) / 2 / (ms * 1e-3),
                      'bound': 'latency / ALU (n^2/2 IoUs), not HBM'})
    n1 = lib.cmr_launch_count()
    clk = clocks.stop() if clocks else None
    if rank != 0:
        return
    head = [c for c in cases if c['op'] == 'roi_align_2d' and c['R'] == 1000 and c['map'] == 'nchw'][0]
    traffic, traffic_src = ncu_traffic('r*_ncu_roi_cl_fwd*_raw.csv', 'roi_align_cl_fwd')
    worst = min(min(c['fwd_api_gbs'], c['bwd_api_gbs']) for c in cases if c['op'] == 'roi_align_2d')
    line = {
        'metric': METRIC_ROI, 'value': head['fwd_kernel_gbs'], 'unit': 'GB/s', 'n_gpus': 1,
        'steps': steps, 'warmup': warmup, 'ms_per_step': head['fwd_kernel_ms'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'ROIAlign + NMS microbench: functions.roi_align_2d forward / '
                               'backward at R = 300..6000 on a 1x1024x50x68 map, 14x14 bins, '
                               'sampling_ratio 0, spatial_scale 1/16; NMS at n = 300..12000, '
                               'thresh 0.7 (BASELINE.json configs[4])',
                   'l2': 'flushed before every timed launch (256 MB overwrite)',
                   'parallelism': 'single GPU'},
        'roofline': {'bound': 'hbm', 'kernel': 'roi_align_cl_fwd_kernel (R = 1000)',
                     'achieved': head['fwd_kernel_gbs'], 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                     'frac': head['fwd_kernel_gbs'] / pk['hbm_gbs'],
                     'peak_source': '%s hbm_gbs (copy bandwidth)' % pk['source'],
                     'traffic': traffic, 'traffic_source': traffic_src,
                     'backward': {'kernel': 'cudaMemset + roi_align_cl_bwd_kernel (R = 1000)',
                                  'achieved': head['bwd_kernel_gbs'],
                                  'frac': head['bwd_kernel_gbs'] / pk['hbm_gbs']},
                     'through_api_worst_case_frac': worst / pk['hbm_gbs']},
        'e2e': {'value': head['fwd_api_gbs'], 'unit': 'GB/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0,
                'note': 'the same forward through functions.roi_align_2d on a plain NCHW CUDA '
                        'tensor (allocation of the output, re-layout of the map, kernel); the '
                        'operator works on device arrays, there is no host copy to time'},
        'cases': cases, 'gpu_launches': int(n1 - n0), 'clocks': clk,
    }
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_roi_nms()
    print(json.dumps(line))


def cpu_baseline_roi_nms(budget_s=10.0):
    """ROIAlign: the oracle's restatement of ROIAlign2D.forward_cpu (pinned to the reference by
    golden vectors; the reference's own pure-Python loop costs ~13 us per output element, the
    restatement is vectorised per RoI, i.e. this favours the CPU side) on a slice of the same
    inputs, extrapolated to R = 1000 x 1024 channels; NMS: NumPy restatement at full size."""
    from oracle import bbox as ob
    from oracle import roi_align as ora
    x, rois, boxes = roi_nms_inputs(0)
    with all_host_threads() as pool:
        R_s, C_s = 500, 512          # half the RoIs x half the channels: a few seconds
        t0 = time.perf_counter()
        ora.roi_align_forward(x[:, :C_s], rois[1000][:R_s], 14, 14, 1. / 16, 0)
        t_roi = time.perf_counter() - t0
        t0 = time.perf_counter()
        ob.non_maximum_suppression(boxes[6000], 0.7)
        t_nms = time.perf_counter() - t0
    full = t_roi * (1000. / R_s) * (1024. / C_s)
    nbytes = 4.0 * (1000 * 1024 * 196 + x.size + 5 * 1000)
    return {'value': nbytes / full / 1e9, 'unit': 'GB/s', 'cores': 1, 'kind': 'port',
            'sample': 'ROIAlign forward on %d RoIs x %d channels of the R = 1000 case (%.2f s), '
                      'linearly extrapolated to 1000 x 1024 (%.0f s); NumPy NMS of 6000 boxes at '
                      'full size: %.2f s' % (R_s, C_s, t_roi, full, t_nms),
            'nms_6000_ms': t_nms * 1e3}


def run_reference_roi_nms(args):
    cb = cpu_baseline_roi_nms()
    v = cb['value']
    nbytes = 4.0 * (1000 * 1024 * 196 + 1024 * 50 * 68 + 5 * 1000)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC_ROI, 'value': v, 'unit': 'GB/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': nbytes / v / 1e6,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': 'ROIAlign + NMS microbench (configs[4])'},
        'cpu_baseline': cb, 'gpu_launches': 0,
        'e2e': {'value': v, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='train', choices=['train', 'r101', 'infer', 'roi_nms'],
                    help='train = BASELINE configs[1] (default), r101 = configs[3], infer = '
                         'configs[2], roi_nms = configs[4]')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--sustain', type=float, default=5.0,
                    help='seconds of back-to-back steps for the `sustained` object (0 = skip)')
    ap.add_argument('--layers', type=int, default=50, choices=[50, 101],
                    help='backbone depth of the train step: 101 is the same as --config r101')
    args = ap.parse_args()
    if args.config == 'r101':
        args.layers = 101
    elif args.layers == 101 and args.config == 'train':
        args.config = 'r101'
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
    if args.config in ('train', 'r101'):
        run_train(args, rank, world, local)
    elif args.config == 'infer':
        run_infer(args, rank, world, local)
    else:
        if rank == 0:
            run_roi_nms(args, rank, world, local)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
