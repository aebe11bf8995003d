"""Device plumbing shared by the model classes: thin wrappers over the C ABI
(include/cmr_b200.h) working on channels-last torch CUDA tensors, and the flat
parameter / gradient / momentum buffers.

Everything the network computes runs in libcmr_b200's kernels; torch is used for
device memory, streams and views only.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _lib

f32 = torch.float32


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv_out(size, k, s, p):
    return (size + 2 * p - k) // s + 1


def conv_gemm(x, w, n, kh=1, kw=1, stride=1, pad=0, out=None, scale=None, bias=None,
              addend=None, mask=None, relu=False, round_out=True, in_c=None, in_ld=None,
              out_hw=None, d_stride=1, d_off=(0, 0), tile_n=0, bcast=None, bcast_group=1,
              bcast_scale=0., tap_cols=0):
    """out[b, oy*d_stride+d_off[0], ox*d_stride+d_off[1], :n] =
    epilogue(sum_{fr,fs,c} x[b, oy*stride-pad+fr, ox*stride-pad+fs, c] * w[n, fr, fs, c]).
    x (B,H,W,C) contiguous NHWC; w (n, kh, kw, in_c) contiguous; see cmr_conv_gemm_tc."""
    B, H, W, C = x.shape
    in_c = C if in_c is None else in_c
    in_ld = C if in_ld is None else in_ld
    oh, ow = out_hw if out_hw else (conv_out(H, kh, stride, pad), conv_out(W, kw, stride, pad))
    if out is None:
        out = torch.empty((B, oh, ow, n), dtype=f32, device=x.device)
    _, dh, dw, dld = out.shape
    desc = _lib.ConvDesc(B, H, W, in_c, in_ld, oh, ow, kh, kw, stride, pad, n, dh, dw, dld,
                         d_stride, d_off[0], d_off[1], int(relu), int(round_out), tile_n,
                         tap_cols)
    ws = conv_workspace(x.device)
    _lib.call('cmr_conv_gemm_tc_ws', ctypes.byref(desc), _p(x), _p(w), _p(out), _p(scale),
              _p(bias), _p(addend), _p(mask), _p(bcast), int(bcast_group), float(bcast_scale),
              _p(ws), ws.numel() if ws is not None else 0, stream())
    return out


_conv_ws = {}
_ws_slot = [0]
# off by default: measured 1-4 % faster per launch alone, 0.5 % slower inside the train step
_split_tail = [os.environ.get('CMR_CONV_SPLIT_TAIL', '0') == '1']


class ws_slot(object):
    """Selects the K-split-tail workspace for the conv_gemm calls inside the block.  A
    workspace serves one launch at a time, so every chain of convolutions that may run next to
    another one (the RPN branch on its side stream) uses its own slot; the main chain is slot 0.
    (Slots rather than stream handles: a CUDA-graph capture runs on its own capture stream.)"""

    def __init__(self, slot):
        self.slot = slot

    def __enter__(self):
        self.old = _ws_slot[0]
        _ws_slot[0] = self.slot

    def __exit__(self, *exc):
        _ws_slot[0] = self.old


def conv_workspace(device):
    """The K-split-tail workspace of cmr_conv_gemm_tc_ws for the current slot (see ws_slot).
    None unless CMR_CONV_SPLIT_TAIL=1.  Allocated on first use -- eagerly, i.e. in the warm-up
    step that precedes a graph capture; never from a graph's private pool (None: no split for
    that launch)."""
    if not _split_tail[0]:
        return None
    key = (device.index, _ws_slot[0])
    ws = _conv_ws.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        n = int(_lib.load().cmr_conv_gemm_ws_bytes())
        ws = torch.empty((n,), dtype=torch.uint8, device=device)
        _conv_ws[key] = ws
    return ws


class _GradSideStream(object):
    """Weight (and bias) gradients on a second stream during the backward pass.

    A layer's weight gradient and its data gradient both only need the layer's output
    gradient; the weight gradients are needed by nobody until the optimizer runs.  Issuing
    them on a side stream lets their CTAs fill the SMs that the data-gradient chain leaves
    idle (the partial last wave of every persistent GEMM launch), and vice versa.  Works
    eagerly and under CUDA-graph capture (the event waits become graph edges)."""

    def __init__(self):
        self.stream = None
        self.active = False
        self.keep = []          # operands stay referenced until the join: the caching
                                # allocator must not hand their memory to the main stream

    def begin(self):
        import os
        if os.environ.get('CMR_GRAD_SIDE', '1') == '0':     # A/B measurement knob
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        self.active = True

    def run(self, fn, *tensors):
        if not self.active:
            return fn()
        ev = torch.cuda.Event()
        ev.record()                              # everything the main stream produced so far
        self.keep.extend(t for t in tensors if t is not None)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            return fn()

    def join(self):
        if self.active:
            done = torch.cuda.Event()
            done.record(self.stream)
            torch.cuda.current_stream().wait_event(done)
        self.active = False
        self.keep = []


grad_side = _GradSideStream()


def wgrad_tap(gy, x, gw, rows, cols, loop_hw, gw_ld, gw_col0=0, gy_stride=1, gy_off=(0, 0),
              gy_c0=0, x_stride=1, x_off=(0, 0), x_c0=0, row_scale=None, taps=(1, 1)):
    """gw[i, gw_col0 + j] += row_scale[i] * sum_pix gy[pix_gy, gy_c0+i] * x[pix_x, x_c0+j]
    (cmr_conv_wgrad_tc).  gy (B,gh,gw,ld), x (B,xh,xw,ld) contiguous NHWC."""
    B, gh, gww, gld = gy.shape
    _, xh, xw, xld = x.shape
    d = _lib.WgradDesc(B, loop_hw[0], loop_hw[1], gh, gww, gld, gy_stride, gy_off[0], gy_off[1],
                       gy_c0, xh, xw, xld, x_stride, x_off[0], x_off[1], x_c0, rows, cols,
                       gw_ld, gw_col0, 0, taps[0], taps[1])
    # an int64 destination is the deterministic mode's fixed-point shadow of the gradient
    fn = 'cmr_conv_wgrad_tc_fixed' if gw.dtype == torch.int64 else 'cmr_conv_wgrad_tc'
    grad_side.run(lambda: _lib.call(fn, ctypes.byref(d), _p(gy), _p(x), _p(gw),
                                    _p(row_scale), stream()), gy, x)


def round_tf32(src, dst=None):
    dst = src if dst is None else dst
    _lib.call('cmr_round_tf32', _p(src), _p(dst), src.numel(), stream())
    return dst


def split3(x, order=0, c_pad=None):
    """(..., C) fp32 -> (..., 3 * c_pad): [hi | lo | hi] (order 0, activations) or
    [hi | hi | lo] (order 1, filters) of the 3 x TF32 parity mode (cmr_split_tf32x3)."""
    C = x.shape[-1]
    c_pad = C if c_pad is None else c_pad
    x = x.contiguous()
    out = torch.empty(tuple(x.shape[:-1]) + (3 * c_pad,), dtype=f32, device=x.device)
    _lib.call('cmr_split_tf32x3', _p(x), x.numel() // C, C, c_pad, order, _p(out), stream())
    return out


def column_sums(g, c0, n, out):
    """out[:n] = sum over all leading axes of g[..., c0:c0+n] (g contiguous)."""
    ld = g.shape[-1]
    rows = g.numel() // ld
    fn = 'cmr_col_sum_fixed' if out.dtype == torch.int64 else 'cmr_col_sum'
    grad_side.run(lambda: _lib.call(fn, _p(g), rows, ld, c0, n, _p(out), stream()), g)
    return out


def fixed_to_float(src, dst, accumulate=False, zero_src=True):
    """Deterministic mode: int64 fixed-point accumulators -> fp32 (cmr_fixed_to_float)."""
    _lib.call('cmr_fixed_to_float', _p(src), _p(dst), src.numel(), int(accumulate),
              int(zero_src), stream())
    return dst


_prep_recorder = None   # list collecting descriptors while a batch table is being built


def prep_dgrad_weight(w, O, T, I, stride_o, stride_t, scale, flip, out, ld_out=None, col0=0):
    ld_out = O if ld_out is None else ld_out
    if _prep_recorder is not None:
        _prep_recorder.append((w, scale, out, stride_o, stride_t, O, T, I, int(flip), ld_out, col0))
        return out
    _lib.call('cmr_prep_dgrad_weight', _p(w), O, T, I, stride_o, stride_t, _p(scale), int(flip),
              _p(out), ld_out, col0, stream())
    return out


class PrepTable(object):
    """Every layer's data-gradient filter re-layout as one launch
    (cmr_prep_dgrad_weight_batch).  Built once: the descriptors point into the flat
    parameter buffers and the layers' persistent w_dgrad banks."""

    def __init__(self, layers, device):
        global _prep_recorder
        _prep_recorder = []
        try:
            for l in layers:
                l.prep_backward()
            recs = _prep_recorder
        finally:
            _prep_recorder = None
        self.keep = recs                       # keeps the tensors alive
        arr = (_lib.PrepDesc * max(len(recs), 1))()
        tiles = 0
        for k, (w, scale, out, so, st, O, T, I, flip, ld_out, col0) in enumerate(recs):
            arr[k] = _lib.PrepDesc(w.data_ptr(), scale.data_ptr() if scale is not None else None,
                                   out.data_ptr(), so, st, O, T, I, flip, ld_out, col0, tiles, 0)
            tiles += ((I + 31) // 32) * ((O + 31) // 32) * T
        self.n, self.tiles = len(recs), tiles
        raw = bytes(arr)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)

    def run(self):
        if self.n:
            _lib.call('cmr_prep_dgrad_weight_batch', _p(self.table), self.n, self.tiles, stream())


def max_pool(x, k, stride, pad, cover_all=True):
    B, H, W, C = x.shape
    if cover_all:
        oh = (H + 2 * pad - k + stride - 1) // stride + 1
        ow = (W + 2 * pad - k + stride - 1) // stride + 1
    else:
        oh, ow = conv_out(H, k, stride, pad), conv_out(W, k, stride, pad)
    y = torch.empty((B, oh, ow, C), dtype=f32, device=x.device)
    _lib.call('cmr_max_pool_nhwc', _p(x), B, H, W, C, k, stride, pad, oh, ow, _p(y), stream())
    return y


def avg_pool(x, round_out=True):
    """(R, h, w, C) -> (R, C) mean over h*w."""
    R, h, w, C = x.shape
    y = torch.empty((R, C), dtype=f32, device=x.device)
    _lib.call('cmr_avg_pool_nhwc_fwd', _p(x), R, h * w, C, _p(y), int(round_out), stream())
    return y


def avg_pool_bwd_accum(g, out, mask, round_out=True):
    R, h, w, C = out.shape
    _lib.call('cmr_avg_pool_nhwc_bwd_accum', _p(g), R, h * w, C, _p(out), _p(mask),
              int(round_out), stream())
    return out


def roi_align_nhwc(x, rois_xy, outh, outw, bin_stride, spatial_scale, sampling_ratio=0,
                   round_out=True):
    N, H, W, C = x.shape
    R = rois_xy.shape[0]
    ohs, ows = -(-outh // bin_stride), -(-outw // bin_stride)
    y = torch.empty((R, ohs, ows, C), dtype=f32, device=x.device)
    _lib.call('cmr_roi_align_nhwc_fwd', _p(x), N, H, W, C, _p(rois_xy), R, outh, outw,
              bin_stride, float(spatial_scale), sampling_ratio, int(round_out), _p(y), stream())
    return y


def roi_align_nhwc_bwd(gy, rois_xy, x_shape, outh, outw, bin_stride, spatial_scale,
                       sampling_ratio=0, accum=None, deterministic=False):
    """-> gx (N,H,W,C); with ``accum`` (N,H,W,C) the RoI gradients are added to it in place
    (no zero fill).  ``deterministic``: the scatter goes through int64 fixed-point words
    (order-independent sums), converted to fp32 afterwards."""
    N, H, W, C = x_shape
    if deterministic:
        fx = torch.zeros((N, H, W, C), dtype=torch.int64, device=gy.device)
        _lib.call('cmr_roi_align_nhwc_bwd_fixed', _p(gy), _p(rois_xy), rois_xy.shape[0], N, H, W,
                  C, outh, outw, bin_stride, float(spatial_scale), sampling_ratio, _p(fx),
                  stream())
        out = accum if accum is not None else torch.empty((N, H, W, C), dtype=f32,
                                                          device=gy.device)
        return fixed_to_float(fx, out, accumulate=accum is not None, zero_src=False)
    if accum is not None:
        _lib.call('cmr_roi_align_nhwc_bwd_accum', _p(gy), _p(rois_xy), rois_xy.shape[0], N, H,
                  W, C, outh, outw, bin_stride, float(spatial_scale), sampling_ratio, _p(accum),
                  stream())
        return accum
    gx = torch.empty((N, H, W, C), dtype=f32, device=gy.device)
    _lib.call('cmr_roi_align_nhwc_bwd', _p(gy), _p(rois_xy), rois_xy.shape[0], N, H, W, C, outh,
              outw, bin_stride, float(spatial_scale), sampling_ratio, _p(gx), stream())
    return gx


def relu_mask(g, mask, round_out=True):
    """g * [mask > 0] in place (rounded to tf32)."""
    _lib.call('cmr_relu_mask', _p(g), _p(mask), _p(g), g.numel(), int(round_out), stream())
    return g


def as_nchw_view(x_nhwc):
    """(B,H,W,C) contiguous -> (B,C,H,W) view on the same memory (channels-last)."""
    return x_nhwc.permute(0, 3, 1, 2)


def to_nhwc(x):
    """Accepts an NCHW-shaped tensor; returns a contiguous (B,H,W,C) tensor, without a
    copy when the memory is already channels-last (what this package produces)."""
    return x.permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------ params --
ALIGN = 64  # floats: every parameter starts on a 256-byte boundary (TMA base alignment)


class FlatStore(object):
    """A set of named fp32 parameters carved out of one flat device buffer, so that
    the gradient all-reduce and the SGD update are single launches over it (the
    reference packs gradients the same way inside ChainerMN's communicator,
    examples/train_common.py:99,177-178)."""

    def __init__(self):
        self.specs = []        # (name, shape, offset)
        self.index = {}
        self.size = 0
        self.data = None

    def add(self, name, shape):
        if self.data is not None:
            raise RuntimeError('store already allocated')
        n = int(np.prod(shape))
        off = self.size
        self.specs.append((name, tuple(shape), off))
        self.index[name] = len(self.specs) - 1
        self.size = off + (n + ALIGN - 1) // ALIGN * ALIGN
        return name

    def allocate(self, device):
        self.data = torch.zeros((max(self.size, 4),), dtype=f32, device=device)
        return self

    def view(self, name, buf=None):
        _, shape, off = self.specs[self.index[name]]
        buf = self.data if buf is None else buf
        return buf[off:off + int(np.prod(shape))].view(shape)

    def names(self):
        return [s[0] for s in self.specs]

    def __contains__(self, name):
        return name in self.index
