"""ROIAlign operator with the reference's interface.

Mirrors ``chainer_mask_rcnn/functions/roi_align_2d.py``: class ``ROIAlign2D``
(:25-524; argument checks :29-47, type checks :49-59) and the wrapper
``roi_align_2d`` (:527-560; ``axes`` handling :555-558).  Both passes run in
``cmr_roi_align_fwd_cl / _bwd_cl`` (feature map channels-last, pooled tensor in the
reference's layout), or in ``cmr_roi_align_fwd / _bwd`` (reference layout AND summation
order: channel counts that are not a multiple of 4, or ``ROIAlign2D.exact_order``);
csrc/roi_align.cu.
"""
import torch

from .. import _lib
from .._array import from_device, InvalidType, to_device


# Both passes run on the channels-last kernels (cmr_roi_align_fwd_cl / _bwd_cl): vector loads /
# vector reductions of 4 channels on the feature map, the pooled tensor kept in the
# reference's (R, C, outh, outw) layout and moved as full 128-byte lines through a
# shared-memory tile.  A feature map that already is channels-last in memory (what this
# package's extractor returns) is used as it is; a plain NCHW map is re-laid once into a
# workspace (cmr_roi_align_fwd_ws / _bwd_ws; the map is ~2 % of the pooled tensor's bytes).
# ``ROIAlign2D.exact_order = True`` selects the reference-layout kernels, which also keep the
# reference's summation order (cmr_roi_align_fwd / _bwd); channel counts that are not a
# multiple of 4 always take them.
def _fast_ok(N, C, H, W, R, outh, outw):
    return (not ROIAlign2D.exact_order and
            _lib.load().cmr_roi_align_cl_supported(N, C, H, W, R, outh, outw) == 1)


def _is_channels_last(x):
    return x.dim() == 4 and x.shape[1] > 1 and x.permute(0, 2, 3, 1).is_contiguous()


def _forward(x, rois, outh, outw, spatial_scale, sampling_ratio):
    N, C, H, W = x.shape
    R = rois.shape[0]
    rois = rois.contiguous()
    y = torch.empty((R, C, outh, outw), dtype=torch.float32, device=x.device)
    if _fast_ok(N, C, H, W, R, outh, outw):
        if _is_channels_last(x):
            _lib.call('cmr_roi_align_fwd_cl', _lib.ptr(x.permute(0, 2, 3, 1)), N, H, W, C,
                      _lib.ptr(rois), R, outh, outw, spatial_scale, sampling_ratio, _lib.ptr(y),
                      _lib.stream_ptr())
            return y
        x = x.contiguous()
        ws = torch.empty((N * C * H * W,), dtype=torch.float32, device=x.device)
        _lib.call('cmr_roi_align_fwd_ws', _lib.ptr(x), N, C, H, W, _lib.ptr(rois), R, outh, outw,
                  spatial_scale, sampling_ratio, _lib.ptr(y), _lib.ptr(ws), ws.numel() * 4,
                  _lib.stream_ptr())
        return y
    x = x.contiguous()
    _lib.call('cmr_roi_align_fwd', _lib.ptr(x), N, C, H, W, _lib.ptr(rois), R, outh,
              outw, spatial_scale, sampling_ratio, _lib.ptr(y), _lib.stream_ptr())
    return y


def _backward(gy, rois, shape, outh, outw, spatial_scale, sampling_ratio, channels_last=False):
    N, C, H, W = shape
    R = rois.shape[0]
    rois = rois.contiguous()
    gy = gy.contiguous()
    if _fast_ok(N, C, H, W, R, outh, outw):
        if channels_last:       # the gradient in the layout of the map it belongs to
            gxt = torch.empty((N, H, W, C), dtype=torch.float32, device=gy.device)
            _lib.call('cmr_roi_align_bwd_cl', _lib.ptr(gy), _lib.ptr(rois), R, N, H, W, C, outh,
                      outw, spatial_scale, sampling_ratio, _lib.ptr(gxt), _lib.stream_ptr())
            return gxt.permute(0, 3, 1, 2)
        gx = torch.empty((N, C, H, W), dtype=torch.float32, device=gy.device)
        ws = torch.empty((N * C * H * W,), dtype=torch.float32, device=gy.device)
        _lib.call('cmr_roi_align_bwd_ws', _lib.ptr(gy), _lib.ptr(rois), R, N, C, H, W, outh,
                  outw, spatial_scale, sampling_ratio, _lib.ptr(gx), _lib.ptr(ws),
                  ws.numel() * 4, _lib.stream_ptr())
        return gx
    gx = torch.empty((N, C, H, W), dtype=torch.float32, device=gy.device)
    _lib.call('cmr_roi_align_bwd', _lib.ptr(gy), _lib.ptr(rois), R, N, C, H, W, outh, outw,
              spatial_scale, sampling_ratio, _lib.ptr(gx), _lib.stream_ptr())
    return gx


class _ROIAlignFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, rois, outh, outw, spatial_scale, sampling_ratio):
        y = _forward(x, rois, outh, outw, spatial_scale, sampling_ratio)
        ctx.save_for_backward(rois)      # only the rois are retained (:62-63)
        ctx.meta = (tuple(x.shape), outh, outw, spatial_scale, sampling_ratio,
                    _is_channels_last(x))
        return y

    @staticmethod
    def backward(ctx, gy):
        rois, = ctx.saved_tensors
        shape, outh, outw, spatial_scale, sampling_ratio, cl = ctx.meta
        gx = _backward(gy, rois, shape, outh, outw, spatial_scale, sampling_ratio, cl)
        return gx, None, None, None, None, None


class ROIAlign2D(object):

    """ROI align over a set of 2d planes (function object, as in the reference)."""

    exact_order = False      # True: always the reference-layout, reference-order kernels

    def __init__(self, outh, outw, spatial_scale, sampling_ratio=0):
        for name, value in (('outh', outh), ('outw', outw),
                            ('sampling_ratio', sampling_ratio)):
            if not (isinstance(value, int) and not isinstance(value, bool) and value >= 0):
                raise TypeError('{} must be positive integer: {}, {}'.format(
                    name, type(value), value))
        if isinstance(spatial_scale, int):
            spatial_scale = float(spatial_scale)
        elif not isinstance(spatial_scale, float):
            raise TypeError('spatial_scale must be float: {}'.format(type(spatial_scale)))
        self.outh, self.outw = outh, outw
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    @staticmethod
    def check_type_forward(x, rois):
        if x.dtype != torch.float32 or x.dim() != 4:
            raise InvalidType('x must be a 4-d float32 array, got {} {}'.format(
                x.dtype, tuple(x.shape)))
        if rois.dtype != torch.float32 or rois.dim() != 2 or rois.shape[1] != 5:
            raise InvalidType('rois must be a (R, 5) float32 array, got {} {}'.format(
                rois.dtype, tuple(rois.shape)))

    def __call__(self, x, rois):
        x, x_np = to_device(x, keep_layout=True)
        rois, _ = to_device(rois)
        self.check_type_forward(x, rois)
        y = _ROIAlignFn.apply(x, rois, self.outh, self.outw, self.spatial_scale,
                              self.sampling_ratio)
        return from_device(y, x_np)

    # Array-level entry points with the reference's names (roi_align_2d.py:162,391).
    def forward_gpu(self, inputs):
        x, rois = inputs
        self._bottom_data_shape = tuple(x.shape)
        with torch.no_grad():
            return self(x, rois),

    def backward_gpu(self, inputs, gy):
        rois, _ = to_device(inputs[1])
        g, g_np = to_device(gy[0])
        gx = _backward(g, rois, self._bottom_data_shape, self.outh, self.outw,
                       self.spatial_scale, self.sampling_ratio)
        return from_device(gx, g_np), None


def roi_align_2d(x, rois, outh, outw, spatial_scale, sampling_ratio=0, axes='xy'):
    """Spatial Region of Interest (ROI) align function.

    Args:
        x: (N, C, H, W) float32 array (torch CUDA tensor, or numpy array which is
            copied to the GPU and back).
        rois: (R, 5) float32, each row (batch_index, x_min, y_min, x_max, y_max) for
            ``axes='xy'`` or (batch_index, y_min, x_min, y_max, x_max) for ``'yx'``.
        outh, outw (int): pooled output size.
        spatial_scale (float): scale applied to the roi coordinates.
        sampling_ratio (int): samples per bin and axis; 0 = adaptive.

    Returns: (R, C, outh, outw) array, differentiable with respect to ``x``.
    """
    if axes not in ['xy', 'yx']:
        raise ValueError('Unsupported axes: {}'.format(axes))
    if axes == 'yx':
        rois = rois[:, [0, 2, 1, 4, 3]]
    return ROIAlign2D(outh, outw, spatial_scale, sampling_ratio)(x, rois)
