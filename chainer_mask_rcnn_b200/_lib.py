"""ctypes binding of libcmr_b200.so (declared in include/cmr_b200.h).

There is no CPU fallback: if the library is missing, or a call returns a
non-zero status, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcmr_b200.so')

c_int = ctypes.c_int
c_float = ctypes.c_float
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
c_longlong = ctypes.c_longlong
c_ulonglong = ctypes.c_ulonglong

# name -> (restype, argtypes); mirrors include/cmr_b200.h one to one.
_SIGNATURES = {
    'cmr_status_string': (ctypes.c_char_p, [c_int]),
    'cmr_version': (c_int, []),
    'cmr_last_cuda_error': (ctypes.c_char_p, []),
    'cmr_launch_count': (c_longlong, []),
    'cmr_prof_enable': (c_int, [c_int]),
    'cmr_prof_collect': (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    'cmr_roi_align_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                  c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    'cmr_roi_align_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    'cmr_roi_align_nhwc_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                       c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                       c_void_p, c_void_p]),
    'cmr_roi_align_nhwc_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_float, c_int,
                                       c_void_p, c_void_p]),
    'cmr_roi_align_nhwc_bwd_accum': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_int, c_int, c_float, c_int,
                                             c_void_p, c_void_p]),
    'cmr_roi_align_cl_supported': (c_int, [c_int] * 7),
    'cmr_roi_align_fwd_cl': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                     c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    'cmr_roi_align_bwd_cl': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    'cmr_roi_align_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'cmr_roi_align_fwd_ws': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                     c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    'cmr_roi_align_bwd_ws': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    'cmr_transpose_batched': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'cmr_affine_channel_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_void_p]),
    'cmr_affine_channel_bwd_workspace_bytes': (c_size_t, [c_int, c_int]),
    'cmr_affine_channel_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    'cmr_bn_fold': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p,
                            c_void_p, c_void_p]),
    'cmr_nms_workspace_bytes': (c_size_t, [c_int]),
    'cmr_nms': (c_int, [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p,
                        c_size_t, c_void_p]),
    'cmr_proposals_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'cmr_proposals': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                              c_float, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    'cmr_set_im2col_tma': (c_int, [c_int]),
    'cmr_set_conv_variant': (c_int, [c_int]),
    'cmr_set_conv_debug': (c_int, [c_void_p]),
    'cmr_conv_gemm_tc': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    'cmr_conv_gemm_tc_ex': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    'cmr_conv_gemm_ws_bytes': (c_size_t, []),
    'cmr_conv_gemm_tc_ws': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                    c_size_t, c_void_p]),
    'cmr_conv_wgrad_tc': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'cmr_round_tf32': (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    'cmr_conv_wgrad_tc_fixed': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    'cmr_col_sum_fixed': (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p,
                                  c_void_p]),
    'cmr_roi_align_nhwc_bwd_fixed': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_int, c_int, c_float, c_int,
                                             c_void_p, c_void_p]),
    'cmr_fixed_to_float': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'cmr_relu_mask': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    'cmr_split_tf32x3': (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p]),
    'cmr_pack_image_nhwc4': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    'cmr_max_pool_nhwc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p, c_void_p]),
    'cmr_avg_pool_nhwc_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'cmr_avg_pool_nhwc_bwd_accum': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                            c_int, c_void_p]),
    'cmr_col_sum': (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p]),
    'cmr_prep_dgrad_weight': (c_int, [c_void_p, c_int, c_int, c_int, c_longlong, c_longlong,
                                      c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    'cmr_prep_dgrad_weight_batch': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'cmr_sgd_momentum': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float,
                                 c_float, c_float, c_void_p, c_void_p]),
    'cmr_rpn_loss': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_longlong,
                             c_int, c_float, c_void_p, c_int, c_void_p, c_void_p]),
    'cmr_roi_loss': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                             c_float, c_void_p, c_int, c_void_p, c_void_p]),
    'cmr_anchor_targets_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'cmr_anchor_targets': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float,
                                   c_float, c_int, c_float, c_float, c_float, c_ulonglong,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    'cmr_proposal_targets_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'cmr_proposal_targets': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_ulonglong, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    'cmr_detections_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'cmr_detections': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                               c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    'cmr_paste_masks': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_longlong,
                                c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p]),
    'cmr_prepare_size': (c_int, [c_int, c_int, c_double, c_double, c_void_p, c_void_p]),
    'cmr_prepare_image': (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_float, c_float,
                                  c_float, c_void_p, c_int, c_int, c_void_p]),
    'cmr_mask_targets': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'cmr_mask_loss': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_void_p, c_int, c_void_p, c_void_p]),
}


class ConvDesc(ctypes.Structure):
    """struct cmr_conv_desc (include/cmr_b200.h)."""
    _fields_ = [(n, c_int) for n in (
        'batch', 'in_h', 'in_w', 'in_c', 'in_ld', 'out_h', 'out_w', 'kh', 'kw', 'stride',
        'pad', 'n', 'd_h', 'd_w', 'd_ld', 'd_stride', 'd_oy', 'd_ox', 'relu', 'round_tf32',
        'tile_n', 'tap_cols')]


class PrepDesc(ctypes.Structure):
    """struct cmr_prep_desc (include/cmr_b200.h)."""
    _fields_ = [('w', c_void_p), ('scale', c_void_p), ('out', c_void_p),
                ('stride_o', c_longlong), ('stride_t', c_longlong)] + \
               [(n, c_int) for n in ('O', 'T', 'I', 'flip', 'ld_out', 'col0', 'tile_begin',
                                     'reserved')]


class WgradDesc(ctypes.Structure):
    """struct cmr_wgrad_desc (include/cmr_b200.h)."""
    _fields_ = [(n, c_int) for n in (
        'batch', 'loop_h', 'loop_w', 'gy_h', 'gy_w', 'gy_ld', 'gy_stride', 'gy_off_y',
        'gy_off_x', 'gy_c0', 'x_h', 'x_w', 'x_ld', 'x_stride', 'x_off_y', 'x_off_x', 'x_c0',
        'rows', 'cols', 'gw_ld', 'gw_col0', 'splits', 'taps_h', 'taps_w')]


EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


class CmrError(RuntimeError):
    """A libcmr_b200 entry point returned a non-zero cmr_status."""


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libcmr_b200.so is missing at {} -- run `python -m '
            'chainer_mask_rcnn_b200.build` (or __graft_entry__.build()). There is no CPU '
            'fallback.'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.cmr_status_string(status).decode()
        detail = lib.cmr_last_cuda_error().decode() if status == -2 else ''
        raise CmrError('{} failed: {} ({}) {}'.format(what, msg, status, detail))


def call(name, *args):
    """Call an int-status entry point and raise CmrError on failure."""
    check(getattr(load(), name)(*args), name)


def stream_ptr():
    """cudaStream_t of torch's current stream, as a void*."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return ctypes.c_void_p(t.data_ptr())
