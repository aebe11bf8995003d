"""Detectron (Caffe2) blob dictionary -> parameters under the reference's names.

The role of ``examples/coco/convert_caffe2_to_chainer.py:45-249``: that script copies the
blobs of a Detectron ``e2e_mask_rcnn_R-50-C4`` / ``R-101-C4`` pickle into a Chainer
``MaskRCNNResNet`` one assignment at a time.  Here the same mapping is expressed as rules
and returns a ``{name: ndarray}`` dict that ``MaskRCNN.load_state_dict`` / ``np.savez``
(-> ``load_npz`` / ``pretrained_model=``) accept:

* ``conv1_w`` is BGR in Detectron: the input-channel axis is reversed (``[:, ::-1]``, :47);
* residual units: ``res{s}_{i}_branch2{a,b,c}`` -> ``conv{1,2,3}``, ``branch1`` -> ``conv4``,
  ``*_bn_s`` / ``*_bn_b`` -> the AffineChannel2D ``bn*/W`` / ``bn*/b``; unit 0 is ``a``,
  unit i >= 1 is ``b{i}``; stages 2-4 live under ``extractor/``, stage 5 under ``head/``;
* box regressors are (dx, dy, dw, dh) per anchor / class in Detectron and (dy, dx, dh, dw)
  here: ``rpn_bbox_pred`` (:186-195) and ``bbox_pred`` (:236-246) have their groups of four
  output channels permuted ``[1, 0, 3, 2]``;
* ``mask_fcn_logits`` drops its background channel (``[1:]``, :251-252).
"""
import pickle

import numpy as np

N_UNITS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}
_BRANCH = (('branch2a', 1), ('branch2b', 2), ('branch2c', 3), ('branch1', 4))


def _swap_yx(a, groups):
    """Permute every group of four leading entries (dx,dy,dw,dh) -> (dy,dx,dh,dw)."""
    rest = a.shape[1:]
    return a.reshape((groups, 4) + rest)[:, [1, 0, 3, 2]].reshape((groups * 4,) + rest)


def load_blobs(path):
    """The ``blobs`` dict of a Detectron ``model_final.pkl``."""
    with open(path, 'rb') as f:
        return pickle.load(f, encoding='latin-1')['blobs']


def detectron_to_params(blobs, n_layers=50, n_anchor=15, n_class=81):
    """-> dict of float32 arrays keyed by the reference's parameter names
    ('extractor/res4/b3/conv2/W', 'head/cls_loc/b', ...)."""
    p = {}

    def put(name, value):
        p[name] = np.ascontiguousarray(value, dtype=np.float32)

    put('extractor/conv1/W', blobs['conv1_w'][:, ::-1])
    put('extractor/conv1/b', blobs['conv1_b'])
    put('extractor/bn1/W', blobs['res_conv1_bn_s'])
    put('extractor/bn1/b', blobs['res_conv1_bn_b'])
    for stage, n_unit in zip((2, 3, 4, 5), N_UNITS[n_layers]):
        root = ('head' if stage == 5 else 'extractor') + '/res%d' % stage
        for i in range(n_unit):
            unit = 'a' if i == 0 else 'b%d' % i
            for branch, k in _BRANCH:
                if k == 4 and i > 0:
                    continue                       # only the first unit has a projection
                src = 'res%d_%d_%s' % (stage, i, branch)
                put('%s/%s/conv%d/W' % (root, unit, k), blobs[src + '_w'])
                put('%s/%s/bn%d/W' % (root, unit, k), blobs[src + '_bn_s'])
                put('%s/%s/bn%d/b' % (root, unit, k), blobs[src + '_bn_b'])
    put('rpn/conv1/W', blobs['conv_rpn_w'])
    put('rpn/conv1/b', blobs['conv_rpn_b'])
    put('rpn/loc/W', _swap_yx(blobs['rpn_bbox_pred_w'], n_anchor))
    put('rpn/loc/b', _swap_yx(blobs['rpn_bbox_pred_b'], n_anchor))
    put('rpn/score/W', blobs['rpn_cls_logits_w'])
    put('rpn/score/b', blobs['rpn_cls_logits_b'])
    put('head/score/W', blobs['cls_score_w'])
    put('head/score/b', blobs['cls_score_b'])
    put('head/cls_loc/W', _swap_yx(blobs['bbox_pred_w'], n_class))
    put('head/cls_loc/b', _swap_yx(blobs['bbox_pred_b'], n_class))
    put('head/deconv6/W', blobs['conv5_mask_w'])
    put('head/deconv6/b', blobs['conv5_mask_b'])
    put('head/mask/W', blobs['mask_fcn_logits_w'][1:])
    put('head/mask/b', blobs['mask_fcn_logits_b'][1:])
    return p


def convert(src_pkl, dst_npz, n_layers=50):
    """Detectron pickle -> npz loadable with ``MaskRCNNResNet(pretrained_model=dst_npz)``."""
    np.savez(dst_npz, **detectron_to_params(load_blobs(src_pkl), n_layers))
    return dst_npz
