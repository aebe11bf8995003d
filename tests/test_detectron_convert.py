"""utils/detectron.py: the Detectron blob -> reference parameter-name mapping
(examples/coco/convert_caffe2_to_chainer.py:45-249).  The reference script needs chainer
and a downloaded pickle; here a synthetic blob dictionary with Detectron's names and
shapes is converted and checked against (a) the parameter inventory of the model
(oracle/model.py names and shapes = what state_dict()/load_npz use) and (b) the
reference's explicit assignments for the permuted tensors."""
import numpy as np
import pytest

from chainer_mask_rcnn_b200.utils import detectron
from oracle import model as om


def _blobs(n_layers, rs):
    units = detectron.N_UNITS[n_layers]
    b = {}

    def conv(name, o, i, k, bias=False):
        b[name + '_w'] = rs.standard_normal((o, i, k, k)).astype(np.float32)
        if bias:
            b[name + '_b'] = rs.standard_normal((o,)).astype(np.float32)

    def bn(name, c):
        b[name + '_bn_s'] = rs.standard_normal((c,)).astype(np.float32)
        b[name + '_bn_b'] = rs.standard_normal((c,)).astype(np.float32)

    conv('conv1', 64, 3, 7, bias=True)
    bn('res_conv1', 64)
    cin = 64
    for stage, n in zip((2, 3, 4, 5), units):
        mid, out = 64 * 2 ** (stage - 2), 256 * 2 ** (stage - 2)
        for i in range(n):
            root = 'res%d_%d_' % (stage, i)
            conv(root + 'branch2a', mid, cin, 1); bn(root + 'branch2a', mid)
            conv(root + 'branch2b', mid, mid, 3); bn(root + 'branch2b', mid)
            conv(root + 'branch2c', out, mid, 1); bn(root + 'branch2c', out)
            if i == 0:
                conv(root + 'branch1', out, cin, 1); bn(root + 'branch1', out)
            cin = out
        if stage == 4:
            cin = 1024
    conv('conv_rpn', 1024, 1024, 3, bias=True)
    conv('rpn_bbox_pred', 60, 1024, 1, bias=True)
    conv('rpn_cls_logits', 15, 1024, 1, bias=True)
    b['cls_score_w'] = rs.standard_normal((81, 2048)).astype(np.float32)
    b['cls_score_b'] = rs.standard_normal((81,)).astype(np.float32)
    b['bbox_pred_w'] = rs.standard_normal((324, 2048)).astype(np.float32)
    b['bbox_pred_b'] = rs.standard_normal((324,)).astype(np.float32)
    b['conv5_mask_w'] = rs.standard_normal((2048, 256, 2, 2)).astype(np.float32)
    b['conv5_mask_b'] = rs.standard_normal((256,)).astype(np.float32)
    conv('mask_fcn_logits', 81, 256, 1, bias=True)
    b['fc1000_w'] = np.zeros((1000, 2048), np.float32)          # ignored, like the reference
    b['conv1_w_momentum'] = np.zeros((64, 3, 7, 7), np.float32)
    return b


@pytest.mark.parametrize('n_layers', [50, 101])
def test_names_and_shapes_match_the_model_inventory(n_layers):
    rs = np.random.RandomState(n_layers)
    blobs = _blobs(n_layers, rs)
    params = detectron.detectron_to_params(blobs, n_layers)
    cfg = om.Config(n_layers=n_layers, n_fg_class=80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14)
    want = om.make_params(cfg, np.random.RandomState(0))
    assert sorted(params) == sorted(want)
    for k in want:
        assert params[k].shape == want[k].shape, k
        assert params[k].dtype == np.float32 and params[k].flags.c_contiguous


def test_permutations_follow_the_reference_assignments():
    rs = np.random.RandomState(1)
    blobs = _blobs(50, rs)
    p = detectron.detectron_to_params(blobs, 50)
    # convert_caffe2_to_chainer.py:47  BGR -> RGB on the input-channel axis
    np.testing.assert_array_equal(p['extractor/conv1/W'], blobs['conv1_w'][:, ::-1])
    # :186-195  rpn box regressor, 15 anchors x (dx,dy,dw,dh) -> (dy,dx,dh,dw)
    W = blobs['rpn_bbox_pred_w'].reshape(15, 4, 1024, 1, 1)[:, [1, 0, 3, 2]].reshape(60, 1024, 1, 1)
    np.testing.assert_array_equal(p['rpn/loc/W'], W)
    b = blobs['rpn_bbox_pred_b'].reshape(15, 4)[:, [1, 0, 3, 2]].reshape(60)
    np.testing.assert_array_equal(p['rpn/loc/b'], b)
    # :236-246  class box regressor
    W = blobs['bbox_pred_w'].reshape(81, 4, 2048)[:, [1, 0, 3, 2], :].reshape(324, 2048)
    np.testing.assert_array_equal(p['head/cls_loc/W'], W)
    # :251-252  background mask channel dropped
    np.testing.assert_array_equal(p['head/mask/W'], blobs['mask_fcn_logits_w'][1:])
    np.testing.assert_array_equal(p['head/mask/b'], blobs['mask_fcn_logits_b'][1:])
    # unit naming: res4_3_branch2b -> extractor/res4/b3/conv2, res5_0_branch1 -> head/res5/a/conv4
    np.testing.assert_array_equal(p['extractor/res4/b3/conv2/W'], blobs['res4_3_branch2b_w'])
    np.testing.assert_array_equal(p['head/res5/a/conv4/W'], blobs['res5_0_branch1_w'])
    np.testing.assert_array_equal(p['head/res5/a/bn4/b'], blobs['res5_0_branch1_bn_b'])


def test_pickle_round_trip(tmp_path):
    import pickle
    blobs = _blobs(50, np.random.RandomState(2))
    src = tmp_path / 'model_final.pkl'
    with open(src, 'wb') as f:
        pickle.dump({'blobs': blobs}, f)
    dst = detectron.convert(str(src), str(tmp_path / 'w.npz'))
    with np.load(dst) as z:
        assert len(z.files) == 174
        np.testing.assert_array_equal(z['rpn/score/W'], blobs['rpn_cls_logits_w'])
