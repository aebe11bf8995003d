"""Size-independent properties at BASELINE.json's full sizes (the oracle comparisons in
the other files stop at sizes a NumPy loop finishes in seconds): greedy NMS of the 12000
train-time candidates is idempotent, leaves no overlapping pair and suppresses only with
a witness; proposals come out in score order inside the image; ROIAlign is linear in its
input on the 1024 x 50 x 68 map of configs[4]."""
import numpy as np
import pytest
import torch

import synth
from chainer_mask_rcnn_b200 import functions, utils
from oracle import bbox as ob

pytestmark = pytest.mark.gpu


def test_nms_12000_idempotent_and_witnessed():
    rs = np.random.RandomState(2024)
    n, thresh = 12000, 0.7
    boxes = synth.clustered_boxes(rs, n, 800, 1333, n_centers=600)
    keep = utils.non_maximum_suppression(boxes, thresh)
    assert keep.dtype == np.int32 and keep[0] == 0            # the best box always survives
    assert (np.diff(keep) > 0).all() and keep[-1] < n          # indices in input (score) order
    kept = boxes[keep]
    again = utils.non_maximum_suppression(kept, thresh)
    np.testing.assert_array_equal(again, np.arange(len(keep), dtype=np.int32))   # idempotent
    # no two survivors overlap by the threshold (first 500 of them, all pairs)
    iou = ob.bbox_iou(kept[:500], kept[:500])
    np.fill_diagonal(iou, 0.)
    assert not (iou >= thresh + 1e-5).any()
    # every suppressed box has a better-scored survivor that overlaps it by the threshold
    suppressed = np.setdiff1d(np.arange(n), keep)
    assert len(suppressed) > 1000
    sample = suppressed[rs.permutation(len(suppressed))[:300]]
    iou = ob.bbox_iou(boxes[sample], kept)
    earlier = keep[None, :] < sample[:, None]
    assert ((iou >= thresh - 1e-5) & earlier).any(axis=1).all()


def test_proposals_full_size_sorted_and_clipped():
    rs = np.random.RandomState(7)
    fh, fw = 51, 84
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (2, 4, 8, 16, 32))
    anchor = ob.enumerate_shifted_anchor(base, 16, fh, fw)
    loc, score = synth.rpn_outputs(rs, len(anchor))
    img = (800, 1333)
    pc = utils.ProposalCreator(min_size=0, n_test_pre_nms=6000, n_test_post_nms=1000)
    for train, n_post in ((True, 2000), (False, 1000)):
        with utils.config.using_config('train', train):
            roi, idx = pc(loc, score, anchor, img, scale=1.6, return_index=True)
        assert 0 < len(idx) <= n_post and len(np.unique(idx)) == len(idx)
        assert (np.diff(score[idx]) < 0).all()                 # descending score, tie-free input
        assert (roi[:, 0] >= 0).all() and (roi[:, 1] >= 0).all()
        assert (roi[:, 2] <= img[0]).all() and (roi[:, 3] <= img[1]).all()
        assert (roi[:, 2] >= roi[:, 0]).all() and (roi[:, 3] >= roi[:, 1]).all()


def test_roi_align_linear_at_full_size():
    rs = np.random.RandomState(3)
    x1 = torch.from_numpy(rs.standard_normal((1, 1024, 50, 68)).astype(np.float32)).cuda()
    x2 = torch.from_numpy(rs.standard_normal((1, 1024, 50, 68)).astype(np.float32)).cuda()
    rois = torch.from_numpy(synth.rois_xy(rs, 300, 1, 800, 1088)).cuda()

    def pool(x):
        return functions.roi_align_2d(x, rois, 14, 14, 1. / 16)
    want = pool(x1) + 2. * pool(x2)
    got = pool(x1 + 2. * x2)
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
