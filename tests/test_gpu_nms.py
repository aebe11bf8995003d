"""NMS and proposal generation on the B200 vs the oracle: integer outputs must be
bit-exact (keep lists, 64-bit suppression masks, proposal anchor indices)."""
import numpy as np
import pytest
import torch

import synth
from chainer_mask_rcnn_b200 import utils
from oracle import bbox as ob

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [1, 2, 63, 64, 65, 300, 1000, 2000])
@pytest.mark.parametrize('thresh', [0.7, 0.5])
def test_keep_list_and_bitmask_bit_exact(n, thresh):
    rs = np.random.RandomState(n)
    boxes = synth.clustered_boxes(rs, n, 800, 1088, n_centers=max(2, n // 12))
    keep, mask = utils.nms_suppression_bitmask(boxes, thresh)
    want_mask = ob.nms_suppression_bitmask(boxes, thresh)
    nb = (n + 63) // 64
    # words below the diagonal block are never produced nor read
    tri = np.arange(nb)[None, :] >= (np.arange(n) // 64)[:, None]
    np.testing.assert_array_equal(mask[tri], want_mask[tri])
    want = ob.non_maximum_suppression(boxes, thresh)
    assert keep.dtype == np.int32
    np.testing.assert_array_equal(keep, want)


def test_empty_score_limit_and_degenerate():
    assert utils.non_maximum_suppression(np.zeros((0, 4), np.float32), 0.5).shape == (0,)
    rs = np.random.RandomState(9)
    boxes = synth.clustered_boxes(rs, 500, 600, 600)
    boxes[10] = boxes[11] = [5, 5, 5, 5]             # zero area: IoU is NaN, never suppresses
    score = synth.tie_free_scores(rs, 500)
    for limit in (None, 1, 17):
        got = utils.non_maximum_suppression(boxes, 0.5, score=score, limit=limit)
        want = ob.non_maximum_suppression(boxes, 0.5, score=score, limit=limit)
        np.testing.assert_array_equal(got, want)
    t = torch.from_numpy(boxes).cuda()
    out = utils.non_maximum_suppression(t, 0.7)
    assert out.is_cuda and out.dtype == torch.int32
    np.testing.assert_array_equal(out.cpu().numpy(), ob.non_maximum_suppression(boxes, 0.7))


def test_large_train_size_keep_list():
    """n = 6000 with the train-time limit, the reference's per-image test budget."""
    rs = np.random.RandomState(123)
    boxes = synth.clustered_boxes(rs, 6000, 800, 1333, n_centers=300)
    got = utils.non_maximum_suppression(boxes, 0.7, limit=1000)
    want = ob.non_maximum_suppression(boxes, 0.7, limit=1000)
    np.testing.assert_array_equal(got, want)


def _anchors(fh, fw, scales):
    base = ob.generate_anchor_base(16, (0.5, 1, 2), scales)
    return ob.enumerate_shifted_anchor(base, 16, fh, fw)


@pytest.mark.parametrize('fh,fw,scales,train', [
    (12, 16, (4, 8, 16, 32), False),        # small, not a power of two
    (38, 63, (4, 8, 16, 32), False),        # BASELINE config 1: VOC 600x1000, 28728 anchors
    (51, 84, (2, 4, 8, 16, 32), True),      # BASELINE config 2: COCO 800x1333, 64260 anchors
])
def test_proposal_indices_bit_exact(fh, fw, scales, train):
    rs = np.random.RandomState(fh)
    anchor = _anchors(fh, fw, scales)
    img = (fh * 16 - 7, fw * 16 - 11)
    loc, score = synth.rpn_outputs(rs, len(anchor))
    params = dict(min_size=0, n_test_pre_nms=6000, n_test_post_nms=1000)
    want_roi, want_idx = ob.ProposalCreator(**params)(loc, score, anchor, img, scale=1.6,
                                                      train=train, return_index=True)
    pc = utils.ProposalCreator(**params)
    with utils.config.using_config('train', train):
        roi, idx = pc(loc, score, anchor, img, scale=1.6, return_index=True)
    np.testing.assert_array_equal(idx, want_idx)
    # coordinates: exp() is evaluated in fp64 on the device, NumPy's fp32 exp may be
    # 1-2 ulp off; everything else is the same fp32 expression
    np.testing.assert_allclose(roi, want_roi, rtol=1e-6, atol=1e-4)
    assert roi.shape[0] <= (2000 if train else 1000)


def test_proposal_min_size_and_batch():
    rs = np.random.RandomState(77)
    anchor = _anchors(20, 30, (4, 8, 16, 32))
    img = (320, 480)
    locs, scores = zip(*[synth.rpn_outputs(rs, len(anchor)) for _ in range(3)])
    params = dict(min_size=16, n_test_pre_nms=3000, n_test_post_nms=300)
    pc = utils.ProposalCreator(**params)
    L = torch.from_numpy(np.stack(locs)).cuda()
    S = torch.from_numpy(np.stack(scores)).cuda()
    A = torch.from_numpy(anchor).cuda()
    rois, idx, cnt = pc.batch(L, S, A, img, scale=1.5, train=False)
    cnt = cnt.cpu().numpy()
    for b in range(3):
        want_roi, want_idx = ob.ProposalCreator(**params)(locs[b], scores[b], anchor, img,
                                                          scale=1.5, train=False,
                                                          return_index=True)
        assert cnt[b] == len(want_idx)
        np.testing.assert_array_equal(idx[b, :cnt[b]].cpu().numpy(), want_idx)
        np.testing.assert_allclose(rois[b, :cnt[b]].cpu().numpy(), want_roi, rtol=1e-6, atol=1e-4)
        assert (idx[b, cnt[b]:] == -1).all()
