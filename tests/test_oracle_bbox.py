"""CPU checks of the chainercv restatements in oracle/bbox.py.  No golden vectors
exist for these in the reference (parity unpinned); the tests pin
the restatement against independent brute-force formulations and known values."""
import numpy as np

import synth
from oracle import bbox as ob


def test_anchor_base_known_values():
    a = ob.generate_anchor_base(16, (0.5, 1, 2), (8, 16, 32))
    assert a.shape == (9, 4) and a.dtype == np.float32
    # ratio 1, scale 8: a 128 x 128 box centred on (8, 8)
    np.testing.assert_allclose(a[3], [-56, -56, 72, 72], atol=1e-4)
    # ratio 0.5 (h = 128*sqrt(.5)), scale 8
    np.testing.assert_allclose(a[0], [8 - 45.2548, 8 - 90.5097, 8 + 45.2548, 8 + 90.5097], atol=1e-3)


def test_shifted_anchor_order():
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (4, 8))
    anc = ob.enumerate_shifted_anchor(base, 16, 3, 5)
    assert anc.shape == (3 * 5 * 6, 4)
    # K row-major over (y, x), A innermost
    k = (2 * 5 + 4) * 6 + 1
    np.testing.assert_allclose(anc[k], base[1] + np.array([32, 64, 32, 64]), atol=1e-5)


def test_loc2bbox_bbox2loc_roundtrip():
    rs = np.random.RandomState(0)
    src = synth.random_boxes(rs, 50, 600, 800)
    dst = synth.random_boxes(rs, 50, 600, 800)
    loc = ob.bbox2loc(src, dst).astype(np.float32)
    back = ob.loc2bbox(src, loc)
    np.testing.assert_allclose(back, dst, atol=2e-2)


def _brute_iou(a, b):
    out = np.zeros((len(a), len(b)), np.float64)
    for i, p in enumerate(a.astype(np.float64)):
        for j, q in enumerate(b.astype(np.float64)):
            h = min(p[2], q[2]) - max(p[0], q[0])
            w = min(p[3], q[3]) - max(p[1], q[1])
            inter = h * w if (h > 0 and w > 0) else 0.
            ua = (p[2] - p[0]) * (p[3] - p[1]) + (q[2] - q[0]) * (q[3] - q[1]) - inter
            out[i, j] = inter / ua
    return out


def test_bbox_iou_matches_bruteforce():
    rs = np.random.RandomState(1)
    a = synth.random_boxes(rs, 30, 300, 400)
    b = synth.random_boxes(rs, 20, 300, 400)
    np.testing.assert_allclose(ob.bbox_iou(a, b), _brute_iou(a, b), atol=1e-5)


def _greedy_reference(boxes, thresh):
    iou = _brute_iou(boxes, boxes)
    keep = []
    for i in range(len(boxes)):
        if all(iou[i, j] < thresh for j in keep):
            keep.append(i)
    return np.array(keep, np.int32)


def test_nms_matches_independent_greedy():
    rs = np.random.RandomState(2)
    boxes = synth.clustered_boxes(rs, 400, 600, 800)
    got = ob.non_maximum_suppression(boxes, 0.7)
    want = _greedy_reference(boxes, 0.7)
    np.testing.assert_array_equal(got, want)
    assert 0 < len(got) < 400


def test_nms_score_and_limit():
    rs = np.random.RandomState(3)
    boxes = synth.clustered_boxes(rs, 200, 300, 300)
    score = synth.tie_free_scores(rs, 200)
    sel = ob.non_maximum_suppression(boxes, 0.5, score=score, limit=7)
    assert len(sel) == 7 and sel.dtype == np.int32
    assert (np.diff(score[sel]) < 0).all()          # descending score order
    assert ob.non_maximum_suppression(boxes[:0], 0.5).shape == (0,)


def _sweep(mask, n):
    """Host sweep of the suppression bitmask (what the chainercv GPU path does)."""
    nb = mask.shape[1]
    remv = np.zeros(nb, np.uint64)
    keep = []
    for i in range(n):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            keep.append(i)
            remv |= mask[i]
    return np.array(keep, np.int32)


def test_bitmask_sweep_equals_greedy():
    rs = np.random.RandomState(4)
    for n in (1, 63, 64, 65, 300):
        boxes = synth.clustered_boxes(rs, n, 400, 400)
        mask = ob.nms_suppression_bitmask(boxes, 0.7)
        assert mask.shape == (n, (n + 63) // 64)
        np.testing.assert_array_equal(_sweep(mask, n), ob.non_maximum_suppression(boxes, 0.7))


def test_proposal_creator_contract():
    rs = np.random.RandomState(5)
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (4, 8, 16, 32))
    anchor = ob.enumerate_shifted_anchor(base, 16, 12, 16)
    loc, score = synth.rpn_outputs(rs, len(anchor))
    pc = ob.ProposalCreator(min_size=0, n_test_pre_nms=600, n_test_post_nms=100)
    roi, idx = pc(loc, score, anchor, (192, 256), scale=1., train=False, return_index=True)
    assert roi.shape[1] == 4 and len(roi) <= 100 and roi.dtype == np.float32
    assert (roi[:, 0] >= 0).all() and (roi[:, 2] <= 192).all() and (roi[:, 3] <= 256).all()
    assert (np.diff(score[idx]) < 0).all()
    # the returned index really is the anchor the proposal was decoded from
    dec = ob.loc2bbox(anchor, loc)
    dec[:, 0::2] = np.clip(dec[:, 0::2], 0, 192)
    dec[:, 1::2] = np.clip(dec[:, 1::2], 0, 256)
    np.testing.assert_array_equal(dec[idx], roi)
    # min_size removes small boxes
    pc2 = ob.ProposalCreator(min_size=16, n_test_pre_nms=600, n_test_post_nms=100)
    roi2 = pc2(loc, score, anchor, (192, 256), scale=2., train=False)
    assert ((roi2[:, 2] - roi2[:, 0]) >= 32).all() and ((roi2[:, 3] - roi2[:, 1]) >= 32).all()


def test_anchor_target_creator_contract():
    rs = np.random.RandomState(6)
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (4, 8, 16))
    anchor = ob.enumerate_shifted_anchor(base, 16, 20, 25)
    bbox = synth.random_boxes(rs, 6, 320, 400, 40., 200.)
    loc, label = ob.AnchorTargetCreator()(bbox, anchor, (320, 400), rng=np.random.RandomState(0))
    assert loc.shape == (len(anchor), 4) and label.shape == (len(anchor),)
    assert set(np.unique(label)) <= {-1, 0, 1}
    assert (label == 1).sum() <= 128 and (label >= 0).sum() <= 256
    assert (label == 1).sum() >= 1
    outside = (anchor[:, 0] < 0) | (anchor[:, 1] < 0) | (anchor[:, 2] > 320) | (anchor[:, 3] > 400)
    assert (label[outside] == -1).all() and (loc[outside] == 0).all()


def test_cross_check_against_torchvision_cpu():
    """chainercv is absent, so these restatements are 'parity unpinned';
    torchvision's CPU ops are an implementation nobody here wrote: same IoU matrix, same
    greedy keep list (torchvision suppresses on IoU > thresh, chainercv on >=: identical on
    boxes without an exact tie), same box decoding as its BoxCoder with unit weights."""
    import torch
    import torchvision
    from torchvision.models.detection._utils import BoxCoder
    rs = np.random.RandomState(42)
    boxes = synth.clustered_boxes(rs, 1500, 600, 800, n_centers=90)
    boxes = boxes[(boxes[:, 2] - boxes[:, 0] > 1) & (boxes[:, 3] - boxes[:, 1] > 1)]
    xyxy = torch.from_numpy(boxes[:, [1, 0, 3, 2]].copy())
    iou_tv = torchvision.ops.box_iou(xyxy[:200], xyxy).numpy()
    np.testing.assert_allclose(ob.bbox_iou(boxes[:200], boxes), iou_tv, rtol=1e-5, atol=1e-6)
    score = synth.tie_free_scores(rs, len(boxes))
    for thresh in (0.3, 0.5, 0.7):
        keep_tv = torchvision.ops.nms(xyxy, torch.from_numpy(score), thresh).numpy()
        keep = ob.non_maximum_suppression(boxes, thresh, score=score)
        np.testing.assert_array_equal(keep, keep_tv)
    # loc2bbox: (dy, dx, dh, dw) on yx boxes == BoxCoder((1,1,1,1)).decode of (dx, dy, dw, dh)
    loc = (rs.standard_normal((len(boxes), 4)) * 0.3).astype(np.float32)
    got = ob.loc2bbox(boxes, loc)
    coder = BoxCoder((1., 1., 1., 1.), bbox_xform_clip=1e9)
    want = coder.decode_single(torch.from_numpy(loc[:, [1, 0, 3, 2]].copy()), xyxy).numpy()
    np.testing.assert_allclose(got[:, [1, 0, 3, 2]], want, rtol=1e-5, atol=1e-3)
    # bbox2loc is its inverse under the same coder
    enc = coder.encode_single(torch.from_numpy(want), xyxy).numpy()
    np.testing.assert_allclose(ob.bbox2loc(boxes, got)[:, [1, 0, 3, 2]], enc, rtol=1e-3, atol=2e-4)
