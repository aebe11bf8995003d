"""Host-side target creators (product code, NumPy) against the oracle and -- where the
reference tree is present -- against the reference's own ProposalTargetCreator run
verbatim with the same NumPy seed (the sampling draws happen in the same order, so
the selected RoIs, labels, box targets and mask targets must be identical)."""
import os

import numpy as np
import pytest

import synth
from chainer_mask_rcnn_b200.models import utils as mu
from oracle import bbox as ob
from oracle import ref_loader


_scene = synth.detection_scene


def test_box_tools_match_oracle():
    rs = np.random.RandomState(0)
    a = synth.random_boxes(rs, 50, 300, 400)
    b = synth.random_boxes(rs, 7, 300, 400)
    np.testing.assert_array_equal(mu.bbox_iou(a, b), ob.bbox_iou(a, b))
    src, dst = a[:7], b
    loc = mu.bbox2loc(src, dst)
    np.testing.assert_allclose(loc, ob.bbox2loc(src, dst), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mu.loc2bbox(src, loc.astype(np.float32)), dst, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(mu.loc2bbox(src, loc.astype(np.float32)),
                               ob.loc2bbox(src, loc.astype(np.float32)), rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_anchor_target_creator_matches_oracle(seed):
    _, bbox, _, _, (H, W) = _scene(seed)
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (4, 8, 16, 32))
    anchor = ob.enumerate_shifted_anchor(base, 16, H // 16, W // 16)
    want_loc, want_label = ob.AnchorTargetCreator()(bbox, anchor, (H, W),
                                                    rng=np.random.RandomState(seed))
    loc, label = mu.AnchorTargetCreator()(bbox, anchor, (H, W), rng=np.random.RandomState(seed))
    np.testing.assert_array_equal(label, want_label)
    np.testing.assert_allclose(loc, want_loc, rtol=1e-6, atol=1e-6)
    assert label.dtype == np.int32 and loc.dtype == np.float32
    assert (label == 1).sum() <= 128 and (label >= 0).sum() <= 256


@pytest.mark.skipif(not ref_loader.reference_available(), reason='reference tree not present')
@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_proposal_target_creator_matches_reference_verbatim(seed):
    ref = ref_loader.load_proposal_target_creator_module()
    roi, bbox, label, mask, _ = _scene(seed)
    np.random.seed(100 + seed)
    want = ref.ProposalTargetCreator(n_sample=128)(roi, bbox, label, mask)
    np.random.seed(100 + seed)
    got = mu.ProposalTargetCreator(n_sample=128)(roi, bbox, label, mask)
    names = ('sample_roi', 'gt_roi_loc', 'gt_roi_label', 'gt_roi_mask')
    for n, g, w in zip(names, got, want):
        assert g.shape == w.shape, n
        if n == 'gt_roi_loc':
            np.testing.assert_allclose(g, w, rtol=1e-5, atol=1e-6, err_msg=n)
        else:
            np.testing.assert_array_equal(g, w, err_msg=n)
    assert (got[2][:32] > 0).all() and (got[2][32:] == 0).all() or got[2].shape[0] < 128
    assert (got[3][got[2] == 0] == -1).all()


def test_proposal_target_creator_golden(golden_dir):
    """Same check against vectors committed from the verbatim reference run
    (tests/golden/make_golden.py), so that it also holds where /root/reference is absent."""
    path = os.path.join(golden_dir, 'proposal_targets.npz')
    g = np.load(path)
    roi, bbox, label, mask, _ = _scene(int(g['scene_seed']))
    np.random.seed(int(g['np_seed']))
    got = mu.ProposalTargetCreator(n_sample=int(g['n_sample']))(roi, bbox, label, mask)
    np.testing.assert_array_equal(got[0], g['sample_roi'])
    np.testing.assert_allclose(got[1], g['gt_roi_loc'], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(got[2], g['gt_roi_label'])
    np.testing.assert_array_equal(got[3], g['gt_roi_mask'])


def test_empty_bbox_raises_like_reference():
    roi, bbox, label, mask, _ = _scene(0)
    with pytest.raises(ValueError):
        mu.ProposalTargetCreator()(roi, bbox[:0], label[:0], mask[:0])
