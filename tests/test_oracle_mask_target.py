"""oracle/mask_target.py against cv2 (the library the reference calls at
models/utils/proposal_target_creator.py:171-172) and against the golden vectors of the
verbatim reference run (tests/golden/proposal_targets.npz)."""
import os

import cv2
import numpy as np
import pytest

import synth
from oracle import bbox as ob
from oracle import mask_target as omt


@pytest.fixture()
def no_ipp():
    """cv2's own bilinear code path (Intel IPP, when compiled in, replaces it with a
    closed-source one whose values differ in the 5th decimal)."""
    have = hasattr(cv2, 'ipp')
    old = cv2.ipp.useIPP() if have else False
    if have:
        cv2.ipp.setUseIPP(False)
    yield
    if have:
        cv2.ipp.setUseIPP(old)


def _crops(n, seed):
    rs = np.random.RandomState(seed)
    for t in range(n):
        h = rs.randint(1, 40) if t % 2 else rs.randint(1, 500)
        w = rs.randint(1, 40) if t % 4 < 2 else rs.randint(1, 500)
        if t % 5 == 0:
            yield rs.rand(h, w).astype(np.float32)
        else:
            yield (rs.rand(h, w) < 0.5).astype(np.float32)
    for hw in ((28, 28), (14, 14), (1, 1), (7, 28), (56, 56)):      # integer-factor paths
        yield (rs.rand(*hw) < 0.5).astype(np.float32)


def test_resize_is_bit_exact_against_cv2(no_ipp):
    n = 0
    for m in _crops(1500, 1):
        np.testing.assert_array_equal(omt.resize_linear_f32(m, 14, 14), cv2.resize(m, (14, 14)))
        n += 1
    assert n > 1500


def test_resize_with_ipp_is_close():
    worst = 0.
    for m in _crops(300, 2):
        worst = max(worst, float(np.abs(omt.resize_linear_f32(m, 14, 14) -
                                        cv2.resize(m, (14, 14))).max()))
    assert worst <= 1e-4


def test_roi_mask_target_matches_reference_formula(no_ipp):
    """The (at most binary) one-hot / resize / argmax pipeline written exactly as in
    proposal_target_creator.py:166-177, with cv2 doing the resize."""
    roi, bbox, label, mask, _ = synth.detection_scene(3, n_gt=8)
    rs = np.random.RandomState(0)
    for r in roi[:120]:
        g = rs.randint(0, len(bbox))
        m = mask[g] * (1 + (rs.randint(0, 3) == 0))           # some label maps with value 2
        ri = np.round(r).astype(np.int32)
        crop = m[ri[0]:ri[2], ri[1]:ri[3]]
        score = (np.arange(crop.max() + 1) == crop[..., None]).astype(np.float32)
        score = cv2.resize(score, (14, 14))
        if score.ndim == 2:
            score = score.reshape(score.shape[:2] + (1,))
        want = np.argmax(score, axis=2).astype(np.int32)
        np.testing.assert_array_equal(omt.roi_mask_target(r, m, 14), want)


def test_golden_reference_run(golden_dir):
    """Mask targets of the verbatim reference run (made with IPP-enabled cv2): the
    restatement may differ only on pixels within IPP's 2e-5 of a tie."""
    g = np.load(os.path.join(golden_dir, 'proposal_targets.npz'))
    roi, bbox, label, mask, _ = synth.detection_scene(int(g['scene_seed']))
    cand = np.concatenate([roi, bbox])
    iou = ob.bbox_iou(cand, bbox)
    pos = np.flatnonzero(g['gt_roi_label'] > 0)
    assert len(pos) > 0
    diff = total = 0
    for j in pos:
        c = np.flatnonzero((cand == g['sample_roi'][j]).all(1))[0]
        got = omt.roi_mask_target(g['sample_roi'][j], mask[iou[c].argmax()], 14)
        diff += int((got != g['gt_roi_mask'][j]).sum())
        total += got.size
    assert diff <= max(1, total // 1000), (diff, total)
    assert (g['gt_roi_mask'][g['gt_roi_label'] == 0] == -1).all()
