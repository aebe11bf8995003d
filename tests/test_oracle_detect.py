"""oracle/detect.py against the golden vectors of MaskRCNN._to_bboxes / segm_results run
verbatim from the reference tree (tests/golden/detect.npz, made by make_golden.py)."""
import os

import numpy as np

import synth
from oracle import detect as od


def _inputs(g):
    rs = np.random.RandomState(int(g['seed']))
    n_class = int(g['n_class'])
    locs, logits, rois, idx = synth.head_outputs(rs, int(g['n_roi']), n_class, 2, 300, 400)
    sizes = [tuple(int(v) for v in s) for s in g['sizes']]
    return locs, logits, rois, idx, sizes, g['scales'], n_class


def test_to_bboxes_matches_reference_run(golden_dir):
    g = np.load(os.path.join(golden_dir, 'detect.npz'))
    locs, logits, rois, idx, sizes, scales, n_class = _inputs(g)
    bboxes, labels, scores = od.to_bboxes(locs, logits, rois, idx, sizes, scales, n_class)
    for i in range(2):
        assert len(bboxes[i]) == 100                         # the positional cut keeps D rows
        np.testing.assert_array_equal(bboxes[i], g['bbox_%d' % i])
        np.testing.assert_array_equal(labels[i], g['label_%d' % i])
        np.testing.assert_array_equal(scores[i], g['score_%d' % i])
        assert labels[i].dtype == np.int32 and bboxes[i].dtype == np.float32


def test_segm_results_matches_reference_run(golden_dir):
    g = np.load(os.path.join(golden_dir, 'detect.npz'))
    sizes = [tuple(int(v) for v in s) for s in g['sizes']]
    for i in range(2):
        prob = g['mask_prob_%d' % i]
        n = len(prob)
        got = od.segm_results(g['bbox_%d' % i][:n], g['label_%d' % i][:n], prob, *sizes[i])
        want = np.unpackbits(g['masks_%d' % i], axis=-1)[..., :sizes[i][1]].astype(bool)
        assert got.shape == want.shape and want.any()
        np.testing.assert_array_equal(got, want)


def test_empty_and_quirk():
    assert od.segm_results(np.zeros((0, 4), np.float32), np.zeros((0,), np.int32),
                           np.zeros((0, 3, 14, 14), np.float32), 10, 12).shape == (0, 10, 12)
    # the positional cut: keeps positions k with argsort(score)[k] >= n - D
    score = np.array([0.9, 0.1, 0.5, 0.7], np.float32)
    bbox = np.array([[0, 0, 5, 5]] * 4, np.float32) + np.arange(4, dtype=np.float32)[:, None]
    b, l, s = od.cut_detections(bbox, np.arange(4, dtype=np.int32), score, 2)
    np.testing.assert_array_equal(l, np.flatnonzero(np.argsort(score) >= 2))
