"""Inference post-processing on the device (csrc/detect.cu) against the oracle's
restatement of MaskRCNN._to_bboxes / _suppress / segm_results, which is itself pinned to
the reference run verbatim (tests/test_oracle_detect.py)."""
import numpy as np
import pytest
import torch

import synth
from chainer_mask_rcnn_b200 import models
from chainer_mask_rcnn_b200.models import mask_rcnn as mr
from oracle import detect as od

pytestmark = pytest.mark.gpu


class _Post(mr.MaskRCNN):
    """Just the post-processing state of a MaskRCNN (no network)."""

    def __init__(self, n_class, detections_per_im=100):
        self._n_class = n_class
        self.loc_normalize_mean = (0., 0., 0., 0.)
        self.loc_normalize_std = (0.1, 0.1, 0.2, 0.2)
        self.nms_thresh, self.score_thresh = 0.5, 0.05
        self._detections_per_im = detections_per_im

    n_class = property(lambda self: self._n_class)


@pytest.mark.parametrize('n_roi,n_class,seed', [(300, 21, 0), (1000, 81, 1), (37, 5, 2)])
def test_to_bboxes_matches_oracle(n_roi, n_class, seed):
    rs = np.random.RandomState(seed)
    sizes, scales = [(300, 400), (280, 390)], np.array([1.6, 1.3], np.float32)
    locs, logits, rois, idx = synth.head_outputs(rs, n_roi, n_class, 2, 300, 400)
    want = od.to_bboxes(locs, logits, rois, idx, sizes, scales, n_class)
    got = _Post(n_class)._to_bboxes(locs, logits, rois, idx, sizes, scales)
    for i in range(2):
        assert len(want[0][i]) > 0
        np.testing.assert_array_equal(got[1][i], want[1][i])             # labels: exact
        np.testing.assert_allclose(got[2][i], want[2][i], rtol=2e-6)     # softmax: fp32 sum order
        np.testing.assert_allclose(got[0][i], want[0][i], rtol=1e-6, atol=1e-4)
        assert got[0][i].dtype == np.float32 and got[1][i].dtype == np.int32


def test_no_cut_and_empty_image():
    rs = np.random.RandomState(3)
    sizes, scales = [(300, 400), (280, 390)], np.array([1.6, 1.6], np.float32)
    locs, logits, rois, idx = synth.head_outputs(rs, 200, 21, 2, 300, 400)
    idx[:] = 0                                            # image 1 has no RoIs at all
    want = od.to_bboxes(locs, logits, rois, idx, sizes, scales, 21, detections_per_im=0)
    got = _Post(21, detections_per_im=0)._to_bboxes(locs, logits, rois, idx, sizes, scales)
    assert len(got[0][1]) == 0 and len(want[0][1]) == 0
    np.testing.assert_array_equal(got[1][0], want[1][0])
    assert len(got[0][0]) > 100


def test_paste_masks_bit_exact():
    rs = np.random.RandomState(4)
    H, W, n, n_fg, M = 300, 400, 40, 20, 14
    b = synth.random_boxes(rs, n, H, W, 4., 350.)
    b[:4] = [[0, 0, H, W], [10.2, 20.7, 10.9, 21.1], [-5, -8, 30, 40], [H - 3, W - 3, H + 9, W + 4]]
    label = rs.randint(0, n_fg, n).astype(np.int32)
    logits = (rs.standard_normal((n, n_fg, M, M)) * 3).astype(np.float32)
    want = od.segm_results(b, label, od.sigmoid(logits), H, W)
    got = mr.segm_results(b, label, od.sigmoid(logits), H, W)
    assert got.dtype == bool and got.shape == want.shape and want.any()
    np.testing.assert_array_equal(got, want)
    # logits + in-kernel sigmoid on a channels-last tensor (what predict feeds)
    dev = torch.device('cuda')
    t = torch.from_numpy(logits).to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    got2 = mr._paste(torch.from_numpy(b).to(dev), torch.from_numpy(label).to(dev), t, H, W, True)
    got2 = got2.cpu().numpy().astype(bool)
    assert (got2 != want).mean() < 1e-5                  # expf vs np.exp at the 0.5 threshold
    assert mr.segm_results(b[:0], label[:0], logits[:0], H, W).shape == (0, H, W)


def test_predict_runs_end_to_end():
    rs = np.random.RandomState(5)
    model = models.MaskRCNNResNet(50, 5, anchor_scales=(4, 8, 16, 32), roi_size=14,
                                  base_channels=32, min_size=160, max_size=240)
    model.score_thresh = 0.0 + 1e-3        # a randomly initialised head: flat probabilities
    imgs = [rs.uniform(0, 255, (3, 120, 150)).astype(np.float32),
            rs.uniform(0, 255, (3, 100, 160)).astype(np.float32)]
    bboxes, masks, labels, scores = model.predict(imgs)
    assert len(bboxes) == len(masks) == len(labels) == len(scores) == 2
    for i, im in enumerate(imgs):
        n = len(bboxes[i])
        assert 0 < n <= 100
        assert masks[i].shape == (n,) + im.shape[1:] and masks[i].dtype == bool
        assert labels[i].shape == (n,) and scores[i].shape == (n,)
        assert (bboxes[i][:, 2] <= im.shape[1]).all() and (bboxes[i][:, 3] <= im.shape[2]).all()
        assert ((labels[i] >= 0) & (labels[i] < 5)).all()


class _Prep(mr.MaskRCNN):
    def __init__(self, min_size, max_size):
        self.min_size, self.max_size = min_size, max_size
        self.mean = np.array([123.152, 115.903, 103.063], np.float32)[:, None, None]


@pytest.mark.parametrize('shapes,min_size,max_size', [
    ([(120, 200), (100, 150)], 200, 400),       # up-scaling, two sizes in one padded batch
    ([(100, 400)], 200, 500),                   # capped by max_size
    ([(160, 203), (161, 322)], 80, 400),        # exact 2x decimation (block mean, cut edge) + plain taps
    ([(96, 128)], 96, 400),                     # identity
    ([(200, 120)], 30, 400),                    # 4x decimation
])
def test_prepare_device_bit_exact(shapes, min_size, max_size):
    """cmr_prepare_image against oracle/prepare.py (itself bit-exact against cv2.resize,
    tests/test_oracle_prepare.py): resize + mean subtraction + zero padding."""
    from oracle import prepare as op
    rs = np.random.RandomState(len(shapes) * 100 + min_size)
    imgs = [rs.uniform(0, 255, (3, h, w)).astype(np.float32) for h, w in shapes]
    m = _Prep(min_size, max_size)
    want, sizes_w, scales_w = op.prepare(imgs, min_size, max_size, m.mean.reshape(-1))
    x, sizes, scales = m._prepare_device(imgs)
    assert sizes == sizes_w and scales == scales_w
    x = x.cpu().numpy()
    Hm, Wm = max(w.shape[1] for w in want), max(w.shape[2] for w in want)
    assert x.shape == (len(imgs), 3, Hm, Wm)
    for i, w in enumerate(want):
        np.testing.assert_array_equal(x[i, :, :w.shape[1], :w.shape[2]], w)
        assert not x[i, :, w.shape[1]:].any() and not x[i, :, :, w.shape[2]:].any()
    # the host-side prepare of the reference API (cv2 as installed: the IPP code path
    # rounds differently) agrees to 4e-5 of the pixel range
    host, _, _ = m.prepare(imgs)
    for i, h in enumerate(host):
        np.testing.assert_allclose(x[i, :, :h.shape[1], :h.shape[2]], h, rtol=0, atol=1e-2)


def test_prepare_golden():
    """Golden vectors of the reference's own MaskRCNN.prepare (tests/golden/prepare.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prepare.npz'))
    for k, (H, W, lo, hi) in enumerate(g['cases']):
        x, sizes, scales = _Prep(int(lo), int(hi))._prepare_device([g['img_%d' % k]])
        assert scales[0] == float(g['scale_%d' % k])
        assert tuple(x.shape[2:]) == g['out_%d' % k].shape[1:]
        np.testing.assert_allclose(x[0].cpu().numpy(), g['out_%d' % k], rtol=0, atol=255e-5)


def test_mask_download_buffers():
    """predict's mask stacks come back as views of page-locked buffers while the budget
    lasts, pageable memory beyond it; either way the bytes are the device's."""
    pool = mr._PinnedDownloads(budget=3000)
    t = torch.arange(2000, dtype=torch.uint8, device='cuda').remainder(2).view(2, 10, 100)
    a = pool.download(t)
    assert pool.outstanding == 2000
    b = pool.download(t)                      # over budget: pageable
    assert pool.outstanding == 2000
    np.testing.assert_array_equal(a, t.cpu().numpy())
    np.testing.assert_array_equal(b, t.cpu().numpy())
    v = a.view(np.bool_)
    del a
    import gc
    gc.collect()
    assert pool.outstanding == 2000 and v[0, 0, 1]     # the view keeps the buffer alive
    del v
    gc.collect()
    assert pool.outstanding == 0
