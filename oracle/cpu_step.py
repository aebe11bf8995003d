"""Bounded CPU sample of the R50-C4 train step, timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py): this is the
``cpu_baseline`` / ``--impl reference`` leg of bench.py -- the reference's
Chainer/NumPy CPU path restated in oracle/model.py (im2col + BLAS sgemm
convolutions, Python-loop ROIAlign vectorised per RoI), since chainer itself cannot
be installed here.

A sample is a fraction ``f`` of ONE image's work with the full-width network and
the step's own layer mix: an (h, w) crop with h*w = f * 800*1333 pixels through the
backbone and RPN (forward + backward) and round(512 * f) sampled RoIs through the
res5 head (forward + backward), i.e. ``f`` of the per-image FLOPs of every stage.
images/s = f / seconds.
"""
import time

import numpy as np

from . import model as om

FULL_H, FULL_W, ROIS_PER_IMAGE = 800, 1333, 512


def sample_shape(fraction):
    h = max(64, int(round(FULL_H * np.sqrt(fraction) / 16)) * 16)
    w = max(64, int(round(FULL_W * np.sqrt(fraction) / 16)) * 16)
    n_roi = max(2, int(round(ROIS_PER_IMAGE * fraction)))
    return h, w, n_roi


class CpuStepSample(object):

    def __init__(self, fraction, n_layers=50, seed=0):
        rs = np.random.RandomState(seed)
        self.cfg = om.Config(n_layers=n_layers, n_fg_class=80, anchor_scales=(2, 4, 8, 16, 32),
                             roi_size=14, base=64)
        self.params = om.make_params(self.cfg, rs)
        h, w, n_roi = sample_shape(fraction)
        self.h, self.w, self.n_roi = h, w, n_roi
        self.fraction = (h * w) / float(FULL_H * FULL_W)
        self.x = (rs.uniform(0, 255, (1, 3, h, w)) - 115.).astype(np.float32)
        y1 = rs.uniform(0, h * 0.6, n_roi); x1 = rs.uniform(0, w * 0.6, n_roi)
        hs = np.exp(rs.uniform(np.log(24), np.log(max(h * 0.4, 32)), n_roi))
        ws = np.exp(rs.uniform(np.log(24), np.log(max(w * 0.4, 32)), n_roi))
        self.rois = np.stack([y1, x1, np.minimum(y1 + hs, h), np.minimum(x1 + ws, w)],
                             axis=1).astype(np.float32)
        self.idx = np.zeros((n_roi,), np.int32)
        self.gt_roi_locs = (rs.standard_normal((n_roi, 4)) * 0.5).astype(np.float32)
        self.gt_roi_labels = rs.randint(0, 81, n_roi).astype(np.int32)
        self.gt_roi_masks = rs.randint(0, 2, (n_roi, 14, 14)).astype(np.int32)
        self.gt_roi_masks[self.gt_roi_labels == 0] = -1
        self._rs = rs
        self._rpn_targets = None

    def step(self):
        """One forward + backward of the sample; returns seconds."""
        t0 = time.perf_counter()
        cfg, p = self.cfg, self.params
        if self._rpn_targets is None:
            fh = self._feat_size(self.h)
            fw = self._feat_size(self.w)
            n_anchor = fh * fw * cfg.n_anchor
            self._rpn_targets = (
                (self._rs.standard_normal((n_anchor, 4)) * 0.3).astype(np.float32),
                self._rs.choice([-1, 0, 1], size=n_anchor, p=[0.9, 0.06, 0.04]).astype(np.int32))
            t0 = time.perf_counter()
        gl, glab = self._rpn_targets
        om.train_step_grads(cfg, p, self.x, self.rois, self.idx, self.gt_roi_locs,
                            self.gt_roi_labels, self.gt_roi_masks, gl, glab)
        return time.perf_counter() - t0

    @staticmethod
    def _feat_size(s):
        s = (s + 2 * 3 - 7) // 2 + 1                 # conv1
        s = (s + 2 * 1 - 3 + 2 - 1) // 2 + 1         # max pool, cover_all
        s = (s - 1) // 2 + 1                         # res3
        return (s - 1) // 2 + 1                      # res4

    def images_per_second(self, seconds):
        return self.fraction / seconds


def calibrate_fraction(budget_s, probe_fraction=1. / 256):
    """Pick the sample fraction whose step takes about ``budget_s`` on this host."""
    probe = CpuStepSample(probe_fraction)
    probe.step()
    t = probe.step()
    f = probe.fraction * budget_s / max(t, 1e-3)
    return float(min(max(f, 1. / 512), 1.0))
