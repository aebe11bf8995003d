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

The host-side stages of the reference's step that do not scale with the crop are timed at
FULL size, once per image (``per_image_host_stages``): ProposalCreator on the full
51x84x15 anchor grid (decode, clip, sort, 12000 -> 2000 greedy NMS; the chainercv CPU
algorithm restated in oracle/bbox.py), AnchorTargetCreator and ProposalTargetCreator with its
128 cv2 mask rasterisations (models/mask_rcnn_train_chain.py:126-158).

    seconds per image = t_sample / f + t_proposals + t_anchor_targets + t_proposal_targets
"""
import time

import numpy as np

from . import model as om

f32 = np.float32

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

    def forward_only(self):
        """Forward pass of the sample with 1000/512 of the train sample's RoIs (the box pass
        of inference runs the head on 1000 proposals per image); returns seconds."""
        cfg, p = self.cfg, self.params
        n = max(2, int(round(1000 * self.fraction)))
        self.n_roi_infer = n
        reps = -(-n // self.n_roi)
        rois = np.tile(self.rois, (reps, 1))[:n]
        t0 = time.perf_counter()
        feat, _ = om.extractor(cfg, p, self.x)
        om.rpn_forward(cfg, p, feat)
        om.head_forward(cfg, p, feat, rois, np.zeros((n,), np.int32))
        return time.perf_counter() - t0

    def per_image_host_stages(self, seed=1):
        """Full-size, per-image host stages of the reference's train step -> dict of seconds."""
        from . import bbox as ob
        rs = np.random.RandomState(seed)
        cfg = self.cfg
        fh, fw = self._feat_size(FULL_H), self._feat_size(FULL_W)
        base = ob.generate_anchor_base(cfg.feat_stride, cfg.ratios, cfg.anchor_scales)
        anchor = ob.enumerate_shifted_anchor(base, cfg.feat_stride, fh, fw)
        n = len(anchor)
        loc = (rs.standard_normal((n, 4)) * 0.3).astype(f32)
        score = (rs.permutation(np.linspace(0, 1, n)) * 12 - 6).astype(f32)
        out = {}
        t0 = time.perf_counter()
        pc = ob.ProposalCreator(**cfg.proposal_creator_params)
        roi = pc(loc, score, anchor, (FULL_H, FULL_W), 1.6, train=True)
        out['proposals'] = time.perf_counter() - t0
        # 40 ground-truth boxes with elliptical masks
        hh = np.exp(rs.uniform(np.log(24), np.log(480), 40))
        ww = np.exp(rs.uniform(np.log(24), np.log(480), 40))
        cy, cx = rs.uniform(0, FULL_H, 40), rs.uniform(0, FULL_W, 40)
        bbox = np.stack([np.clip(cy - hh / 2, 0, FULL_H), np.clip(cx - ww / 2, 0, FULL_W),
                         np.clip(cy + hh / 2, 0, FULL_H), np.clip(cx + ww / 2, 0, FULL_W)],
                        1).astype(f32)
        bbox = bbox[(bbox[:, 2] - bbox[:, 0] > 4) & (bbox[:, 3] - bbox[:, 1] > 4)]
        label = rs.randint(0, 80, len(bbox)).astype(np.int32)
        t0 = time.perf_counter()
        ob.AnchorTargetCreator()(bbox, anchor, (FULL_H, FULL_W))
        out['anchor_targets'] = time.perf_counter() - t0
        masks = np.zeros((len(bbox), FULL_H, FULL_W), np.int32)
        for i, (y1, x1, y2, x2) in enumerate(bbox.astype(int)):
            masks[i, y1:y2, x1:x2] = 1
        t0 = time.perf_counter()
        from .mask_target import proposal_targets
        proposal_targets(roi, bbox, label, masks, rs=rs)
        out['proposal_targets'] = time.perf_counter() - t0
        return out

    def images_per_second(self, seconds, extra=None):
        per_image = seconds / self.fraction + (sum(extra.values()) if extra else 0.)
        return 1.0 / per_image

    def describe(self, seconds, extra=None):
        txt = ('%.4f of one image per step: %dx%d crop through backbone+RPN and %d RoIs through '
               'the res5 head, forward+backward, %.1f s; NumPy im2col + BLAS sgemm restatement '
               'of the Chainer CPU path (oracle/model.py), linearly extrapolated'
               % (self.fraction, self.h, self.w, self.n_roi, seconds))
        if extra:
            txt += ('; plus, at full size per image: ProposalCreator 12000->2000 %.2f s, '
                    'AnchorTargetCreator %.2f s, ProposalTargetCreator (128 cv2 mask crops) '
                    '%.2f s' % (extra.get('proposals', 0.), extra.get('anchor_targets', 0.),
                                extra.get('proposal_targets', 0.)))
        return txt


def calibrate_fraction(budget_s, probe_fraction=1. / 256):
    """Pick the sample fraction whose step takes about ``budget_s`` on this host."""
    probe = CpuStepSample(probe_fraction)
    probe.step()
    t = probe.step()
    f = probe.fraction * budget_s / max(t, 1e-3)
    return float(min(max(f, 1. / 512), 1.0))
