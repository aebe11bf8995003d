"""Generate golden vectors by running the reference's own code (dev container only).

Run from the repo root:  python tests/golden/make_golden.py

The reference files are executed verbatim from /root/reference through
oracle/ref_loader.py (stub ``chainer`` module); nothing is copied.  Outputs are
small .npz fixtures committed next to this script:

  roi_align_unit.npz    reference unit-test fixture, tests/functions_tests/
                        test_roi_align_2d.py:20-39 (np.random.seed(0) first; the
                        reference test is unseeded), sampling_ratio 0, 1, 2
  roi_align_check.npz   the 8x8 hand-typed map of tests/functions_tests/
                        check_roi_align_2d.py:24-47, 3 RoIs, 2x2 output
  roi_align_random.npz  random-normal maps (the unit fixture is a ramp, which
                        hides sampling-ratio errors): feature-stride-16 boxes,
                        7x7 and 14x14 outputs, sampling_ratio 0 and 2
  affine_channel.npz    functions/affine_channel_2d.py forward/backward
  bn_fold.npz           models/resnet_extractor.py:16-44 (_get_affine_from_bn through
                        _convert_bn_to_affine) on a small tree of BatchNormalization stubs
  proposal_targets.npz  models/utils/proposal_target_creator.py on a seeded scene
  detect.npz            MaskRCNN._to_bboxes (+ _suppress) and segm_results of
                        models/mask_rcnn.py on seeded head outputs (tests/synth.py
                        head_outputs); loc2bbox / NMS served by oracle/bbox.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

from oracle import ref_loader  # noqa: E402


def unit_fixture():
    np.random.seed(0)
    N, C = 3, 3
    x = np.arange(N * C * 12 * 8, dtype=np.float32).reshape((N, C, 12, 8))
    np.random.shuffle(x)
    x = 2 * x / x.size - 1
    rois = np.array([[0, 1, 1, 6, 6], [2, 6, 2, 7, 11], [1, 3, 1, 5, 10],
                     [0, 3, 3, 3, 3]], dtype=np.float32)
    outh, outw, scale = 5, 7, 0.6
    gy = np.random.uniform(-1, 1, (4, C, outh, outw)).astype(np.float32)
    out = dict(x=x.astype(np.float32), rois=rois, gy=gy, outh=outh, outw=outw,
               spatial_scale=scale)
    for ratio in (0, 1, 2):
        y = ref_loader.ref_roi_align_forward(x, rois, outh, outw, scale, ratio)
        gx = ref_loader.ref_roi_align_backward(x.shape, rois, gy, outh, outw,
                                               scale, ratio)
        out['y_r%d' % ratio] = y
        out['gx_r%d' % ratio] = gx
    return out


def check_fixture():
    x = np.array([
        [0.88, 0.44, 0.14, 0.16, 0.37, 0.77, 0.96, 0.27],
        [0.19, 0.45, 0.57, 0.16, 0.63, 0.29, 0.71, 0.70],
        [0.66, 0.26, 0.82, 0.64, 0.54, 0.73, 0.59, 0.26],
        [0.85, 0.34, 0.76, 0.84, 0.29, 0.75, 0.62, 0.25],
        [0.32, 0.74, 0.21, 0.39, 0.34, 0.03, 0.33, 0.48],
        [0.20, 0.14, 0.16, 0.13, 0.73, 0.65, 0.96, 0.32],
        [0.19, 0.69, 0.09, 0.86, 0.88, 0.07, 0.01, 0.48],
        [0.83, 0.24, 0.97, 0.04, 0.24, 0.35, 0.50, 0.91],
    ], dtype=np.float32)[None, None]
    rois = np.array([[0, 0, 0, 2, 2], [0, 0, 0, 3, 2], [0, 0, 2, 6, 7]],
                    dtype=np.float32)
    y = ref_loader.ref_roi_align_forward(x, rois, 2, 2, 1.0, 0)
    gy = np.ones_like(y)
    gx = ref_loader.ref_roi_align_backward(x.shape, rois, gy, 2, 2, 1.0, 0)
    return dict(x=x, rois=rois, y=y, gy=gy, gx=gx)


def random_fixture():
    rs = np.random.RandomState(1234)
    N, C, H, W = 2, 3, 13, 17
    x = rs.standard_normal((N, C, H, W)).astype(np.float32)
    img_h, img_w = H * 16, W * 16
    R = 7
    hh = np.exp(rs.uniform(np.log(16), np.log(img_h), R))
    ww = np.exp(rs.uniform(np.log(16), np.log(img_w), R))
    cy = rs.uniform(0, img_h, R)
    cx = rs.uniform(0, img_w, R)
    y1 = np.clip(cy - hh / 2, 0, img_h)
    y2 = np.clip(cy + hh / 2, 0, img_h)
    x1 = np.clip(cx - ww / 2, 0, img_w)
    x2 = np.clip(cx + ww / 2, 0, img_w)
    b = rs.randint(0, N, R)
    rois = np.stack([b, x1, y1, x2, y2], axis=1).astype(np.float32)
    # one RoI covering the whole image (largest adaptive grid, edge clamps)
    rois[0] = [1, 0, 0, img_w, img_h]
    out = dict(x=x, rois=rois, spatial_scale=1. / 16)
    for (oh, ow) in ((7, 7), (14, 14)):
        gy = rs.standard_normal((R, C, oh, ow)).astype(np.float32)
        out['gy_%d' % oh] = gy
        for ratio in (0, 2):
            y = ref_loader.ref_roi_align_forward(x, rois, oh, ow, 1. / 16, ratio)
            gx = ref_loader.ref_roi_align_backward(x.shape, rois, gy, oh, ow,
                                                   1. / 16, ratio)
            out['y_%d_r%d' % (oh, ratio)] = y
            out['gx_%d_r%d' % (oh, ratio)] = gx
    return out


def affine_fixture():
    mod = ref_loader.load_affine_channel_module()
    rs = np.random.RandomState(7)
    x = rs.standard_normal((2, 5, 4, 3)).astype(np.float32)
    W = rs.uniform(0.5, 1.5, (1, 5, 1, 1)).astype(np.float32)
    b = rs.standard_normal((1, 5, 1, 1)).astype(np.float32)
    gy = rs.standard_normal(x.shape).astype(np.float32)
    f = mod.AffineChannel2DFunction()
    y, = f.forward((x, W, b))
    gx, gW, gb = f.backward((x, W, b), (gy,))
    return dict(x=x, W=W, b=b, gy=gy, y=y, gx=gx, gW=gW, gb=gb)


BN_FOLD_LINKS = (('bn1', 64), ('res2/a/bn1', 37), ('res2/a/bn4', 256), ('res3/b2/bn3', 8))


def bn_fold_fixture():
    """The reference's BatchNormalization -> AffineChannel2D conversion run verbatim
    (ref_loader.ref_convert_bn_to_affine); variances from tiny to large so that the
    1e-5 epsilon matters for some channels."""
    rs = np.random.RandomState(3)
    out = {}
    tree = {}
    for path, c in BN_FOLD_LINKS:
        gamma = rs.uniform(0.2, 2.0, c).astype(np.float32)
        beta = rs.standard_normal(c).astype(np.float32)
        mean = (rs.standard_normal(c) * 3).astype(np.float32)
        var = np.exp(rs.uniform(np.log(1e-7), np.log(50.), c)).astype(np.float32)
        tree[path] = (gamma, beta, mean, var)
        for name, a in zip(('gamma', 'beta', 'avg_mean', 'avg_var'), tree[path]):
            out['%s/%s' % (path, name)] = a
    for path, (W, b) in ref_loader.ref_convert_bn_to_affine(tree).items():
        out[path + '/W'], out[path + '/b'] = W, b
    return out


def proposal_target_fixture():
    """The reference's own ProposalTargetCreator (models/utils/proposal_target_creator.py:
    63-184) run verbatim on a seeded scene (tests/synth.py detection_scene)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import synth
    mod = ref_loader.load_proposal_target_creator_module()
    scene_seed, np_seed, n_sample = 5, 1234, 64
    roi, bbox, label, mask, _ = synth.detection_scene(scene_seed)
    np.random.seed(np_seed)
    sr, gl, glab, gm = mod.ProposalTargetCreator(n_sample=n_sample)(roi, bbox, label, mask)
    return dict(scene_seed=scene_seed, np_seed=np_seed, n_sample=n_sample, sample_roi=sr,
                gt_roi_loc=gl.astype(np.float32), gt_roi_label=glab, gt_roi_mask=gm)


def detect_fixture():
    """MaskRCNN._to_bboxes / segm_results (models/mask_rcnn.py:63-107,178-261) verbatim."""
    sys.path.insert(0, os.path.dirname(HERE))
    import cv2
    import synth
    seed, n_roi, n_class = 11, 500, 21
    sizes, scales = [(300, 400), (280, 390)], np.array([1.6, 1.3], np.float32)
    rs = np.random.RandomState(seed)
    locs, logits, rois, idx = synth.head_outputs(rs, n_roi, n_class, 2, 300, 400)
    bboxes, labels, scores = ref_loader.ref_to_bboxes(
        locs, logits, rois, idx, sizes, scales, n_class, (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2),
        0.05, 0.5, 100)
    out = dict(seed=seed, n_roi=n_roi, n_class=n_class, sizes=np.asarray(sizes), scales=scales)
    mod = ref_loader.load_mask_rcnn_module()
    had = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)       # OpenCV's own bilinear code (see oracle/mask_target.py)
    for i in range(2):
        out['bbox_%d' % i], out['label_%d' % i], out['score_%d' % i] = \
            bboxes[i], labels[i], scores[i]
        n = min(len(bboxes[i]), 12)
        prob = (1 / (1 + np.exp(-rs.standard_normal((n, n_class - 1, 14, 14))))).astype(np.float32)
        out['mask_prob_%d' % i] = prob
        out['masks_%d' % i] = np.packbits(mod.segm_results(
            bboxes[i][:n], labels[i][:n], prob, sizes[i][0], sizes[i][1]), axis=-1)
    cv2.ipp.setUseIPP(had)
    return out


PREPARE_CASES = [  # (H, W, min_size, max_size)
    (60, 100, 100, 400),    # 5/3 up
    (50, 200, 100, 250),    # capped by max_size: 1.25
    (80, 103, 40, 400),     # exactly 0.5: OpenCV's 2x2 block mean; 103 * 0.5 rounds up to 52,
                            # so the last column is a block cut by the edge
    (48, 64, 48, 400),      # identity
    (100, 60, 15, 400),     # 0.25: plain bilinear taps, no block mean
]


def prepare_fixture():
    """MaskRCNN.prepare (models/mask_rcnn.py:152-176) verbatim, cv2 as installed (IPP on)."""
    rs = np.random.RandomState(5)
    mean = (123.152, 115.903, 103.063)
    out = dict(mean=np.asarray(mean, np.float32), cases=np.asarray(PREPARE_CASES))
    for k, (H, W, lo, hi) in enumerate(PREPARE_CASES):
        img = rs.uniform(0, 255, (3, H, W)).astype(np.float32)
        prepared, sizes, scales = ref_loader.ref_prepare([img], lo, hi, mean)
        assert sizes[0] == (H, W)
        out['img_%d' % k] = img
        out['out_%d' % k] = prepared[0]
        out['scale_%d' % k] = np.float64(scales[0])
    return out


def main():
    assert ref_loader.reference_available(), 'needs /root/reference'
    np.savez_compressed(os.path.join(HERE, 'roi_align_unit.npz'), **unit_fixture())
    np.savez_compressed(os.path.join(HERE, 'roi_align_check.npz'), **check_fixture())
    np.savez_compressed(os.path.join(HERE, 'roi_align_random.npz'), **random_fixture())
    np.savez_compressed(os.path.join(HERE, 'affine_channel.npz'), **affine_fixture())
    np.savez_compressed(os.path.join(HERE, 'bn_fold.npz'), **bn_fold_fixture())
    np.savez_compressed(os.path.join(HERE, 'proposal_targets.npz'), **proposal_target_fixture())
    np.savez_compressed(os.path.join(HERE, 'detect.npz'), **detect_fixture())
    np.savez_compressed(os.path.join(HERE, 'prepare.npz'), **prepare_fixture())
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
