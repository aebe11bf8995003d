"""CPU restatement of ProposalTargetCreator's mask-target rasterisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows chainer_mask_rcnn/models/utils/proposal_target_creator.py:163-177: for every
sampled foreground RoI, ``np.round`` the box, crop the assigned instance mask, turn the
crop into one-hot planes, ``cv2.resize`` every plane to (mask_size, mask_size) (default
INTER_LINEAR, float32) and take the argmax over the planes.

``cv2.resize`` is a third-party dependency (opencv-python, unpinned in
requirements.txt).  ``resize_linear_f32`` restates OpenCV's own float32 bilinear code
(modules/imgproc/src/resize.cpp: ``resizeGeneric_`` with ``HResizeLinear`` /
``VResizeLinear``) in NumPy with the same fp32 rounding order:

* ``scale = 1 / (dsize / ssize)`` in double; for destination index d:
  ``f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s``;
* columns: ``s < 0 -> (s, f) = (0, 0)``, ``s >= ssize - 1 -> (s, f) = (ssize - 1, 0)``;
* rows: source rows are clipped to ``[0, ssize - 1]``, the weights are NOT clamped;
* horizontal pass first, ``r = S[x0] * (1 - fx) + S[x1] * fx``, then
  ``out = r[y0] * (1 - fy) + r[y1] * fy``, each product and sum rounded to fp32.

PINNED: bit-exact against ``cv2.resize`` 4.13 with ``cv2.ipp.setUseIPP(False)`` on
thousands of random crops (tests/test_oracle_mask_target.py).  With Intel IPP enabled
(the default of the opencv-python wheel on x86) cv2 takes a different, closed-source
path whose values differ by up to ~2e-5; the argmax then differs on ~1e-4 of the
pixels (those within 2e-5 of a tie).  The golden vectors of the verbatim reference run
(tests/golden/proposal_targets.npz, made with IPP on) are checked with that allowance.
"""
import numpy as np

f32 = np.float32


def _coeffs(ssize, dsize, clamp_weights):
    scale = 1. / (float(dsize) / ssize)
    i0 = np.zeros(dsize, np.int64)
    i1 = np.zeros(dsize, np.int64)
    w = np.zeros((dsize, 2), f32)
    for d in range(dsize):
        f = f32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = f32(f - f32(s))
        if clamp_weights:
            if s < 0:
                s, f = 0, f32(0)
            if s >= ssize - 1:
                s, f = ssize - 1, f32(0)
        i0[d] = min(max(s, 0), ssize - 1)
        i1[d] = min(max(s + 1, 0), ssize - 1)
        w[d, 0] = f32(1.) - f
        w[d, 1] = f
    return i0, i1, w


def resize_linear_f32(src, dh, dw):
    """cv2.resize(src, (dw, dh)) for a 2-D float32 array, INTER_LINEAR."""
    src = np.asarray(src, f32)
    h, w = src.shape
    x0, x1, wx = _coeffs(w, dw, True)
    y0, y1, wy = _coeffs(h, dh, False)
    rows = (src[:, x0] * wx[:, 0][None]).astype(f32) + (src[:, x1] * wx[:, 1][None]).astype(f32)
    rows = rows.astype(f32)
    out = (rows[y0] * wy[:, 0][:, None]).astype(f32) + (rows[y1] * wy[:, 1][:, None]).astype(f32)
    return out.astype(f32)


def roi_mask_target(roi, mask, mask_size=14):
    """One RoI (y1,x1,y2,x2 float32) and its instance mask (H,W) int -> (ms,ms) int32
    (proposal_target_creator.py:166-177).  An empty crop gives zeros (the reference
    raises on ``.max()`` of an empty array; sampled foreground RoIs are never empty)."""
    r = np.round(np.asarray(roi, f32)).astype(np.int32)
    crop = mask[r[0]:r[2], r[1]:r[3]]
    if crop.size == 0:
        return np.zeros((mask_size, mask_size), np.int32)
    planes = [resize_linear_f32((crop == v).astype(f32), mask_size, mask_size)
              for v in range(int(crop.max()) + 1)]
    return np.argmax(np.stack(planes, axis=2), axis=2).astype(np.int32)


def mask_targets(sample_roi, gt_assign, n_pos, masks, mask_size=14):
    """Batched form used by the device kernel's test: sample_roi (B,n,4), gt_assign (B,n),
    n_pos (B,), masks (B,G,H,W) -> (B,n,ms,ms) int32, -1 on rows >= n_pos[b]."""
    B, n, _ = sample_roi.shape
    out = np.full((B, n, mask_size, mask_size), -1, np.int32)
    for b in range(B):
        for j in range(int(n_pos[b])):
            out[b, j] = roi_mask_target(sample_roi[b, j], masks[b][gt_assign[b, j]], mask_size)
    return out


def proposal_targets(roi, bbox, label, mask, n_sample=512, pos_ratio=0.25, pos_iou_thresh=0.5,
                     neg_iou_thresh_hi=0.5, neg_iou_thresh_lo=0.0, mask_size=14,
                     loc_normalize_mean=(0., 0., 0., 0.), loc_normalize_std=(0.1, 0.1, 0.2, 0.2),
                     rs=np.random):
    """ProposalTargetCreator.__call__ (models/utils/proposal_target_creator.py:115-184):
    concat proposals + ground truth, IoU, sample <= round(n_sample * pos_ratio) foreground
    and the rest background RoIs with ``rs.choice`` (the reference uses the global NumPy
    RNG), normalised bbox2loc, and the per-foreground-RoI mask rasterisation.  Used by the
    CPU baseline (oracle/cpu_step.py); the product's host sampler is pinned separately against
    the reference file run verbatim (tests/test_host_targets.py)."""
    from . import bbox as ob
    roi = np.concatenate((np.asarray(roi, f32), np.asarray(bbox, f32)), axis=0)
    pos_per_image = np.round(n_sample * pos_ratio)
    iou = ob.bbox_iou(roi, bbox)
    assign = iou.argmax(axis=1)
    max_iou = iou.max(axis=1)
    gt_label = label[assign] + 1
    pos = np.where(max_iou >= pos_iou_thresh)[0]
    n_pos = int(min(pos_per_image, pos.size))
    if pos.size > 0:
        pos = rs.choice(pos, size=n_pos, replace=False)
    neg = np.where((max_iou < neg_iou_thresh_hi) & (max_iou >= neg_iou_thresh_lo))[0]
    n_neg = int(min(n_sample - n_pos, neg.size))
    if neg.size > 0:
        neg = rs.choice(neg, size=n_neg, replace=False)
    keep = np.append(pos, neg)
    gt_label = gt_label[keep]
    gt_label[n_pos:] = 0
    sample_roi = roi[keep]
    loc = ob.bbox2loc(sample_roi, bbox[assign[keep]])
    loc = (loc - np.array(loc_normalize_mean, f32)) / np.array(loc_normalize_std, f32)
    gt_mask = -np.ones((len(sample_roi), mask_size, mask_size), np.int32)
    for i, p in enumerate(pos):
        gt_mask[i] = roi_mask_target(sample_roi[i], mask[assign[p]], mask_size)
    return sample_roi, loc.astype(f32), gt_label.astype(np.int32), gt_mask
