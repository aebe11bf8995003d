"""Seeded synthetic inputs shared by the oracle-vs-CUDA tests and bench.py
(SURVEY.md 8d).  Pure NumPy; no reference or oracle code."""
import numpy as np


def random_boxes(rs, n, img_h, img_w, lo=16., hi=512.):
    """(n, 4) float32 (y1, x1, y2, x2): log-uniform sizes, uniform centres, clipped."""
    h = np.exp(rs.uniform(np.log(lo), np.log(hi), n))
    w = np.exp(rs.uniform(np.log(lo), np.log(hi), n))
    cy = rs.uniform(0, img_h, n)
    cx = rs.uniform(0, img_w, n)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_h)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_w)
    return b.astype(np.float32)


def clustered_boxes(rs, n, img_h, img_w, n_centers=40):
    """Boxes jittered around a few objects, so that NMS has real work to do."""
    centers = random_boxes(rs, n_centers, img_h, img_w, 32., 400.)
    pick = rs.randint(0, n_centers, n)
    b = centers[pick].astype(np.float64)
    size = np.stack([b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1)
    jitter = rs.normal(0, 0.12, (n, 4)) * np.concatenate([size, size], axis=1)
    b = b + jitter
    y1 = np.minimum(b[:, 0], b[:, 2]); y2 = np.maximum(b[:, 0], b[:, 2])
    x1 = np.minimum(b[:, 1], b[:, 3]); x2 = np.maximum(b[:, 1], b[:, 3])
    b = np.stack([y1, x1, y2, x2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_h)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_w)
    return b.astype(np.float32)


def tie_free_scores(rs, n):
    """A random permutation of linspace(0, 1, n): no two scores are equal."""
    return rs.permutation(np.linspace(0, 1, n)).astype(np.float32)


def rois_xy(rs, n, n_img, img_h, img_w, lo=16., hi=512.):
    """(n, 5) float32 rows (batch_index, x1, y1, x2, y2) in image coordinates."""
    b = random_boxes(rs, n, img_h, img_w, lo, hi)
    idx = rs.randint(0, n_img, n).astype(np.float32)
    return np.stack([idx, b[:, 1], b[:, 0], b[:, 3], b[:, 2]], axis=1).astype(np.float32)


def rpn_outputs(rs, n_anchor, loc_std=0.3):
    """loc (n, 4) ~ N(0, loc_std), tie-free scores (n,)."""
    loc = (rs.standard_normal((n_anchor, 4)) * loc_std).astype(np.float32)
    score = (tie_free_scores(rs, n_anchor) * 12 - 6).astype(np.float32)
    return loc, score


def detection_scene(seed, n_gt=6, H=320, W=416, n_roi=300):
    """A small detection scene: ground-truth boxes with elliptical instance masks and
    candidate RoIs (half jittered copies of the ground truth, half random).
    -> roi (n_roi,4), bbox, label, mask (n_gt,H,W) int32, (H, W)."""
    rs = np.random.RandomState(seed)
    bbox = random_boxes(rs, n_gt, H, W, 40., 200.)
    bbox = bbox[(bbox[:, 2] - bbox[:, 0] > 10) & (bbox[:, 3] - bbox[:, 1] > 10)]
    label = rs.randint(0, 20, len(bbox)).astype(np.int32)
    mask = np.zeros((len(bbox), H, W), np.int32)
    yy, xx = np.mgrid[:H, :W]
    for i, (y1, x1, y2, x2) in enumerate(bbox):
        cy, cx, ry, rx = (y1 + y2) / 2, (x1 + x2) / 2, (y2 - y1) / 2, (x2 - x1) / 2
        mask[i] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).astype(np.int32)
    jitter = bbox[rs.randint(0, len(bbox), n_roi // 2)] + rs.normal(0, 8, (n_roi // 2, 4))
    roi = np.concatenate([jitter, random_boxes(rs, n_roi - n_roi // 2, H, W)])
    roi = np.stack([np.minimum(roi[:, 0], roi[:, 2]), np.minimum(roi[:, 1], roi[:, 3]),
                    np.maximum(roi[:, 0], roi[:, 2]) + 1, np.maximum(roi[:, 1], roi[:, 3]) + 1],
                   axis=1)
    roi[:, 0::2] = np.clip(roi[:, 0::2], 0, H)
    roi[:, 1::2] = np.clip(roi[:, 1::2], 0, W)
    return roi.astype(np.float32), bbox, label, mask, (H, W)


def head_outputs(rs, n_roi, n_class, n_img, img_h, img_w, scale=1.6):
    """Box-head outputs for the inference post-processing: clustered RoIs (in the scaled
    image), per-class offsets ~ N(0, 0.5) (normalised units) and logits in which a few
    object classes stand out per cluster, so that per-class NMS has overlapping boxes.
    -> roi_cls_locs (R,4C), roi_scores (R,C), rois (R,4), roi_indices (R,) sorted."""
    rois = clustered_boxes(rs, n_roi, img_h * scale, img_w * scale, n_centers=12)
    idx = np.sort(rs.randint(0, n_img, n_roi)).astype(np.int32)
    locs = (rs.standard_normal((n_roi, 4 * n_class)) * 0.5).astype(np.float32)
    logits = rs.standard_normal((n_roi, n_class)).astype(np.float32)
    hot = rs.randint(1, n_class, n_roi)
    logits[np.arange(n_roi), hot % 4 + 1] += rs.uniform(2, 6, n_roi).astype(np.float32)
    logits[np.arange(n_roi), hot] += rs.uniform(0, 4, n_roi).astype(np.float32)
    return locs, logits, rois.astype(np.float32), idx
