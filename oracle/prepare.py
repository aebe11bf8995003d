"""CPU restatement of ``MaskRCNN.prepare`` (the host-side image preparation of predict).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows chainer_mask_rcnn/models/mask_rcnn.py:152-176: scale so that the short side is
``min_size`` (capped so that the long side stays <= ``max_size``), ``cv2.resize(img_hwc,
None, fx=scale, fy=scale)`` (INTER_LINEAR, float32), subtract the per-channel mean.

``cv2.resize`` with ``fx, fy`` (opencv-python, unpinned in requirements.txt;
modules/imgproc/src/resize.cpp) differs from the ``dsize`` form restated in
oracle/mask_target.py in three ways that are kept here:

* ``dsize = (cvRound(W * fx), cvRound(H * fy))`` -- round half to even;
* the coordinate scale is ``1 / fx`` (NOT ``ssize / dsize``);
* INTER_LINEAR with an exact 2x decimation in both directions is silently switched to
  INTER_AREA: every destination pixel is ``(a + b + c + d) * 0.25`` of its 2x2 block;
  equal sizes are a plain copy.

PINNED: bit-exact against ``cv2.resize`` 4.13 with ``cv2.ipp.setUseIPP(False)`` for
random float images and scales (tests/test_oracle_prepare.py); with IPP enabled cv2
differs in the 5th significant digit, the golden vectors (tests/golden/prepare.npz, made
with IPP on through the reference's own expression) are checked with rtol 1e-5.
"""
import numpy as np

f32 = np.float32


def cv_round(v):
    """cvRound: nearest integer, ties to even (lrint)."""
    return int(np.rint(v))


def _coeffs(ssize, dsize, scale, clamp_weights):
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


def out_size(H, W, fx, fy):
    return cv_round(H * fy), cv_round(W * fx)


def _area2(img, dh, dw):
    """resizeAreaFast_ with a 2x2 block (the scalar path OpenCV takes for 3 channels):
    full blocks ((a + b) + c) + d times 0.25f; blocks cut by the right/bottom edge (odd
    source size, destination rounded up) average the pixels that exist, sum / count."""
    C, H, W = img.shape
    out = np.zeros((C, dh, dw), f32)
    fh, fw = min(dh, H // 2), min(dw, W // 2)
    a = img[:, 0:2 * fh:2, 0:2 * fw:2]
    b = img[:, 0:2 * fh:2, 1:2 * fw:2]
    c = img[:, 1:2 * fh:2, 0:2 * fw:2]
    d = img[:, 1:2 * fh:2, 1:2 * fw:2]
    out[:, :fh, :fw] = (((a + b).astype(f32) + c).astype(f32) + d).astype(f32) * f32(0.25)
    for dy in range(dh):
        for dx in range(dw):
            if dy < fh and dx < fw:
                continue
            if 2 * dy >= H or 2 * dx >= W:
                continue                      # stays 0
            acc, count = np.zeros(C, f32), 0
            for sy in range(2):
                if 2 * dy + sy >= H:
                    break
                for sx in range(2):
                    if 2 * dx + sx >= W:
                        break
                    acc = (acc + img[:, 2 * dy + sy, 2 * dx + sx]).astype(f32)
                    count += 1
            out[:, dy, dx] = acc / f32(count)
    return out


def resize_fxfy(img, fx, fy):
    """cv2.resize(img_hwc, None, fx=fx, fy=fy) on a (C, H, W) float32 array -> (C, h, w)."""
    img = np.asarray(img, f32)
    C, H, W = img.shape
    dh, dw = out_size(H, W, fx, fy)
    if (dh, dw) == (H, W):
        return img.copy()
    sx, sy = 1. / fx, 1. / fy
    eps = np.finfo(np.float64).eps
    if abs(sx - 2) < eps and abs(sy - 2) < eps:
        return _area2(img, dh, dw)
    x0, x1, wx = _coeffs(W, dw, sx, True)
    y0, y1, wy = _coeffs(H, dh, sy, False)
    rows = ((img[:, :, x0] * wx[:, 0]).astype(f32) + (img[:, :, x1] * wx[:, 1]).astype(f32)).astype(f32)
    out = (rows[:, y0] * wy[:, 0][None, :, None]).astype(f32) + \
        (rows[:, y1] * wy[:, 1][None, :, None]).astype(f32)
    return out.astype(f32)


def prepare_scale(H, W, min_size, max_size):
    """(:158-165) Python-float arithmetic."""
    scale = 1.
    if min_size:
        scale = min_size / min(H, W)
    if max_size and scale * max(H, W) > max_size:
        scale = max_size / max(H, W)
    return scale


def prepare(imgs, min_size, max_size, mean):
    """-> prepared (list of (3, h, w) float32), sizes, scales."""
    mean = np.asarray(mean, f32).reshape(3, 1, 1)
    prepared, sizes, scales = [], [], []
    for img in imgs:
        _, H, W = img.shape
        scale = prepare_scale(H, W, min_size, max_size)
        out = resize_fxfy(img, scale, scale)
        prepared.append((out - mean).astype(f32, copy=False))
        sizes.append((H, W))
        scales.append(scale)
    return prepared, sizes, scales
