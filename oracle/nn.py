"""NumPy restatement of the Chainer layers on the Mask R-CNN hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: chainer
(requirements.txt:1) is not vendored under the reference tree and cannot be
installed offline, and the reference's tests pin no convolution / pooling / loss
value.  The functions restate Chainer's CPU algorithms in fp32
(convolution = im2col + tensordot -> BLAS sgemm, as chainer.functions.convolution_2d
does on NumPy arrays) in NCHW, the reference's layout, anchored on these call
sites:

  convolution_2d / Convolution2D    models/region_proposal_network.py:75-80,124-131
                                    models/mask_rcnn_resnet.py:141-143,194
  BuildingBlock (BottleneckA/B)     models/mask_rcnn_resnet.py:131-133,181;
                                    models/resnet_extractor.py:76-90
  max_pooling_2d(3, 2, pad=1)       models/resnet_extractor.py:67-69 (cover_all=True)
  average_pooling_2d(7, stride=7)   models/mask_rcnn_resnet.py:187
  Linear                            models/mask_rcnn_resnet.py:134-135,188-190
  Deconvolution2D(2048,256,2,s=2)   models/mask_rcnn_resnet.py:138-139,192-193
  affine_channel_2d                 functions/affine_channel_2d.py:10-56 (in-tree; pinned
                                    by tests/golden/affine_channel.npz)
  sigmoid / softmax cross entropy,  models/mask_rcnn_train_chain.py:165,173,176-178,
  smooth L1                         192-213 (in-tree for smooth L1)

``tests/test_oracle_nn.py`` cross-checks every function here (forward and
backward) against torch CPU fp32 ops / autograd.
"""
import numpy as np

f32 = np.float32


# ------------------------------------------------------------------ conv ----
def _out_size(size, k, s, p, cover_all=False):
    if cover_all:
        return (size + 2 * p - k + s - 1) // s + 1
    return (size + 2 * p - k) // s + 1


def im2col(x, kh, kw, sy, sx, ph, pw, pval=0, cover_all=False):
    n, c, h, w = x.shape
    oh = _out_size(h, kh, sy, ph, cover_all)
    ow = _out_size(w, kw, sx, pw, cover_all)
    img = np.pad(x, ((0, 0), (0, 0), (ph, ph + sy - 1), (pw, pw + sx - 1)),
                 mode='constant', constant_values=(pval,))
    col = np.ndarray((n, c, kh, kw, oh, ow), dtype=x.dtype)
    for j in range(kh):
        jlim = j + sy * oh
        for i in range(kw):
            ilim = i + sx * ow
            col[:, :, j, i, :, :] = img[:, :, j:jlim:sy, i:ilim:sx]
    return col


def col2im(col, sy, sx, ph, pw, h, w):
    n, c, kh, kw, oh, ow = col.shape
    img = np.zeros((n, c, h + 2 * ph + sy - 1, w + 2 * pw + sx - 1), dtype=col.dtype)
    for j in range(kh):
        jlim = j + sy * oh
        for i in range(kw):
            ilim = i + sx * ow
            img[:, :, j:jlim:sy, i:ilim:sx] += col[:, :, j, i]
    return img[:, :, ph:h + ph, pw:w + pw]


def conv2d(x, W, b=None, stride=1, pad=0):
    """x (n,c,h,w), W (o,c,kh,kw) -> (n,o,oh,ow)."""
    kh, kw = W.shape[2:]
    col = im2col(x, kh, kw, stride, stride, pad, pad)
    y = np.tensordot(col, W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)
    if b is not None:
        y += b
    return np.rollaxis(y, 3, 1)


def conv2d_backward(x, W, gy, stride=1, pad=0, need_gx=True):
    """-> (gx, gW, gb)."""
    kh, kw = W.shape[2:]
    n, c, h, w = x.shape
    col = im2col(x, kh, kw, stride, stride, pad, pad)
    gW = np.tensordot(gy, col, ((0, 2, 3), (0, 4, 5))).astype(W.dtype, copy=False)
    gb = gy.sum(axis=(0, 2, 3))
    gx = None
    if need_gx:
        gcol = np.tensordot(W, gy, (0, 1)).astype(x.dtype, copy=False)
        gcol = np.rollaxis(gcol, 3)
        gx = col2im(gcol, stride, stride, pad, pad, h, w)
    return gx, gW, gb


def deconv2d(x, W, b=None, stride=2):
    """Deconvolution2D without padding.  x (n,c,h,w), W (c,o,kh,kw) -> (n,o,oh,ow)."""
    kh, kw = W.shape[2:]
    n, c, h, w = x.shape
    oh, ow = stride * (h - 1) + kh, stride * (w - 1) + kw
    gcol = np.tensordot(W, x, (0, 1)).astype(x.dtype, copy=False)   # (o,kh,kw,n,h,w)
    gcol = np.rollaxis(gcol, 3)
    y = col2im(gcol, stride, stride, 0, 0, oh, ow)
    if b is not None:
        y += b.reshape(1, -1, 1, 1)
    return y


def deconv2d_backward(x, W, gy, stride=2):
    kh, kw = W.shape[2:]
    col = im2col(gy, kh, kw, stride, stride, 0, 0)                  # (n,o,kh,kw,h,w)
    gW = np.tensordot(x, col, ((0, 2, 3), (0, 4, 5))).astype(W.dtype, copy=False)
    gx = np.tensordot(col, W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)
    gx = np.rollaxis(gx, 3, 1)
    gb = gy.sum(axis=(0, 2, 3))
    return gx, gW, gb


def linear(x, W, b=None):
    x = x.reshape(len(x), -1)
    y = x.dot(W.T).astype(x.dtype, copy=False)
    if b is not None:
        y += b
    return y


def linear_backward(x, W, gy):
    x2 = x.reshape(len(x), -1)
    return gy.dot(W).astype(x.dtype).reshape(x.shape), gy.T.dot(x2).astype(W.dtype), gy.sum(0)


# --------------------------------------------------------------- pooling ----
def max_pooling_2d(x, k, stride, pad, cover_all=True):
    col = im2col(x, k, k, stride, stride, pad, pad, pval=-np.inf, cover_all=cover_all)
    n, c, kh, kw, oh, ow = col.shape
    return col.reshape(n, c, kh * kw, oh, ow).max(axis=2)


def average_pooling_2d(x, k, stride):
    col = im2col(x, k, k, stride, stride, 0, 0)
    return col.mean(axis=(2, 3)).astype(x.dtype)


def average_pooling_2d_backward(x_shape, gy, k, stride):
    n, c, h, w = x_shape
    oh, ow = gy.shape[2:]
    gcol = np.tile(gy[:, :, None, None], (1, 1, k, k, 1, 1)).astype(gy.dtype)
    return col2im(gcol, stride, stride, 0, 0, h, w) / f32(k * k)


# ---------------------------------------------------------- element-wise ----
def affine_channel_2d(x, W, b):
    """functions/affine_channel_2d.py:10-21 with W, b of shape (C,)."""
    return W.reshape(1, -1, 1, 1) * x + b.reshape(1, -1, 1, 1)


def affine_channel_2d_backward(x, W, gy):
    """functions/affine_channel_2d.py:40-56: (gx, gW, gb)."""
    gx = W.reshape(1, -1, 1, 1) * gy
    gW = (x * gy).sum(axis=(0, 2, 3))
    gb = gy.sum(axis=(0, 2, 3))
    return gx, gW, gb


def bn_to_affine(gamma, beta, avg_mean, avg_var, eps=1e-5):
    """_get_affine_from_bn (models/resnet_extractor.py:16-29): the AffineChannel2D that
    replaces a BatchNormalization in test mode, W = gamma / sqrt(var + 1e-5),
    b = beta - mean * W, all float32 (pinned: tests/golden/bn_fold.npz holds the outputs of
    the reference function run verbatim)."""
    std = np.sqrt(avg_var.astype(f32) + f32(eps))
    W = (gamma.astype(f32) / std).astype(f32)
    b = (beta.astype(f32) - avg_mean.astype(f32) * W).astype(f32)
    return W, b


def convert_bn_to_affine(params):
    """_convert_bn_to_affine (models/resnet_extractor.py:32-44) on a flat parameter dict:
    every '<link>/{gamma,beta,avg_mean,avg_var}' group (Chainer BatchNormalization; its
    counter 'N' is dropped) becomes '<link>/{W,b}'; everything else is passed through."""
    out = {}
    for key, value in params.items():
        root, _, leaf = key.rpartition('/')
        if leaf in ('beta', 'avg_mean', 'avg_var', 'N') and (root + '/gamma') in params:
            continue
        if leaf == 'gamma':
            out[root + '/W'], out[root + '/b'] = bn_to_affine(
                value, params[root + '/beta'], params[root + '/avg_mean'],
                params[root + '/avg_var'])
        else:
            out[key] = value
    return out


def relu(x):
    return np.maximum(x, 0)


# ----------------------------------------------------------------- losses ---
def sigmoid_cross_entropy(x, t):
    """chainer.functions.sigmoid_cross_entropy(normalize=True): t int32, -1 ignored,
    mean over the non-ignored elements.  Returns (loss, gx)."""
    x = x.astype(f32)
    ignore = (t == -1)
    count = max(int((~ignore).sum()), 1)
    loss = -(~ignore * (x * (t - (x >= 0)) - np.log1p(np.exp(-np.abs(x)))))
    loss = f32(loss.astype(f32).sum() / count)
    sig = (np.tanh(x * f32(0.5)) * f32(0.5) + f32(0.5)).astype(f32)
    gx = ((~ignore) * (sig - t) / f32(count)).astype(f32)
    return loss, gx


def softmax(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(x.dtype)


def softmax_cross_entropy(x, t):
    """Mean over rows with t != -1.  Returns (loss, gx)."""
    x = x.astype(f32)
    m = x.max(axis=1, keepdims=True)
    logz = m + np.log(np.exp(x - m).sum(axis=1, keepdims=True))
    logp = x - logz
    valid = t != -1
    count = max(int(valid.sum()), 1)
    idx = np.where(valid, t, 0)
    loss = f32(-(logp[np.arange(len(t)), idx] * valid).sum() / count)
    g = np.exp(logp)
    g[np.arange(len(t)), idx] -= 1
    g = (g * valid[:, None] / f32(count)).astype(f32)
    return loss, g


def smooth_l1_loss_sum(x, t, in_weight, sigma):
    """models/mask_rcnn_train_chain.py:192-202.  Returns (sum, d sum / dx)."""
    sigma2 = f32(sigma ** 2)
    diff = in_weight * (x - t)
    abs_diff = np.abs(diff)
    flag = (abs_diff < (1. / sigma2)).astype(f32)
    y = flag * (sigma2 / 2.) * np.square(diff) + (1 - flag) * (abs_diff - 0.5 / sigma2)
    g = in_weight * (flag * sigma2 * diff + (1 - flag) * np.sign(diff))
    return f32(y.sum()), g.astype(f32)


def fast_rcnn_loc_loss(pred_loc, gt_loc, gt_label, sigma):
    """models/mask_rcnn_train_chain.py:205-213.  Returns (loss, d loss / d pred_loc)."""
    in_weight = np.zeros_like(gt_loc)
    in_weight[gt_label > 0] = 1
    s, g = smooth_l1_loss_sum(pred_loc, gt_loc, in_weight, sigma)
    n = f32((gt_label >= 0).sum())
    return f32(s / n), (g / n).astype(f32)
