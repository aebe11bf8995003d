// Inference post-processing on the device (sm_100a): the work MaskRCNN.predict does on the
// host between the two head passes and after the second one.
//
//   cmr_detections   MaskRCNN._to_bboxes + _suppress (chainer_mask_rcnn/models/
//                    mask_rcnn.py:178-243): softmax, per-class box decoding (loc2bbox),
//                    clipping, score threshold and per-class NMS for a whole batch.  The
//                    reference loops over 80 classes per image calling a NumPy NMS each
//                    time; here every (image, class) pair is one row of a batch: its
//                    candidates (RoIs whose class probability exceeds the threshold) are
//                    sorted by score in shared memory, and ONE batched NMS launch (mask +
//                    sweep, csrc/nms.cu) resolves all B * (n_class - 1) rows in parallel.
//   cmr_paste_masks  segm_results (mask_rcnn.py:63-107): every detection's 14x14 mask
//                    probability map is padded to 16x16, resized (cv2 INTER_LINEAR, fp32)
//                    to its expanded integer box, thresholded at 0.5 and written into a
//                    (n, H, W) uint8 image stack.
//
// Box arithmetic follows the NumPy expression order in fp32 without FMA contraction (the
// un-normalisation of the offsets runs in fp64 and is rounded once, like the reference's
// float64 mean/std tiles).
#include <math.h>

#include "common.cuh"

namespace cmr {

int launch_sort_desc_u64(unsigned long long* keys, int n_pad, int rows, cudaStream_t st);
int launch_nms_batch(const float* boxes, const int* labels, const int* n_arr, int n_max, int B,
                     float thresh, int32_t* keep, int32_t* n_keep, unsigned long long* mask,
                     cudaStream_t st);

namespace {

__device__ __forceinline__ unsigned int float_to_sortable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sortable_to_float(unsigned int s) {
  return __uint_as_float((s & 0x80000000u) ? (s & 0x7fffffffu) : ~s);
}

// One warp per RoI row: softmax over the class logits, one key per foreground class into
// that class's candidate row.  keys (B, n_class-1, n_pad): sortable(prob) << 32 | roi for
// a probability above the threshold, 0 = no candidate (also every roi >= n_roi[img]).
__global__ void __launch_bounds__(256)
det_score_kernel(const float* __restrict__ score, int ld_score, const int* __restrict__ n_roi,
                 int max_roi, int n_class, int n_pad, float score_thresh,
                 unsigned long long* __restrict__ keys) {
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_pad) return;
  unsigned long long* kcol = keys + (size_t)img * (n_class - 1) * n_pad + r;
  if (r >= max_roi || r >= n_roi[img]) {
    for (int c = 1 + lane; c < n_class; c += 32) kcol[(size_t)(c - 1) * n_pad] = 0ull;
    return;
  }
  const float* x = score + ((size_t)img * max_roi + r) * ld_score;
  float m = -INFINITY;
  for (int c = lane; c < n_class; c += 32) m = fmaxf(m, __ldg(x + c));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < n_class; c += 32) s += expf(__fsub_rn(__ldg(x + c), m));
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int c = lane; c < n_class; c += 32) {
    if (c == 0) continue;
    const float p = __fdiv_rn(expf(__fsub_rn(__ldg(x + c), m)), s);
    kcol[(size_t)(c - 1) * n_pad] =
        p > score_thresh ? ((unsigned long long)float_to_sortable(p) << 32) | (unsigned int)r : 0ull;
  }
}

// Rank k of class c's sorted candidates -> decoded, clipped box of (roi, c) and its score;
// the thread standing on the last non-zero key also writes the row's candidate count.
__global__ void __launch_bounds__(256)
det_gather_kernel(const unsigned long long* __restrict__ keys, int n_pad,
                  const float* __restrict__ cls_loc, int ld_loc, const float4* __restrict__ rois,
                  int max_roi, int n_class, const float* __restrict__ img_info, double4 mean,
                  double4 stdv, float4* __restrict__ box, float* __restrict__ prob,
                  int* __restrict__ n_valid) {
  const int img = blockIdx.z, l = blockIdx.y + 1;
  const int row = img * (n_class - 1) + blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= max_roi) return;
  const unsigned long long* krow = keys + (size_t)row * n_pad;
  const unsigned long long key = krow[k];
  if (key == 0ull) {
    if (k == 0) n_valid[row] = 0;
    return;
  }
  if (k + 1 >= max_roi || krow[k + 1] == 0ull) n_valid[row] = k + 1;
  const int r = (int)(key & 0xffffffffull);
  const float scale = __ldg(img_info + 3 * img), H = __ldg(img_info + 3 * img + 1),
              W = __ldg(img_info + 3 * img + 2);
  const float4 rr = __ldg(rois + (size_t)img * max_roi + r);
  // roi / scale, then loc2bbox(roi, loc * std + mean)
  const float y1 = __fdiv_rn(rr.x, scale), x1 = __fdiv_rn(rr.y, scale);
  const float y2 = __fdiv_rn(rr.z, scale), x2 = __fdiv_rn(rr.w, scale);
  const float* lp = cls_loc + ((size_t)img * max_roi + r) * ld_loc + 4 * l;
  const float dy = (float)((double)__ldg(lp + 0) * stdv.x + mean.x);
  const float dx = (float)((double)__ldg(lp + 1) * stdv.y + mean.y);
  const float dh = (float)((double)__ldg(lp + 2) * stdv.z + mean.z);
  const float dw = (float)((double)__ldg(lp + 3) * stdv.w + mean.w);
  const float h = __fsub_rn(y2, y1), w = __fsub_rn(x2, x1);
  const float cy = __fadd_rn(y1, __fmul_rn(0.5f, h)), cx = __fadd_rn(x1, __fmul_rn(0.5f, w));
  const float ncy = __fadd_rn(__fmul_rn(dy, h), cy), ncx = __fadd_rn(__fmul_rn(dx, w), cx);
  const float nh = __fmul_rn((float)exp((double)dh), h);
  const float nw = __fmul_rn((float)exp((double)dw), w);
  float by1 = __fsub_rn(ncy, __fmul_rn(0.5f, nh)), bx1 = __fsub_rn(ncx, __fmul_rn(0.5f, nw));
  float by2 = __fadd_rn(ncy, __fmul_rn(0.5f, nh)), bx2 = __fadd_rn(ncx, __fmul_rn(0.5f, nw));
  by1 = fminf(fmaxf(by1, 0.f), H);
  by2 = fminf(fmaxf(by2, 0.f), H);
  bx1 = fminf(fmaxf(bx1, 0.f), W);
  bx2 = fminf(fmaxf(bx2, 0.f), W);
  const size_t o = (size_t)row * max_roi + k;
  box[o] = make_float4(by1, bx1, by2, bx2);
  prob[o] = sortable_to_float((unsigned int)(key >> 32));
}

// One CTA per image: the classes' survivor lists are concatenated in class order (the
// order of the reference's per-class loop, mask_rcnn.py:186-201).
constexpr int kEmitThreads = 256;
__global__ void __launch_bounds__(kEmitThreads)
det_emit_kernel(const float4* __restrict__ box, const float* __restrict__ prob,
                const int32_t* __restrict__ keep, const int32_t* __restrict__ n_keep, int max_roi,
                int n_class, int max_cand, float4* __restrict__ det_bbox,
                int32_t* __restrict__ det_label, float* __restrict__ det_score,
                int32_t* __restrict__ n_det) {
  extern __shared__ int offs[];            // n_class entries: exclusive prefix of n_keep
  const int img = blockIdx.x, t = threadIdx.x;
  const int n_fg = n_class - 1;
  if (t == 0) {
    int acc = 0;
    for (int c = 0; c < n_fg; ++c) {
      offs[c] = acc;
      acc += n_keep[img * n_fg + c];
    }
    offs[n_fg] = acc;
    n_det[img] = min(acc, max_cand);
  }
  __syncthreads();
  const int total = min(offs[n_fg], max_cand);
  for (int c = 0; c < n_fg; ++c) {
    const int row = img * n_fg + c;
    const int n = offs[c + 1] - offs[c];
    for (int q = t; q < n; q += kEmitThreads) {
      const int dst = offs[c] + q;
      if (dst >= max_cand) break;
      const size_t s = (size_t)row * max_roi + keep[(size_t)row * max_roi + q];
      const size_t o = (size_t)img * max_cand + dst;
      det_bbox[o] = box[s];
      det_label[o] = c;                  // foreground class ids start at 0 (:196-197)
      det_score[o] = prob[s];
    }
  }
  for (int q = total + t; q < max_cand; q += kEmitThreads) {
    const size_t o = (size_t)img * max_cand + q;
    det_bbox[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    det_label[o] = -1;
    det_score[o] = 0.f;
  }
}

// ------------------------------------------------------------- mask paste ----
// cv2.resize(INTER_LINEAR, fp32) coefficients of destination index d (see targets.cu).
struct Lin {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lin lin_coeff(int d, int ssize, int dsize, bool clamp_weights) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_weights) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
  }
  Lin l;
  l.i0 = min(max(s, 0), ssize - 1);
  l.i1 = min(max(s + 1, 0), ssize - 1);
  l.w0 = __fsub_rn(1.f, f);
  l.w1 = f;
  return l;
}

constexpr int kMaxMask = 30;   // mask_size + 2 <= 32

// grid = (row blocks, detections); a CTA paints rows of one detection's box.
__global__ void __launch_bounds__(256)
paste_masks_kernel(const float4* __restrict__ bbox, const int* __restrict__ label,
                   const float* __restrict__ mask_prob, long long sn, long long sc_, long long sy,
                   long long sx, int ms, int H, int W, int apply_sigmoid,
                   uint8_t* __restrict__ out) {
  __shared__ float padded[(kMaxMask + 2) * (kMaxMask + 2)];
  const int k = blockIdx.y;
  const int P = ms + 2;
  const float4 b = bbox[k];                       // (y1, x1, y2, x2)
  // expand_boxes: fp32 arithmetic on the float32 box (NumPy keeps the array's dtype when
  // it is multiplied by Python floats), truncated toward zero by astype(int32)
  const float sc = (float)(((double)ms + 2.0) / (double)ms);
  const float w_half = __fmul_rn(__fmul_rn(__fsub_rn(b.w, b.y), 0.5f), sc);
  const float h_half = __fmul_rn(__fmul_rn(__fsub_rn(b.z, b.x), 0.5f), sc);
  const float x_c = __fmul_rn(__fadd_rn(b.w, b.y), 0.5f), y_c = __fmul_rn(__fadd_rn(b.z, b.x), 0.5f);
  const int x0 = (int)__fsub_rn(x_c, w_half), x1 = (int)__fadd_rn(x_c, w_half);
  const int y0 = (int)__fsub_rn(y_c, h_half), y1 = (int)__fadd_rn(y_c, h_half);
  const int w = max(x1 - x0 + 1, 1), h = max(y1 - y0 + 1, 1);
  const int xa = max(x0, 0), xb = min(x1 + 1, W), ya = max(y0, 0), yb = min(y1 + 1, H);
  if (xb <= xa || yb <= ya) return;
  const float* src = mask_prob + (long long)k * sn + (long long)label[k] * sc_;
  for (int t = threadIdx.x; t < P * P; t += blockDim.x) {
    const int py = t / P, px = t - py * P;
    float v = 0.f;
    if (py >= 1 && py <= ms && px >= 1 && px <= ms) {
      v = __ldg(src + (py - 1) * sy + (px - 1) * sx);
      if (apply_sigmoid) v = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));   // F.sigmoid, fp32
    }
    padded[t] = v;
  }
  __syncthreads();
  uint8_t* img = out + (size_t)k * H * W;
  const int rows = yb - ya, cols = xb - xa;
  const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  for (int t = threadIdx.x; t < (r_end - r_begin) * cols; t += blockDim.x) {
    const int ry = r_begin + t / cols, rx = t % cols;
    const int dy = ya + ry - y0, dx = xa + rx - x0;     // position inside the resized mask
    const Lin ly = lin_coeff(dy, P, h, false), lx = lin_coeff(dx, P, w, true);
    const float top = __fadd_rn(__fmul_rn(padded[ly.i0 * P + lx.i0], lx.w0),
                                __fmul_rn(padded[ly.i0 * P + lx.i1], lx.w1));
    const float bot = __fadd_rn(__fmul_rn(padded[ly.i1 * P + lx.i0], lx.w0),
                                __fmul_rn(padded[ly.i1 * P + lx.i1], lx.w1));
    const float v = __fadd_rn(__fmul_rn(top, ly.w0), __fmul_rn(bot, ly.w1));
    img[(size_t)(ya + ry) * W + xa + rx] = v > 0.5f ? 1 : 0;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct DetWs {
  size_t keys, n_valid, box, prob, keep, n_keep, mask, total;
  int n_pad, rows;
};
DetWs det_ws(int B, int max_roi, int n_class) {
  DetWs w;
  w.n_pad = next_pow2(max_roi);
  w.rows = B * (n_class - 1);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nb = (max_roi + 63) / 64;
  const size_t rows = w.rows;
  w.keys = take(sizeof(unsigned long long) * rows * w.n_pad);
  w.n_valid = take(sizeof(int) * rows);
  w.box = take(sizeof(float4) * rows * max_roi);
  w.prob = take(sizeof(float) * rows * max_roi);
  w.keep = take(sizeof(int32_t) * rows * max_roi);
  w.n_keep = take(sizeof(int32_t) * rows);
  w.mask = take(sizeof(unsigned long long) * rows * max_roi * nb);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_detections_workspace_bytes(int B, int max_roi, int n_class, int max_cand) {
  (void)max_cand;
  if (B <= 0 || max_roi <= 0 || n_class <= 1) return 256;
  return det_ws(B, max_roi, n_class).total;
}

extern "C" int cmr_detections(const float* cls_loc, int ld_loc, const float* score, int ld_score,
                              const float* rois, const int32_t* n_roi, int B, int max_roi,
                              int n_class, const float* img_info, const double* loc_mean,
                              const double* loc_std, float score_thresh, float nms_thresh,
                              int max_cand, float* det_bbox, int32_t* det_label,
                              float* det_score, int32_t* n_det, void* workspace,
                              size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(cls_loc && score && rois && n_roi && img_info && loc_mean && loc_std);
  CMR_REQUIRE(det_bbox && det_label && det_score && n_det && workspace);
  CMR_REQUIRE(B > 0 && max_roi > 0 && n_class > 1 && max_cand > 0);
  CMR_REQUIRE((long long)B * (n_class - 1) < 65536 && n_class <= 4096);
  CMR_REQUIRE(ld_loc >= 4 * n_class && ld_score >= n_class);
  CMR_REQUIRE(max_roi <= 64 * 2048);
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(rois) | reinterpret_cast<uintptr_t>(det_bbox)) & 15) == 0);
  const DetWs w = det_ws(B, max_roi, n_class);
  if (workspace_bytes < w.total) return CMR_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
  int* n_valid = reinterpret_cast<int*>(ws + w.n_valid);
  float4* box = reinterpret_cast<float4*>(ws + w.box);
  float* prob = reinterpret_cast<float*>(ws + w.prob);
  int32_t* keep = reinterpret_cast<int32_t*>(ws + w.keep);
  int32_t* n_keep = reinterpret_cast<int32_t*>(ws + w.n_keep);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + w.mask);

  det_score_kernel<<<dim3(ceil_div(w.n_pad, 8), B), 256, 0, st>>>(
      score, ld_score, n_roi, max_roi, n_class, w.n_pad, score_thresh, keys);
  CMR_LAUNCH_CHECK();
  // every (image, class) row is sorted on its own: one shared-memory pass for <= 4096 RoIs
  int rc = launch_sort_desc_u64(keys, w.n_pad, w.rows, st);
  if (rc != CMR_OK) return rc;
  const double4 mean = make_double4(loc_mean[0], loc_mean[1], loc_mean[2], loc_mean[3]);
  const double4 stdv = make_double4(loc_std[0], loc_std[1], loc_std[2], loc_std[3]);
  det_gather_kernel<<<dim3(ceil_div(max_roi, 256), n_class - 1, B), 256, 0, st>>>(
      keys, w.n_pad, cls_loc, ld_loc, reinterpret_cast<const float4*>(rois), max_roi, n_class,
      img_info, mean, stdv, box, prob, n_valid);
  CMR_LAUNCH_CHECK();
  // the (image, class) rows are independent NMS problems: one batched launch
  rc = launch_nms_batch(reinterpret_cast<const float*>(box), nullptr, n_valid, max_roi, w.rows,
                        nms_thresh, keep, n_keep, mask, st);
  if (rc != CMR_OK) return rc;
  det_emit_kernel<<<B, kEmitThreads, sizeof(int) * (n_class + 1), st>>>(
      box, prob, keep, n_keep, max_roi, n_class, max_cand, reinterpret_cast<float4*>(det_bbox),
      det_label, det_score, n_det);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_paste_masks(const float* bbox, const int32_t* label, const float* mask_prob,
                               long long stride_n, long long stride_c, long long stride_y,
                               long long stride_x, int n, int mask_size, int H, int W,
                               int apply_sigmoid, uint8_t* out, void* stream) {
  CMR_REQUIRE(n >= 0 && mask_size > 0 && mask_size <= kMaxMask && H > 0 && W > 0);
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(bbox && label && mask_prob && out && n < 65536);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(bbox) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  CMR_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)n * H * W, st));
  paste_masks_kernel<<<dim3(8, n), 256, 0, st>>>(reinterpret_cast<const float4*>(bbox), label,
                                                 mask_prob, stride_n, stride_c, stride_y,
                                                 stride_x, mask_size, H, W, apply_sigmoid, out);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
