// Image preparation of MaskRCNN.predict on the device
// (chainer_mask_rcnn/models/mask_rcnn.py:152-176): cv2.resize(img, None, fx, fy)
// (INTER_LINEAR, float32) followed by the per-channel mean subtraction, written straight
// into the zero-padded (3, out_h, out_w) planes the extractor consumes.
//
// The arithmetic restates OpenCV's float path operation by operation (fp32 products and
// sums, no FMA contraction; see oracle/prepare.py for the three fx/fy-specific rules: the
// destination size is cvRound(size * f), the coordinate scale is 1 / f, and an exact 2x
// decimation is computed as the 2x2 block mean).  HBM-bound: 4 * 3 * (H*W + out_h*out_w)
// bytes per image.
#include <math.h>

#include "common.cuh"

namespace cmr {
namespace {

enum { kModeCopy = 0, kModeArea2 = 1, kModeLinear = 2 };

struct PrepParams {
  const float* img;
  float* out;
  int H, W, h, w, out_h, out_w, mode;
  double sx, sy;
  float mean[3];
};

__device__ __forceinline__ void lin(int d, int ssize, double scale, bool clamp_weights, int& i0,
                                    int& i1, float& w0, float& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_weights) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
  }
  i0 = min(max(s, 0), ssize - 1);
  i1 = min(max(s + 1, 0), ssize - 1);
  w0 = __fsub_rn(1.f, f);
  w1 = f;
}

__global__ void __launch_bounds__(256) prepare_image_kernel(const PrepParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.out_w) return;
  const size_t splane = (size_t)p.H * p.W, dplane = (size_t)p.out_h * p.out_w;
  float* o = p.out + (size_t)y * p.out_w + x;
  if (y >= p.h || x >= p.w) {          // concat_examples(padding=0)
    o[0] = 0.f; o[dplane] = 0.f; o[2 * dplane] = 0.f;
    return;
  }
  if (p.mode == kModeCopy) {
    const float* s = p.img + (size_t)y * p.W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * dplane] = __fsub_rn(__ldg(s + c * splane), p.mean[c]);
    return;
  }
  if (p.mode == kModeArea2) {
    const int sy = 2 * y, sx = 2 * x;
    const bool full = sy + 1 < p.H && sx + 1 < p.W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* s = p.img + c * splane;
      float v = 0.f;
      if (full) {
        const float a = __ldg(s + (size_t)sy * p.W + sx), b = __ldg(s + (size_t)sy * p.W + sx + 1);
        const float cc = __ldg(s + (size_t)(sy + 1) * p.W + sx);
        const float d = __ldg(s + (size_t)(sy + 1) * p.W + sx + 1);
        v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), cc), d), 0.25f);
      } else if (sy < p.H && sx < p.W) {   // block cut by the right / bottom edge
        float acc = 0.f;
        int count = 0;
        for (int dy = 0; dy < 2 && sy + dy < p.H; ++dy)
          for (int dx = 0; dx < 2 && sx + dx < p.W; ++dx) {
            acc = __fadd_rn(acc, __ldg(s + (size_t)(sy + dy) * p.W + sx + dx));
            ++count;
          }
        v = __fdiv_rn(acc, (float)count);
      }
      o[c * dplane] = __fsub_rn(v, p.mean[c]);
    }
    return;
  }
  int x0, x1, y0, y1;
  float wx0, wx1, wy0, wy1;
  lin(x, p.W, p.sx, true, x0, x1, wx0, wx1);
  lin(y, p.H, p.sy, false, y0, y1, wy0, wy1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* s = p.img + c * splane;
    const float top = __fadd_rn(__fmul_rn(__ldg(s + (size_t)y0 * p.W + x0), wx0),
                                __fmul_rn(__ldg(s + (size_t)y0 * p.W + x1), wx1));
    const float bot = __fadd_rn(__fmul_rn(__ldg(s + (size_t)y1 * p.W + x0), wx0),
                                __fmul_rn(__ldg(s + (size_t)y1 * p.W + x1), wx1));
    const float v = __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, wy1));
    o[c * dplane] = __fsub_rn(v, p.mean[c]);
  }
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_prepare_size(int H, int W, double fx, double fy, int* h, int* w) {
  CMR_REQUIRE(H > 0 && W > 0 && fx > 0 && fy > 0 && h && w);
  // cvRound: nearest, ties to even (the default rounding mode of nearbyint)
  const double dh = nearbyint((double)H * fy), dw = nearbyint((double)W * fx);
  CMR_REQUIRE(dh >= 1 && dw >= 1 && dh < 1e9 && dw < 1e9);
  *h = (int)dh;
  *w = (int)dw;
  return CMR_OK;
}

extern "C" int cmr_prepare_image(const float* img, int H, int W, double fx, double fy,
                                 float mean0, float mean1, float mean2, float* out, int out_h,
                                 int out_w, void* stream) {
  CMR_REQUIRE(img && out);
  PrepParams p;
  int st = cmr_prepare_size(H, W, fx, fy, &p.h, &p.w);
  if (st != CMR_OK) return st;
  CMR_REQUIRE(out_h >= p.h && out_w >= p.w && out_h < 65536);
  p.img = img; p.out = out;
  p.H = H; p.W = W; p.out_h = out_h; p.out_w = out_w;
  p.sx = 1.0 / fx; p.sy = 1.0 / fy;
  p.mean[0] = mean0; p.mean[1] = mean1; p.mean[2] = mean2;
  const double eps = 2.220446049250313e-16;
  if (p.h == H && p.w == W) p.mode = kModeCopy;
  else if (fabs(p.sx - 2.0) < eps && fabs(p.sy - 2.0) < eps) p.mode = kModeArea2;
  else p.mode = kModeLinear;
  dim3 grid((out_w + 255) / 256, out_h);
  prepare_image_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
