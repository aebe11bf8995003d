// The five losses of MaskRCNNTrainChain.__call__ (models/mask_rcnn_train_chain.py:
// 160-181) and their gradients with respect to the network outputs, computed on the
// channels-last tensors the convolution kernels produce.  Gradients are written
// (tf32-rounded, they are GEMM operands of the backward pass) into zero-padded
// buffers whose channel layout is what the fused data-gradient GEMMs read.
//
//   smooth L1   _smooth_l1_loss / _fast_rcnn_loc_loss   (:192-213)
//   sigmoid CE  chainer.functions.sigmoid_cross_entropy (normalize=True, -1 = ignore)
//   softmax CE  chainer.functions.softmax_cross_entropy (mean over t != -1)
//
// `losses` is a device float[8]: [0] rpn_loc [1] rpn_cls [2] roi_loc [3] roi_cls
// [4] roi_mask; words [5..7] are int32 normalisation counters (scratch).
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cmr {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Adds the block's total of `v` to *dst (one atomic per block).
__device__ __forceinline__ void block_atomic_add(float v, float* dst) {
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (blockDim.x + 31) / 32 ? part[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0 && t != 0.f) atomicAdd(dst, t);
  }
}

__global__ void __launch_bounds__(256)
count_ge0_kernel(const int* __restrict__ t, size_t n, int* __restrict__ count) {
  int c = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    c += (__ldg(t + i) >= 0) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// d = w * (x - t);  |d| < 1/s2 ? s2/2 d^2 : |d| - 0.5/s2;   dy/dx = w * (...)
__device__ __forceinline__ float smooth_l1(float x, float t, float w, float s2, float* grad) {
  const float d = w * (x - t);
  const float ad = fabsf(d);
  if (ad < 1.0f / s2) {
    *grad = w * s2 * d;
    return 0.5f * s2 * d * d;
  }
  *grad = w * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  return ad - 0.5f / s2;
}

// loss term -(x (t - [x >= 0]) - log1p(exp(-|x|))), gradient sigmoid(x) - t.
__device__ __forceinline__ float sigmoid_ce(float x, int t, float* grad) {
  const float sig = 1.0f / (1.0f + expf(-x));
  *grad = sig - (float)t;
  return -(x * ((float)t - (x >= 0.f ? 1.f : 0.f)) - log1pf(expf(-fabsf(x))));
}

// One thread per anchor (pixel p, anchor a).
__global__ void __launch_bounds__(256)
rpn_loss_kernel(const float* __restrict__ loc, int ld_loc, const float* __restrict__ score,
                int ld_score, const float4* __restrict__ gt_loc, const int* __restrict__ gt_label,
                long long P, int A, float sigma2, float* __restrict__ g, int ld_g,
                float* __restrict__ losses, const int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float l_loc = 0.f, l_cls = 0.f;
  if (i < P * A) {
    const long long p = i / A;
    const int a = (int)(i - p * A);
    const float inv_n = 1.0f / (float)max(*count, 1);
    const int label = __ldg(gt_label + i);
    const float w = label > 0 ? 1.f : 0.f;
    const float4 t = __ldg(gt_loc + i);
    const float* lp = loc + p * ld_loc + a * 4;
    float* gp = g + p * ld_g + a * 4;
    const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr;
      l_loc += smooth_l1(__ldg(lp + k), tt[k], w, sigma2, &gr);
      gp[k] = tc::round_tf32(gr * inv_n);
    }
    float gs = 0.f;
    if (label >= 0) {
      l_cls = sigmoid_ce(__ldg(score + p * ld_score + a), label, &gs);
      gs *= inv_n;
    }
    g[p * ld_g + 4 * A + a] = tc::round_tf32(gs);
    l_loc *= inv_n;
    l_cls *= inv_n;
    if (a == 0)
      for (int c = 5 * A; c < ld_g; ++c) g[p * ld_g + c] = 0.f;
  }
  block_atomic_add(l_loc, losses + 0);
  __syncthreads();
  block_atomic_add(l_cls, losses + 1);
}

// One warp per RoI row.
__global__ void __launch_bounds__(256)
roi_loss_kernel(const float* __restrict__ cls_loc, int ld_cl, const float* __restrict__ score,
                int ld_sc, const float4* __restrict__ gt_loc, const int* __restrict__ gt_label,
                int R, int n_class, float sigma2, float* __restrict__ g, int ld_g,
                float* __restrict__ losses, const int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float l_loc = 0.f, l_cls = 0.f;
  if (r < R) {
    const float inv_n = 1.0f / (float)max(*count, 1);
    const int label = __ldg(gt_label + r);
    float* gr = g + (size_t)r * ld_g;
    for (int c = lane; c < ld_g; c += 32) gr[c] = 0.f;
    __syncwarp();
    if (label >= 0) {
      // localisation: the 4 outputs of the labelled class (weight 1 when label > 0)
      if (lane < 4) {
        const float4 t = __ldg(gt_loc + r);
        const float tt = lane == 0 ? t.x : (lane == 1 ? t.y : (lane == 2 ? t.z : t.w));
        float gl;
        l_loc = smooth_l1(__ldg(cls_loc + (size_t)r * ld_cl + label * 4 + lane), tt,
                          label > 0 ? 1.f : 0.f, sigma2, &gl) * inv_n;
        gr[label * 4 + lane] = tc::round_tf32(gl * inv_n);
      }
      // softmax cross entropy
      const float* sp = score + (size_t)r * ld_sc;
      float m = -INFINITY;
      for (int c = lane; c < n_class; c += 32) m = fmaxf(m, __ldg(sp + c));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f;
      for (int c = lane; c < n_class; c += 32) s += expf(__ldg(sp + c) - m);
      s = warp_sum(s);
      const float logz = m + logf(s);
      for (int c = lane; c < n_class; c += 32) {
        const float logp = __ldg(sp + c) - logz;
        float gv = expf(logp);
        if (c == label) {
          gv -= 1.f;
          l_cls = -logp * inv_n;
        }
        gr[4 * n_class + c] = tc::round_tf32(gv * inv_n);
      }
    }
  }
  block_atomic_add(l_loc, losses + 2);
  __syncthreads();
  block_atomic_add(l_cls, losses + 3);
}

// One thread per channel quad of the (R, HW, ld_g) gradient buffer (ld_g % 4 == 0): a
// 16-byte store each; only the quad holding the RoI's class reads a logit.
__global__ void __launch_bounds__(256)
mask_loss_kernel(const float* __restrict__ masks, int ld_m, const int* __restrict__ gt_label,
                 const int* __restrict__ gt_mask, size_t total4, int HW, int n_fg,
                 float4* __restrict__ g, int ld_g4, float* __restrict__ losses,
                 const int* __restrict__ count) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (e < total4) {
    const int c4 = (int)(e % ld_g4);
    const size_t pix = e / ld_g4;
    const int r = (int)(pix / HW);
    int sel = __ldg(gt_label + r) - 1;  // roi_masks[arange, label - 1]: -1 wraps to the last class
    if (sel < 0) sel += n_fg;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((sel >> 2) == c4) {
      const int t = __ldg(gt_mask + pix);
      if (t >= 0) {
        float gv;
        const float inv_n = 1.0f / (float)max(*count, 1);
        l = sigmoid_ce(__ldg(masks + pix * ld_m + sel), t, &gv) * inv_n;
        gv = tc::round_tf32(gv * inv_n);
        const int k = sel & 3;
        out.x = k == 0 ? gv : 0.f;
        out.y = k == 1 ? gv : 0.f;
        out.z = k == 2 ? gv : 0.f;
        out.w = k == 3 ? gv : 0.f;
      }
    }
    g[e] = out;
  }
  // most blocks hold no selected element at all: skip their reduction
  if (__syncthreads_or(l != 0.f)) block_atomic_add(l, losses + 4);
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_rpn_loss(const float* loc, int ld_loc, const float* score, int ld_score,
                            const float* gt_loc, const int32_t* gt_label, long long n_pixel,
                            int n_anchor, float sigma, float* g, int ld_g, float* losses,
                            void* stream) {
  CMR_REQUIRE(loc && score && gt_loc && gt_label && g && losses);
  CMR_REQUIRE(n_pixel > 0 && n_anchor > 0 && ld_g >= 5 * n_anchor && ld_loc >= 4 * n_anchor &&
              ld_score >= n_anchor);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(gt_loc) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  int* count = reinterpret_cast<int*>(losses + 5);
  CMR_CUDA_TRY(cudaMemsetAsync(losses, 0, 2 * sizeof(float), st));
  CMR_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), st));
  const long long n = n_pixel * n_anchor;
  count_ge0_kernel<<<(unsigned)min((long long)sm_count() * 8, ceil_div_ll(n, 256)), 256, 0, st>>>(
      gt_label, (size_t)n, count);
  CMR_LAUNCH_CHECK();
  rpn_loss_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(
      loc, ld_loc, score, ld_score, reinterpret_cast<const float4*>(gt_loc), gt_label, n_pixel,
      n_anchor, sigma * sigma, g, ld_g, losses, count);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_roi_loss(const float* cls_loc, int ld_cls_loc, const float* score,
                            int ld_score, const float* gt_loc, const int32_t* gt_label, int R,
                            int n_class, float sigma, float* g, int ld_g, float* losses,
                            void* stream) {
  CMR_REQUIRE(cls_loc && score && gt_loc && gt_label && g && losses);
  CMR_REQUIRE(R > 0 && n_class > 0 && ld_g >= 5 * n_class && ld_cls_loc >= 4 * n_class &&
              ld_score >= n_class);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(gt_loc) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  int* count = reinterpret_cast<int*>(losses + 6);
  CMR_CUDA_TRY(cudaMemsetAsync(losses + 2, 0, 2 * sizeof(float), st));
  CMR_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), st));
  count_ge0_kernel<<<ceil_div(R, 256), 256, 0, st>>>(gt_label, (size_t)R, count);
  CMR_LAUNCH_CHECK();
  roi_loss_kernel<<<ceil_div(R, 8), 256, 0, st>>>(cls_loc, ld_cls_loc, score, ld_score,
                                                  reinterpret_cast<const float4*>(gt_loc),
                                                  gt_label, R, n_class, sigma * sigma, g, ld_g,
                                                  losses, count);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_mask_loss(const float* masks, int ld_masks, const int32_t* gt_label,
                             const int32_t* gt_mask, int R, int HW, int n_fg, float* g, int ld_g,
                             float* losses, void* stream) {
  CMR_REQUIRE(masks && gt_label && gt_mask && g && losses && R > 0 && HW > 0 && n_fg > 0);
  CMR_REQUIRE(ld_masks >= n_fg && ld_g >= n_fg && ld_g % 4 == 0);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  int* count = reinterpret_cast<int*>(losses + 7);
  CMR_CUDA_TRY(cudaMemsetAsync(losses + 4, 0, sizeof(float), st));
  CMR_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), st));
  const size_t npix = (size_t)R * HW;
  count_ge0_kernel<<<(unsigned)min((long long)sm_count() * 8, ceil_div_ll(npix, 256)), 256, 0,
                     st>>>(gt_mask, npix, count);
  CMR_LAUNCH_CHECK();
  const size_t total4 = npix * (ld_g / 4);
  mask_loss_kernel<<<(unsigned)ceil_div_ll(total4, 256), 256, 0, st>>>(
      masks, ld_masks, gt_label, gt_mask, total4, HW, n_fg, reinterpret_cast<float4*>(g),
      ld_g / 4, losses, count);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
