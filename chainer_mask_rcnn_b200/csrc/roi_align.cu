// ROIAlign forward / backward for sm_100a.
//
// Semantics: chainer_mask_rcnn/functions/roi_align_2d.py:179-284 (forward) and
// :405-518 (backward): no "aligned" half-pixel shift, malformed RoIs forced to
// 1x1, adaptive sampling grid ceil(roi/pooled) when sampling_ratio == 0, samples
// outside [-1,H]x[-1,W] skipped but still counted in the divisor.
//
// Two layouts:
//  * NCHW  (the reference operator's layout; drop-in `functions.roi_align_2d`).
//    One CTA = one RoI x a chunk of kChanPerCta channels.  A thread owns one
//    output bin (ph,pw), computes the sample geometry once and reuses it for
//    every channel of the chunk (4*kChanPerCta independent loads in flight per
//    sample), then writes bins of one (roi,channel) plane contiguously.
//  * NHWC  (what the model uses internally).  One CTA = one RoI x one output
//    bin; lanes run over channels as float4, so every tap is a fully coalesced
//    512 B warp load and the geometry is CTA-uniform.  `bin_stride` produces only
//    every bin_stride-th bin (res5.a reads the 14x14 pool with stride 2).
#include "common.cuh"

namespace cmr {
namespace {

struct RoiGeom {
  int batch;
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w;
  float inv_count_den;  // count = grid_h * grid_w (as float)
};

// roi_align_2d.py:184-211.  All fp32, IEEE ops without contraction so the
// truncations / comparisons below see the same values as the reference.
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi,
                                                float scale, int outh, int outw,
                                                int sampling_ratio) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  float end_w = __fmul_rn(roi[3], scale);
  float end_h = __fmul_rn(roi[4], scale);
  float roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.0f);
  float roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.0f);
  g.bin_h = __fdiv_rn(roi_h, (float)outh);
  g.bin_w = __fdiv_rn(roi_w, (float)outw);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_h, (float)outh));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_w, (float)outw));
  g.inv_count_den = (float)(g.grid_h * g.grid_w);
  return g;
}

struct AxisTap {
  int low, high;
  float l, h;
  bool valid;
};

// roi_align_2d.py:216-262 for one axis.
__device__ __forceinline__ AxisTap axis_tap(float start, float bin, int p, int i,
                                            int grid, int limit) {
  AxisTap t;
  float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
  t.valid = !(c < -1.0f || c > (float)limit);
  if (c <= 0.0f) c = 0.0f;
  int low = (int)c;
  if (low >= limit - 1) {
    low = limit - 1;
    t.high = low;
    c = (float)low;
  } else {
    t.high = low + 1;
  }
  t.low = low;
  t.l = __fsub_rn(c, (float)low);
  t.h = __fsub_rn(1.0f, t.l);
  return t;
}

// ------------------------------------------------------------------ NCHW --
constexpr int kChanPerCta = 8;

template <int CB>
__global__ void __launch_bounds__(256)
roi_align_nchw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ rois,
                          float* __restrict__ y, int C, int H, int W, int outh, int outw,
                          float scale, int sampling_ratio, int chunks_per_roi) {
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const int P = outh * outw;
  const size_t HW = (size_t)H * W;
  const float* __restrict__ plane0 = x + ((size_t)g.batch * C + c0) * HW;
  float* __restrict__ out0 = y + ((size_t)r * C + c0) * P;
  const int nch = min(CB, C - c0);

  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int ph = p / outw;
    const int pw = p - ph * outw;
    float acc[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
        if (!(ty.valid && tx.valid)) continue;
        const float w1 = ty.h * tx.h, w2 = ty.h * tx.l, w3 = ty.l * tx.h, w4 = ty.l * tx.l;
        const int o1 = ty.low * W + tx.low, o2 = ty.low * W + tx.high;
        const int o3 = ty.high * W + tx.low, o4 = ty.high * W + tx.high;
        if (nch == CB) {
          float v1[CB], v2[CB], v3[CB], v4[CB];
#pragma unroll
          for (int k = 0; k < CB; ++k) {
            const float* pl = plane0 + (size_t)k * HW;
            v1[k] = __ldg(pl + o1);
            v2[k] = __ldg(pl + o2);
            v3[k] = __ldg(pl + o3);
            v4[k] = __ldg(pl + o4);
          }
#pragma unroll
          for (int k = 0; k < CB; ++k)
            acc[k] += w1 * v1[k] + w2 * v2[k] + w3 * v3[k] + w4 * v4[k];
        } else {
          for (int k = 0; k < nch; ++k) {
            const float* pl = plane0 + (size_t)k * HW;
            acc[k] += w1 * __ldg(pl + o1) + w2 * __ldg(pl + o2) + w3 * __ldg(pl + o3) +
                      w4 * __ldg(pl + o4);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (k < nch) out0[(size_t)k * P + p] = __fdiv_rn(acc[k], g.inv_count_den);
  }
}

template <int CB>
__global__ void __launch_bounds__(256)
roi_align_nchw_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ rois,
                          float* __restrict__ gx, int C, int H, int W, int outh, int outw,
                          float scale, int sampling_ratio, int chunks_per_roi) {
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const int P = outh * outw;
  const size_t HW = (size_t)H * W;
  float* __restrict__ plane0 = gx + ((size_t)g.batch * C + c0) * HW;
  const float* __restrict__ in0 = gy + ((size_t)r * C + c0) * P;
  const int nch = min(CB, C - c0);

  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int ph = p / outw;
    const int pw = p - ph * outw;
    float gval[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) gval[k] = (k < nch) ? __ldg(in0 + (size_t)k * P + p) : 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
        if (!(ty.valid && tx.valid)) continue;
        const float w1 = ty.h * tx.h, w2 = ty.h * tx.l, w3 = ty.l * tx.h, w4 = ty.l * tx.l;
        const int o1 = ty.low * W + tx.low, o2 = ty.low * W + tx.high;
        const int o3 = ty.high * W + tx.low, o4 = ty.high * W + tx.high;
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          if (k < nch) {
            float* pl = plane0 + (size_t)k * HW;
            // g_k = top_diff * w_k / count  (roi_align_2d.py:501-504)
            atomicAdd(pl + o1, __fdiv_rn(gval[k] * w1, g.inv_count_den));
            atomicAdd(pl + o2, __fdiv_rn(gval[k] * w2, g.inv_count_den));
            atomicAdd(pl + o3, __fdiv_rn(gval[k] * w3, g.inv_count_den));
            atomicAdd(pl + o4, __fdiv_rn(gval[k] * w4, g.inv_count_den));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ NHWC --
// grid = (R * oh_s * ow_s), block = C/4 threads (<= 1024) looping if C/4 larger.
__global__ void __launch_bounds__(256)
roi_align_nhwc_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ rois,
                          float4* __restrict__ y, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale,
                          int sampling_ratio) {
  const int P = oh_s * ow_s;
  const int r = blockIdx.x / P;
  const int p = blockIdx.x - r * P;
  const int ph = (p / ow_s) * bin_stride;
  const int pw = (p % ow_s) * bin_stride;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const float4* __restrict__ img = x + (size_t)g.batch * H * W * C4;
  float4* __restrict__ out = y + (size_t)blockIdx.x * C4;

  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
        if (!(ty.valid && tx.valid)) continue;
        const float w1 = ty.h * tx.h, w2 = ty.h * tx.l, w3 = ty.l * tx.h, w4 = ty.l * tx.l;
        const float4 v1 = __ldg(img + (size_t)(ty.low * W + tx.low) * C4 + c);
        const float4 v2 = __ldg(img + (size_t)(ty.low * W + tx.high) * C4 + c);
        const float4 v3 = __ldg(img + (size_t)(ty.high * W + tx.low) * C4 + c);
        const float4 v4 = __ldg(img + (size_t)(ty.high * W + tx.high) * C4 + c);
        acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
        acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
        acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
        acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
      }
    }
    const float d = g.inv_count_den;
    out[c] = make_float4(__fdiv_rn(acc.x, d), __fdiv_rn(acc.y, d), __fdiv_rn(acc.z, d),
                         __fdiv_rn(acc.w, d));
  }
}

__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
  // sm_90+: 128-bit vector reduction to global memory.
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256)
roi_align_nhwc_bwd_kernel(const float4* __restrict__ gy, const float* __restrict__ rois,
                          float4* __restrict__ gx, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale,
                          int sampling_ratio) {
  const int P = oh_s * ow_s;
  const int r = blockIdx.x / P;
  const int p = blockIdx.x - r * P;
  const int ph = (p / ow_s) * bin_stride;
  const int pw = (p % ow_s) * bin_stride;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  float4* __restrict__ img = gx + (size_t)g.batch * H * W * C4;
  const float4* __restrict__ in = gy + (size_t)blockIdx.x * C4;
  const float d = g.inv_count_den;

  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    const float4 gv = __ldg(in + c);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
        if (!(ty.valid && tx.valid)) continue;
        const float w[4] = {ty.h * tx.h, ty.h * tx.l, ty.l * tx.h, ty.l * tx.l};
        const int o[4] = {ty.low * W + tx.low, ty.low * W + tx.high, ty.high * W + tx.low,
                          ty.high * W + tx.high};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 t = make_float4(__fdiv_rn(gv.x * w[k], d), __fdiv_rn(gv.y * w[k], d),
                                 __fdiv_rn(gv.z * w[k], d), __fdiv_rn(gv.w * w[k], d));
          red_add_f4(img + (size_t)o[k] * C4 + c, t);
        }
      }
    }
  }
}

int pick_threads(int positions) {
  int t = ((positions + 31) / 32) * 32;
  if (t < 64) t = 64;
  if (t > 256) t = 256;
  return t;
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_roi_align_fwd(const float* x, int N, int C, int H, int W, const float* rois,
                                 int R, int outh, int outw, float spatial_scale,
                                 int sampling_ratio, float* y, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x && rois && y);
  const int chunks = ceil_div(C, kChanPerCta);
  CMR_REQUIRE((long long)R * chunks < (1ll << 31));
  roi_align_nchw_fwd_kernel<kChanPerCta>
      <<<R * chunks, pick_threads(outh * outw), 0, as_stream(stream)>>>(
          x, rois, y, C, H, W, outh, outw, spatial_scale, sampling_ratio, chunks);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_roi_align_bwd(const float* gy, const float* rois, int R, int N, int C, int H,
                                 int W, int outh, int outw, float spatial_scale,
                                 int sampling_ratio, float* gx, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && gx);
  CMR_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)N * C * H * W, as_stream(stream)));
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(gy && rois);
  const int chunks = ceil_div(C, kChanPerCta);
  CMR_REQUIRE((long long)R * chunks < (1ll << 31));
  roi_align_nchw_bwd_kernel<kChanPerCta>
      <<<R * chunks, pick_threads(outh * outw), 0, as_stream(stream)>>>(
          gy, rois, gx, C, H, W, outh, outw, spatial_scale, sampling_ratio, chunks);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_roi_align_nhwc_fwd(const float* x, int N, int H, int W, int C,
                                      const float* rois, int R, int outh, int outw,
                                      int bin_stride, float spatial_scale, int sampling_ratio,
                                      float* y, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && bin_stride >= 1 && C % 4 == 0);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x && rois && y);
  const int oh_s = ceil_div(outh, bin_stride), ow_s = ceil_div(outw, bin_stride);
  CMR_REQUIRE((long long)R * oh_s * ow_s < (1ll << 31));
  const int C4 = C / 4;
  const int threads = C4 >= 256 ? 256 : (C4 >= 128 ? 128 : (C4 >= 64 ? 64 : 32));
  roi_align_nhwc_fwd_kernel<<<R * oh_s * ow_s, threads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), rois, reinterpret_cast<float4*>(y), H, W, C4, outh,
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_roi_align_nhwc_bwd(const float* gy, const float* rois, int R, int N, int H,
                                      int W, int C, int outh, int outw, int bin_stride,
                                      float spatial_scale, int sampling_ratio, float* gx,
                                      void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && bin_stride >= 1 && C % 4 == 0 && gx);
  CMR_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)N * C * H * W, as_stream(stream)));
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(gy && rois);
  const int oh_s = ceil_div(outh, bin_stride), ow_s = ceil_div(outw, bin_stride);
  CMR_REQUIRE((long long)R * oh_s * ow_s < (1ll << 31));
  const int C4 = C / 4;
  const int threads = C4 >= 256 ? 256 : (C4 >= 128 ? 128 : (C4 >= 64 ? 64 : 32));
  roi_align_nhwc_bwd_kernel<<<R * oh_s * ow_s, threads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(gy), rois, reinterpret_cast<float4*>(gx), H, W, C4, outh,
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
