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

// Per-(bin, sample) taps of one axis, packed {low offset, high offset, l, h}; a
// negative low offset marks a skipped sample.  The geometry of a RoI is
// CTA-uniform, so it is computed once per CTA into shared memory whenever
// pooled * grid <= kMaxTaps on both axes (always, for RoIs clipped to the image);
// otherwise threads compute their taps on the fly (same arithmetic).
constexpr int kMaxTaps = 512;

struct TapTables {
  float4 y[kMaxTaps];
  float4 x[kMaxTaps];
};

__device__ __forceinline__ float4 packed_tap(float start, float bin, int p, int i, int grid,
                                             int limit, int mul) {
  const AxisTap t = axis_tap(start, bin, p, i, grid, limit);
  return make_float4(__int_as_float(t.valid ? t.low * mul : -1), __int_as_float(t.high * mul),
                     t.l, t.h);
}

struct TapSource {
  const TapTables* tt;
  RoiGeom g;
  int H, W, y_mul, x_mul;
  bool tabled;
  __device__ __forceinline__ float4 y(int ph, int iy) const {
    return tabled ? tt->y[ph * g.grid_h + iy]
                  : packed_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, y_mul);
  }
  __device__ __forceinline__ float4 x(int pw, int ix) const {
    return tabled ? tt->x[pw * g.grid_w + ix]
                  : packed_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, x_mul);
  }
};

__device__ __forceinline__ TapSource make_taps(TapTables& tt, const RoiGeom& g, int outh,
                                               int outw, int H, int W, int y_mul, int x_mul) {
  TapSource ts;
  ts.tt = &tt;
  ts.g = g;
  ts.H = H;
  ts.W = W;
  ts.y_mul = y_mul;
  ts.x_mul = x_mul;
  ts.tabled = (long long)outh * g.grid_h <= kMaxTaps && (long long)outw * g.grid_w <= kMaxTaps;
  if (ts.tabled) {
    for (int e = threadIdx.x; e < outh * g.grid_h; e += blockDim.x) {
      const int ph = e / g.grid_h;
      tt.y[e] = packed_tap(g.start_h, g.bin_h, ph, e - ph * g.grid_h, g.grid_h, H, y_mul);
    }
    for (int e = threadIdx.x; e < outw * g.grid_w; e += blockDim.x) {
      const int pw = e / g.grid_w;
      tt.x[e] = packed_tap(g.start_w, g.bin_w, pw, e - pw * g.grid_w, g.grid_w, W, x_mul);
    }
  }
  __syncthreads();
  return ts;
}

// ------------------------------------------------------------------ NCHW --
constexpr int kChanPerCta = 8;

template <int CB>
__global__ void __launch_bounds__(256)
roi_align_nchw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ rois,
                          float* __restrict__ y, int C, int H, int W, int outh, int outw,
                          float scale, int sampling_ratio, int chunks_per_roi) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const TapSource taps = make_taps(tt, g, outh, outw, H, W, W, 1);
  const int P = outh * outw;
  const size_t HW = (size_t)H * W;
  const float* __restrict__ plane0 = x + ((size_t)g.batch * C + c0) * HW;
  float* __restrict__ out0 = y + ((size_t)r * C + c0) * P;
  const int nch = min(CB, C - c0);
  // count is a power of two in the common cases (1, 2, 4, 16): multiplying by the
  // exact reciprocal is then bit-identical to the reference's division.
  const int cnt = g.grid_h * g.grid_w;
  const bool exact_inv = (cnt & (cnt - 1)) == 0;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);

  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int ph = p / outw;
    const int pw = p - ph * outw;
    float acc[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float4 ty = taps.y(ph, iy);
      const int yl = __float_as_int(ty.x), yh = __float_as_int(ty.y);
      if (yl < 0) continue;
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float4 tx = taps.x(pw, ix);
        const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
        if (xl < 0) continue;
        const float w1 = ty.w * tx.w, w2 = ty.w * tx.z, w3 = ty.z * tx.w, w4 = ty.z * tx.z;
        const int o1 = yl + xl, o2 = yl + xh, o3 = yh + xl, o4 = yh + xh;
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
      if (k < nch)
        out0[(size_t)k * P + p] = exact_inv ? acc[k] * inv : __fdiv_rn(acc[k], g.inv_count_den);
  }
}

template <int CB>
__global__ void __launch_bounds__(256)
roi_align_nchw_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ rois,
                          float* __restrict__ gx, int C, int H, int W, int outh, int outw,
                          float scale, int sampling_ratio, int chunks_per_roi) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const TapSource taps = make_taps(tt, g, outh, outw, H, W, W, 1);
  const int P = outh * outw;
  const size_t HW = (size_t)H * W;
  float* __restrict__ plane0 = gx + ((size_t)g.batch * C + c0) * HW;
  const float* __restrict__ in0 = gy + ((size_t)r * C + c0) * P;
  const int nch = min(CB, C - c0);
  const float d = g.inv_count_den;

  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int ph = p / outw;
    const int pw = p - ph * outw;
    float gval[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) gval[k] = (k < nch) ? __ldg(in0 + (size_t)k * P + p) : 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float4 ty = taps.y(ph, iy);
      const int yl = __float_as_int(ty.x), yh = __float_as_int(ty.y);
      if (yl < 0) continue;
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float4 tx = taps.x(pw, ix);
        const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
        if (xl < 0) continue;
        const float w1 = ty.w * tx.w, w2 = ty.w * tx.z, w3 = ty.z * tx.w, w4 = ty.z * tx.z;
        const int o1 = yl + xl, o2 = yl + xh, o3 = yh + xl, o4 = yh + xh;
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          if (k < nch) {
            float* pl = plane0 + (size_t)k * HW;
            // g_k = top_diff * w_k / count  (roi_align_2d.py:501-504)
            atomicAdd(pl + o1, __fdiv_rn(gval[k] * w1, d));
            atomicAdd(pl + o2, __fdiv_rn(gval[k] * w2, d));
            atomicAdd(pl + o3, __fdiv_rn(gval[k] * w3, d));
            atomicAdd(pl + o4, __fdiv_rn(gval[k] * w4, d));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ NHWC --
__device__ __forceinline__ float round_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
  // sm_90+: 128-bit vector reduction to global memory.
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One CTA per (RoI, produced output row); lanes over channels as float4, two channel
// quads per thread (c and c + blockDim.x) so that the sample geometry -- table lookups,
// the four bilinear weights and offsets -- is computed once per 8 channels.  The division
// by the sample count is a multiplication by its reciprocal (exact when the count is a
// power of two, else within 1 ulp of the reference's division; the NCHW drop-in kernels
// keep the exact division).
// forward: src = x (N,H,W,C), dst = y;   backward: src = gy, dst = gx.
template <bool kBackward>
__global__ void __launch_bounds__(128)
roi_align_nhwc_kernel(const float4* __restrict__ src, const float* __restrict__ rois,
                      float4* __restrict__ dst, int H, int W, int C4, int outh, int outw,
                      int bin_stride, int oh_s, int ow_s, float scale, int sampling_ratio,
                      int round_out) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / oh_s;
  const int row = blockIdx.x - r * oh_s;
  const int ph = row * bin_stride;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const TapSource taps = make_taps(tt, g, outh, outw, H, W, W * C4, C4);
  const size_t img_off = (size_t)g.batch * H * W * C4;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  const size_t bin0 = ((size_t)r * oh_s + row) * ow_s;
  const int T = blockDim.x;

  for (int c = threadIdx.x; c < C4; c += 2 * T) {
    const bool two = c + T < C4;
    for (int q = 0; q < ow_s; ++q) {
      const int pw = q * bin_stride;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;   // fwd: sums; bwd: scaled gy
      if (kBackward) {
        a0 = __ldg(src + (bin0 + q) * C4 + c);
        a0.x *= inv; a0.y *= inv; a0.z *= inv; a0.w *= inv;
        if (two) {
          a1 = __ldg(src + (bin0 + q) * C4 + c + T);
          a1.x *= inv; a1.y *= inv; a1.z *= inv; a1.w *= inv;
        }
      }
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const float4 ty = taps.y(ph, iy);
        const int yl = __float_as_int(ty.x), yh = __float_as_int(ty.y);
        if (yl < 0) continue;
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const float4 tx = taps.x(pw, ix);
          const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
          if (xl < 0) continue;
          const float w[4] = {ty.w * tx.w, ty.w * tx.z, ty.z * tx.w, ty.z * tx.z};
          const int o[4] = {yl + xl, yl + xh, yh + xl, yh + xh};
          if (!kBackward) {
            const float4* img = src + img_off + c;
            const float4 v1 = __ldg(img + o[0]), v2 = __ldg(img + o[1]);
            const float4 v3 = __ldg(img + o[2]), v4 = __ldg(img + o[3]);
            a0.x += w[0] * v1.x + w[1] * v2.x + w[2] * v3.x + w[3] * v4.x;
            a0.y += w[0] * v1.y + w[1] * v2.y + w[2] * v3.y + w[3] * v4.y;
            a0.z += w[0] * v1.z + w[1] * v2.z + w[2] * v3.z + w[3] * v4.z;
            a0.w += w[0] * v1.w + w[1] * v2.w + w[2] * v3.w + w[3] * v4.w;
            if (two) {
              const float4 u1 = __ldg(img + o[0] + T), u2 = __ldg(img + o[1] + T);
              const float4 u3 = __ldg(img + o[2] + T), u4 = __ldg(img + o[3] + T);
              a1.x += w[0] * u1.x + w[1] * u2.x + w[2] * u3.x + w[3] * u4.x;
              a1.y += w[0] * u1.y + w[1] * u2.y + w[2] * u3.y + w[3] * u4.y;
              a1.z += w[0] * u1.z + w[1] * u2.z + w[2] * u3.z + w[3] * u4.z;
              a1.w += w[0] * u1.w + w[1] * u2.w + w[2] * u3.w + w[3] * u4.w;
            }
          } else {
            float4* img = dst + img_off + c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              red_add_f4(img + o[k],
                         make_float4(a0.x * w[k], a0.y * w[k], a0.z * w[k], a0.w * w[k]));
              if (two)
                red_add_f4(img + o[k] + T,
                           make_float4(a1.x * w[k], a1.y * w[k], a1.z * w[k], a1.w * w[k]));
            }
          }
        }
      }
      if (!kBackward) {
        float4 o0 = make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv);
        float4 o1 = make_float4(a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv);
        if (round_out) {
          o0.x = round_tf32_rn(o0.x); o0.y = round_tf32_rn(o0.y);
          o0.z = round_tf32_rn(o0.z); o0.w = round_tf32_rn(o0.w);
          o1.x = round_tf32_rn(o1.x); o1.y = round_tf32_rn(o1.y);
          o1.z = round_tf32_rn(o1.z); o1.w = round_tf32_rn(o1.w);
        }
        dst[(bin0 + q) * C4 + c] = o0;
        if (two) dst[(bin0 + q) * C4 + c + T] = o1;
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

// two channel quads per thread
int nhwc_threads(int C4) { return C4 >= 256 ? 128 : (C4 >= 128 ? 64 : 32); }

}  // namespace
}  // namespace cmr

using namespace cmr;

namespace {
// Algorithmic bytes of one ROIAlign pass: the pooled tensor once, the
// feature map once, the RoI table.
double roi_align_bytes(int R, int C, int oh, int ow, int N, int H, int W) {
  return 4.0 * ((double)R * C * oh * ow + (double)N * C * H * W + 5.0 * R);
}
}  // namespace

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
                                      int round_tf32, float* y, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && bin_stride >= 1 && C % 4 == 0);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x && rois && y);
  const int oh_s = ceil_div(outh, bin_stride), ow_s = ceil_div(outw, bin_stride);
  CMR_REQUIRE((long long)R * oh_s < (1ll << 31));
  CMR_REQUIRE((long long)N * H * W * (C / 4) < (1ll << 31));
  prof_begin(kProfRoiAlign, roi_align_bytes(R, C, oh_s, ow_s, N, H, W), as_stream(stream));
  roi_align_nhwc_kernel<false><<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), rois, reinterpret_cast<float4*>(y), H, W, C / 4, outh,
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio, round_tf32);
  prof_end(as_stream(stream));
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
  CMR_REQUIRE((long long)R * oh_s < (1ll << 31));
  CMR_REQUIRE((long long)N * H * W * (C / 4) < (1ll << 31));
  prof_begin(kProfRoiAlignBwd, roi_align_bytes(R, C, oh_s, ow_s, N, H, W), as_stream(stream));
  roi_align_nhwc_kernel<true><<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(gy), rois, reinterpret_cast<float4*>(gx), H, W, C / 4, outh,
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio, 0);
  prof_end(as_stream(stream));
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
