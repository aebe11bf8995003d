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
//  * NHWC  (what the model uses internally).  One CTA = one RoI x one produced
//    output row; lanes run over channels as float4, so every tap is a fully
//    coalesced 512 B warp load and the geometry is CTA-uniform.  The bilinear sum is
//    evaluated separably over a two-column register window (see the kernels).
//    `bin_stride` produces only every bin_stride-th bin (res5.a reads the 14x14 pool
//    with stride 2).
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

// only_ph >= 0: the CTA reads the vertical taps of that output row only.
__device__ __forceinline__ TapSource make_taps(TapTables& tt, const RoiGeom& g, int outh,
                                               int outw, int H, int W, int y_mul, int x_mul,
                                               int only_ph = -1) {
  TapSource ts;
  ts.tt = &tt;
  ts.g = g;
  ts.H = H;
  ts.W = W;
  ts.y_mul = y_mul;
  ts.x_mul = x_mul;
  ts.tabled = (long long)outh * g.grid_h <= kMaxTaps && (long long)outw * g.grid_w <= kMaxTaps;
  if (ts.tabled) {
    const int y_first = only_ph >= 0 ? only_ph * g.grid_h : 0;
    const int y_end = only_ph >= 0 ? y_first + g.grid_h : outh * g.grid_h;
    for (int e = y_first + threadIdx.x; e < y_end; e += blockDim.x) {
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
// quads per thread (c and c + blockDim.x) so that the sample geometry is looked up once
// per 8 channels.
//
// The bilinear sum is evaluated separably.  Every sample of an output row uses the same
// vertical taps, so a feature column enters the row only through its vertical blend
//     g[x] = sum over iy of  hy(iy) * f[y_low(iy)][x] + ly(iy) * f[y_high(iy)][x],
// and bin (ph, pw) = 1/count * sum over ix of  hx(ix) * g[x_low(ix)] + lx(ix) * g[x_high(ix)].
// The samples of a row run left to right, so the thread keeps just the two blended columns
// under the current sample in registers and fetches a feature column (2 * grid_h loads)
// only when the window moves: 2 * grid_h * (columns the RoI spans) loads per row instead of
// 4 * grid_h * grid_w * pooled_w -- neighbouring bins of all but the largest RoIs share
// their columns, and those re-reads were what bound the kernel (L1 bandwidth).  Same
// products as the reference, summed in a different order (<= a few ulp apart; the NCHW
// drop-in kernels keep the reference's order).  The division by the sample count is a
// multiplication by its reciprocal.
// acc += w * v on a channel quad as two packed FFMA2 (sm_100: two fp32 FMAs per issue
// slot; the kernels below are issue-bound).  Same single rounding as the scalar FMA.
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  asm("{\n"
      ".reg .b64 a, b, ww, p, q;\n"
      "mov.b64 ww, {%8, %8};\n"
      "mov.b64 a, {%0, %1};\n"
      "mov.b64 b, {%2, %3};\n"
      "mov.b64 p, {%4, %5};\n"
      "mov.b64 q, {%6, %7};\n"
      "fma.rn.f32x2 a, ww, p, a;\n"
      "fma.rn.f32x2 b, ww, q, b;\n"
      "mov.b64 {%0, %1}, a;\n"
      "mov.b64 {%2, %3}, b;\n"
      "}"
      : "+f"(acc.x), "+f"(acc.y), "+f"(acc.z), "+f"(acc.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(w));
}

// Taps of one output row / of all output columns as the row kernels walk them: from the
// CTA's shared-memory tables (kTabled, the case for every RoI that fits the image) or
// recomputed on the fly (RoIs more than kMaxTaps / pooled samples tall or wide).
template <bool kTabled>
struct RowTaps {
  const float4* ytab;       // grid_h entries of row ph
  const float4* xtab;       // pooled_w * grid_w entries
  RoiGeom g;
  int ph, H, W, y_mul, x_mul;
  __device__ __forceinline__ float4 y(int iy) const {
    if (kTabled) return ytab[iy];
    return packed_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H, y_mul);
  }
  __device__ __forceinline__ float4 x(int pw, int ix) const {
    if (kTabled) return xtab[pw * g.grid_w + ix];
    return packed_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W, x_mul);
  }
};

template <bool kTabled>
__device__ __forceinline__ RowTaps<kTabled> row_taps(const TapTables& tt, const RoiGeom& g,
                                                     int ph, int H, int W, int y_mul,
                                                     int x_mul) {
  RowTaps<kTabled> t;
  t.ytab = tt.y + ph * g.grid_h;
  t.xtab = tt.x;
  t.g = g;
  t.ph = ph; t.H = H; t.W = W; t.y_mul = y_mul; t.x_mul = x_mul;
  return t;
}

// g = vertical blend of the feature column at `col` (its pixel in image row 0; col1 = the
// thread's second channel quad) for the CTA's output row.
// (The row kernels keep tap offsets in BYTES, so that an address is one 64-bit add.)
__device__ __forceinline__ const float4* at(const char* base, int byte_off) {
  return reinterpret_cast<const float4*>(base + byte_off);
}
__device__ __forceinline__ float4* at(char* base, int byte_off) {
  return reinterpret_cast<float4*>(base + byte_off);
}

template <bool kTabled>
__device__ __forceinline__ void blend_column(const RowTaps<kTabled>& taps, int grid_h,
                                             const char* __restrict__ col,
                                             const char* __restrict__ col1, bool two,
                                             float4& g0, float4& g1) {
  g0 = make_float4(0.f, 0.f, 0.f, 0.f);
  g1 = g0;
  for (int iy = 0; iy < grid_h; ++iy) {
    const float4 ty = taps.y(iy);
    const int yl = __float_as_int(ty.x), yh = __float_as_int(ty.y);
    if (yl < 0) continue;
    const float4 v = __ldg(at(col, yl)), u = __ldg(at(col, yh));
    if (two) {
      const float4 v1 = __ldg(at(col1, yl)), u1 = __ldg(at(col1, yh));
      fma4(g1, ty.w, v1);
      fma4(g1, ty.z, u1);
    }
    fma4(g0, ty.w, v);
    fma4(g0, ty.z, u);
  }
}

template <bool kTabled>
__device__ __forceinline__ void roi_align_fwd_row(const RowTaps<kTabled>& taps,
                                                  const float4* __restrict__ img,
                                                  float4* __restrict__ out, int C4, int T,
                                                  int ow_s, int bin_stride, float inv,
                                                  int round_out) {
  const int grid_h = taps.g.grid_h, grid_w = taps.g.grid_w;
  for (int c = threadIdx.x; c < C4; c += 2 * T) {
    const bool two = c + T < C4;
    const char* img0 = reinterpret_cast<const char*>(img + c);
    const char* img1 = reinterpret_cast<const char*>(img + c + T);
    float4* o = out + c;
    int lo = -1, hi = -1;                       // x offsets of the two cached blended columns
    float4 g0lo, g1lo, g0hi, g1hi;
    g0lo = g1lo = g0hi = g1hi = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < ow_s; ++q, o += C4) {
      const int pw = q * bin_stride;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
      for (int ix = 0; ix < grid_w; ++ix) {
        const float4 tx = taps.x(pw, ix);
        const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
        if (xl < 0) continue;
        if (xl != lo) {
          if (xl == hi) {
            g0lo = g0hi; g1lo = g1hi;
          } else {
            blend_column<kTabled>(taps, grid_h, img0 + xl, img1 + xl, two, g0lo, g1lo);
          }
          lo = xl;
          hi = -1;
        }
        if (xh != hi) {
          if (xh == lo) {                       // right border: both taps on the last column
            g0hi = g0lo; g1hi = g1lo;
          } else {
            blend_column<kTabled>(taps, grid_h, img0 + xh, img1 + xh, two, g0hi, g1hi);
          }
          hi = xh;
        }
        fma4(a0, tx.w, g0lo);
        fma4(a0, tx.z, g0hi);
        fma4(a1, tx.w, g1lo);
        fma4(a1, tx.z, g1hi);
      }
      float4 o0 = make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv);
      float4 o1 = make_float4(a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv);
      if (round_out) {
        o0.x = round_tf32_rn(o0.x); o0.y = round_tf32_rn(o0.y);
        o0.z = round_tf32_rn(o0.z); o0.w = round_tf32_rn(o0.w);
        o1.x = round_tf32_rn(o1.x); o1.y = round_tf32_rn(o1.y);
        o1.z = round_tf32_rn(o1.z); o1.w = round_tf32_rn(o1.w);
      }
      o[0] = o0;
      if (two) o[T] = o1;
    }
  }
}

__global__ void __launch_bounds__(128)   // (128, 8) = 64 registers was measured: slower
roi_align_nhwc_fwd_kernel(const float4* __restrict__ src, const float* __restrict__ rois,
                          float4* __restrict__ dst, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale, int sampling_ratio,
                          int round_out) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / oh_s;
  const int row = blockIdx.x - r * oh_s;
  const int ph = row * bin_stride;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const TapSource ts = make_taps(tt, g, outh, outw, H, W, W * C4 * 16, C4 * 16, ph);
  const float4* img = src + (size_t)g.batch * H * W * C4;
  float4* out = dst + ((size_t)r * oh_s + row) * ow_s * C4;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  if (ts.tabled)
    roi_align_fwd_row<true>(row_taps<true>(tt, g, ph, H, W, W * C4 * 16, C4 * 16), img, out, C4,
                            blockDim.x, ow_s, bin_stride, inv, round_out);
  else
    roi_align_fwd_row<false>(row_taps<false>(tt, g, ph, H, W, W * C4 * 16, C4 * 16), img, out, C4,
                             blockDim.x, ow_s, bin_stride, inv, round_out);
}

// Backward, the transpose of the above: the gradients of a row's bins are first gathered
// per feature column (h[x] = sum of hx * gy over the samples whose low tap is x, lx * gy
// over those whose high tap is x), and a column is scattered through the vertical taps
// when the window leaves it: 2 * grid_h vector reductions per column of the RoI.
template <bool kTabled>
__device__ __forceinline__ void scatter_column(const RowTaps<kTabled>& taps, int grid_h,
                                               char* __restrict__ col,
                                               char* __restrict__ col1, bool two,
                                               const float4& h0, const float4& h1) {
  for (int iy = 0; iy < grid_h; ++iy) {
    const float4 ty = taps.y(iy);
    const int yl = __float_as_int(ty.x), yh = __float_as_int(ty.y);
    if (yl < 0) continue;
    red_add_f4(at(col, yl), make_float4(h0.x * ty.w, h0.y * ty.w, h0.z * ty.w, h0.w * ty.w));
    red_add_f4(at(col, yh), make_float4(h0.x * ty.z, h0.y * ty.z, h0.z * ty.z, h0.w * ty.z));
    if (two) {
      red_add_f4(at(col1, yl), make_float4(h1.x * ty.w, h1.y * ty.w, h1.z * ty.w, h1.w * ty.w));
      red_add_f4(at(col1, yh), make_float4(h1.x * ty.z, h1.y * ty.z, h1.z * ty.z, h1.w * ty.z));
    }
  }
}

template <bool kTabled>
__device__ __forceinline__ void roi_align_bwd_row(const RowTaps<kTabled>& taps,
                                                  const float4* __restrict__ gyrow,
                                                  float4* __restrict__ img, int C4, int T,
                                                  int ow_s, int bin_stride, float inv) {
  const int grid_h = taps.g.grid_h, grid_w = taps.g.grid_w;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = threadIdx.x; c < C4; c += 2 * T) {
    const bool two = c + T < C4;
    char* img0 = reinterpret_cast<char*>(img + c);
    char* img1 = reinterpret_cast<char*>(img + c + T);
    const float4* gp = gyrow + c;
    int lo = -1, hi = -1;
    float4 h0lo = zero, h1lo = zero, h0hi = zero, h1hi = zero;
    // the gradient of the next bin is requested while the current one is scattered
    float4 n0 = __ldg(gp), n1 = two ? __ldg(gp + T) : zero;
    for (int q = 0; q < ow_s; ++q) {
      const int pw = q * bin_stride;
      float4 a0 = n0, a1 = n1;
      gp += C4;
      if (q + 1 < ow_s) {
        n0 = __ldg(gp);
        if (two) n1 = __ldg(gp + T);
      }
      a0.x *= inv; a0.y *= inv; a0.z *= inv; a0.w *= inv;
      a1.x *= inv; a1.y *= inv; a1.z *= inv; a1.w *= inv;
      for (int ix = 0; ix < grid_w; ++ix) {
        const float4 tx = taps.x(pw, ix);
        const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
        if (xl < 0) continue;
        if (xl != lo) {
          if (lo >= 0)
            scatter_column<kTabled>(taps, grid_h, img0 + lo, img1 + lo, two, h0lo, h1lo);
          if (xl == hi) {
            h0lo = h0hi; h1lo = h1hi;
          } else {
            if (hi >= 0)
              scatter_column<kTabled>(taps, grid_h, img0 + hi, img1 + hi, two, h0hi, h1hi);
            h0lo = zero; h1lo = zero;
          }
          lo = xl;
          hi = -1;
          h0hi = zero; h1hi = zero;
        }
        fma4(h0lo, tx.w, a0);
        fma4(h1lo, tx.w, a1);
        if (xh == lo) {                         // right border: the high tap is the same pixel
          fma4(h0lo, tx.z, a0);
          fma4(h1lo, tx.z, a1);
        } else {
          if (xh != hi) {
            if (hi >= 0)
              scatter_column<kTabled>(taps, grid_h, img0 + hi, img1 + hi, two, h0hi, h1hi);
            hi = xh;
            h0hi = zero; h1hi = zero;
          }
          fma4(h0hi, tx.z, a0);
          fma4(h1hi, tx.z, a1);
        }
      }
    }
    if (lo >= 0) scatter_column<kTabled>(taps, grid_h, img0 + lo, img1 + lo, two, h0lo, h1lo);
    if (hi >= 0) scatter_column<kTabled>(taps, grid_h, img0 + hi, img1 + hi, two, h0hi, h1hi);
  }
}

__global__ void __launch_bounds__(128, 6)   // 80 registers: 6 CTAs per SM (+5 % over 95 / 5)
roi_align_nhwc_bwd_kernel(const float4* __restrict__ gy, const float* __restrict__ rois,
                          float4* __restrict__ gx, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale, int sampling_ratio) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / oh_s;
  const int row = blockIdx.x - r * oh_s;
  const int ph = row * bin_stride;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio);
  const TapSource ts = make_taps(tt, g, outh, outw, H, W, W * C4 * 16, C4 * 16, ph);
  float4* img = gx + (size_t)g.batch * H * W * C4;
  const float4* gyrow = gy + ((size_t)r * oh_s + row) * ow_s * C4;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  if (ts.tabled)
    roi_align_bwd_row<true>(row_taps<true>(tt, g, ph, H, W, W * C4 * 16, C4 * 16), gyrow, img, C4,
                            blockDim.x, ow_s, bin_stride, inv);
  else
    roi_align_bwd_row<false>(row_taps<false>(tt, g, ph, H, W, W * C4 * 16, C4 * 16), gyrow, img, C4,
                             blockDim.x, ow_s, bin_stride, inv);
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
  CMR_REQUIRE((long long)H * W * C * 4 < (1ll << 31));   // byte offsets inside an image are ints
  prof_begin(kProfRoiAlign, roi_align_bytes(R, C, oh_s, ow_s, N, H, W), as_stream(stream));
  roi_align_nhwc_fwd_kernel<<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
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
  CMR_REQUIRE((long long)H * W * C * 4 < (1ll << 31));   // byte offsets inside an image are ints
  prof_begin(kProfRoiAlignBwd, roi_align_bytes(R, C, oh_s, ow_s, N, H, W), as_stream(stream));
  roi_align_nhwc_bwd_kernel<<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(gy), rois, reinterpret_cast<float4*>(gx), H, W, C / 4, outh,
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio);
  prof_end(as_stream(stream));
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
