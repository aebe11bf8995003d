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
#include <stdlib.h>

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
// A batch index outside [0, n_img) (or NaN) makes the RoI empty: no samples, zero output,
// no gradient -- never an out-of-bounds access.
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi,
                                                float scale, int outh, int outw,
                                                int sampling_ratio, int n_img) {
  RoiGeom g;
  const float bf = roi[0];
  if (!(bf >= 0.0f && bf < (float)n_img)) {
    g.batch = 0;
    g.start_w = g.start_h = 0.f;
    g.bin_w = g.bin_h = 1.f;
    g.grid_h = g.grid_w = 0;
    g.inv_count_den = 1.f;
    return g;
  }
  g.batch = (int)bf;
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
                          float scale, int sampling_ratio, int chunks_per_roi, int n_img) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
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
                          float scale, int sampling_ratio, int chunks_per_roi, int n_img) {
  __shared__ TapTables tt;
  const int r = blockIdx.x / chunks_per_roi;
  const int c0 = (blockIdx.x - r * chunks_per_roi) * CB;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
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

// ---------------------------------------------------------------- merged rows --
// The vertical taps of one output row, merged per feature row.  Consecutive samples of a bin
// are at most one pixel apart (grid = ceil(bin size)), so the 2 * grid_h (row, weight) taps
// of an output row fall on a short run of consecutive feature rows; summing the weights per
// row first turns 2 * grid_h loads (forward) / vector reductions (backward) per feature
// column into (rows spanned) of them: ~1.5x fewer for the usual RoIs.  Same products as the
// reference up to the order of the weight sum (a few ulp).
constexpr int kRowSlots = 8;      // feature rows one output row may span on this path
constexpr int kColTaps = 256;     // pooled_w * grid_w x-taps kept in shared memory

struct RowBlend {
  int first;                // byte offset of the first feature row, -1: every sample skipped
  int n;                    // rows spanned (0 when every sample is skipped)
  float w[kRowSlots];
};

// Threads e < nph * kRowSlots fill rb[0 .. nph) for output rows ph0, ph0 + ph_step, ...;
// returns false (CTA-uniform, through *ok) when some row spans more than kRowSlots rows.
__device__ __forceinline__ void build_row_blends(RowBlend* rb, int* ok, const RoiGeom& g,
                                                 int ph0, int ph_step, int nph, int H,
                                                 int row_bytes) {
  for (int e = threadIdx.x; e < nph * kRowSlots; e += blockDim.x) {
    const int i = e / kRowSlots, k = e - i * kRowSlots;
    const int ph = ph0 + i * ph_step;
    int lo = 1 << 30, hi = -1;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
      if (!t.valid) continue;
      lo = min(lo, t.low);
      hi = max(hi, t.high);
    }
    const int row = lo + k;
    float w = 0.f;
    if (hi >= 0 && row <= hi) {
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
        if (!t.valid) continue;
        if (t.low == row) w += t.h;
        if (t.high == row) w += t.l;
      }
    }
    rb[i].w[k] = w;
    if (k == 0) {
      rb[i].first = hi >= 0 ? lo * row_bytes : -1;
      rb[i].n = hi >= 0 ? hi - lo + 1 : 0;
      if (hi - lo + 1 > kRowSlots) atomicExch(ok, 0);
    }
  }
}

__device__ __forceinline__ void build_col_taps(float4* xt, const RoiGeom& g, int outw, int W,
                                               int px_bytes) {
  for (int e = threadIdx.x; e < outw * g.grid_w; e += blockDim.x) {
    const int pw = e / g.grid_w;
    xt[e] = packed_tap(g.start_w, g.bin_w, pw, e - pw * g.grid_w, g.grid_w, W, px_bytes);
  }
}

// g[j] = sum over the rows of rb of w * f[row][column at byte offset xoff], for the NQ
// channel quads of the thread (quad j sits qs bytes after quad j - 1)
template <int NQ>
__device__ __forceinline__ void blend_rows(const RowBlend& rb, const char* __restrict__ img,
                                           int xoff, int row_bytes, int qs, float4 (&g)[NQ]) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const char* p = img + rb.first + xoff;
  int k = 0;
  for (; k + 1 < rb.n; k += 2, p += 2 * row_bytes) {
    float4 v0[NQ], v1[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      v0[j] = __ldg(reinterpret_cast<const float4*>(p + j * qs));
      v1[j] = __ldg(reinterpret_cast<const float4*>(p + row_bytes + j * qs));
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      fma4(g[j], rb.w[k], v0[j]);
      fma4(g[j], rb.w[k + 1], v1[j]);
    }
  }
  if (k < rb.n) {
#pragma unroll
    for (int j = 0; j < NQ; ++j)
      fma4(g[j], rb.w[k], __ldg(reinterpret_cast<const float4*>(p + j * qs)));
  }
}

// FIXED: the destination is the deterministic mode's int64 fixed-point image (common.cuh):
// same element order, 8 bytes per element, four scalar 64-bit reductions per quad.
__device__ __forceinline__ void red_fixed_f4(char* addr, const float4& v) {
  long long* q = reinterpret_cast<long long*>(addr);
  red_fixed(q, v.x); red_fixed(q + 1, v.y); red_fixed(q + 2, v.z); red_fixed(q + 3, v.w);
}

template <int NQ, bool FIXED = false>
__device__ __forceinline__ void scatter_rows(const RowBlend& rb, char* __restrict__ img, int xoff,
                                             int row_bytes, int qs, const float4 (&h)[NQ]) {
  constexpr int kW = FIXED ? 2 : 1;           // bytes per element relative to fp32
  char* p = img + (size_t)kW * (rb.first + xoff);
  for (int k = 0; k < rb.n; ++k, p += kW * row_bytes) {
    const float w = rb.w[k];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const float4 v = make_float4(h[j].x * w, h[j].y * w, h[j].z * w, h[j].w * w);
      if (FIXED) red_fixed_f4(p + kW * j * qs, v);
      else red_add_f4(reinterpret_cast<float4*>(p + j * qs), v);
    }
  }
}

// One output row of NQ channel quads, forward: bins left to right over a two-column window
// of vertically blended feature columns (described above); emit(q, j, value) receives
// finished bin q (bin index pw0 + q * pw_step) of quad j.
template <int NQ, typename Emit>
__device__ __forceinline__ void walk_row_fwd(const RowBlend& rb, const float4* __restrict__ xt,
                                             int grid_w, int pw0, int pw_step, int npw,
                                             const char* __restrict__ img, int row_bytes, int qs,
                                             float inv, Emit emit) {
  int lo = -1, hi = -1;
  float4 glo[NQ], ghi[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) glo[j] = ghi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = 0; q < npw; ++q) {
    const int pw = pw0 + q * pw_step;
    float4 a[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rb.n > 0) {
      for (int ix = 0; ix < grid_w; ++ix) {
        const float4 tx = xt[pw * grid_w + ix];
        const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
        if (xl < 0) continue;
        if (xl != lo) {
          if (xl == hi) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) glo[j] = ghi[j];
          } else {
            blend_rows<NQ>(rb, img, xl, row_bytes, qs, glo);
          }
          lo = xl;
          hi = -1;
        }
        if (xh != hi) {
          if (xh == lo) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) ghi[j] = glo[j];
          } else {
            blend_rows<NQ>(rb, img, xh, row_bytes, qs, ghi);
          }
          hi = xh;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          fma4(a[j], tx.w, glo[j]);
          fma4(a[j], tx.z, ghi[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j)
      emit(q, j, make_float4(a[j].x * inv, a[j].y * inv, a[j].z * inv, a[j].w * inv));
  }
}

// Backward of the same walk: fetch(q, j) returns the gradient of bin q of quad j.
template <int NQ, bool FIXED = false, typename Fetch>
__device__ __forceinline__ void walk_row_bwd(const RowBlend& rb, const float4* __restrict__ xt,
                                             int grid_w, int pw0, int pw_step, int npw,
                                             char* __restrict__ img, int row_bytes, int qs,
                                             float inv, Fetch fetch) {
  if (rb.n <= 0) return;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  int lo = -1, hi = -1;
  float4 hlo[NQ], hhi[NQ], nxt[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    hlo[j] = hhi[j] = zero;
    nxt[j] = fetch(0, j);
  }
  for (int q = 0; q < npw; ++q) {
    const int pw = pw0 + q * pw_step;
    float4 a[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      a[j] = make_float4(nxt[j].x * inv, nxt[j].y * inv, nxt[j].z * inv, nxt[j].w * inv);
      if (q + 1 < npw) nxt[j] = fetch(q + 1, j);   // requested while bin q is scattered
    }
    for (int ix = 0; ix < grid_w; ++ix) {
      const float4 tx = xt[pw * grid_w + ix];
      const int xl = __float_as_int(tx.x), xh = __float_as_int(tx.y);
      if (xl < 0) continue;
      if (xl != lo) {
        if (lo >= 0) scatter_rows<NQ, FIXED>(rb, img, lo, row_bytes, qs, hlo);
        if (xl == hi) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) hlo[j] = hhi[j];
        } else {
          if (hi >= 0) scatter_rows<NQ, FIXED>(rb, img, hi, row_bytes, qs, hhi);
#pragma unroll
          for (int j = 0; j < NQ; ++j) hlo[j] = zero;
        }
        lo = xl;
        hi = -1;
#pragma unroll
        for (int j = 0; j < NQ; ++j) hhi[j] = zero;
      }
#pragma unroll
      for (int j = 0; j < NQ; ++j) fma4(hlo[j], tx.w, a[j]);
      if (xh == lo) {                           // right border: the high tap is the same pixel
#pragma unroll
        for (int j = 0; j < NQ; ++j) fma4(hlo[j], tx.z, a[j]);
      } else {
        if (xh != hi) {
          if (hi >= 0) scatter_rows<NQ, FIXED>(rb, img, hi, row_bytes, qs, hhi);
          hi = xh;
#pragma unroll
          for (int j = 0; j < NQ; ++j) hhi[j] = zero;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) fma4(hhi[j], tx.z, a[j]);
      }
    }
  }
  if (lo >= 0) scatter_rows<NQ, FIXED>(rb, img, lo, row_bytes, qs, hlo);
  if (hi >= 0) scatter_rows<NQ, FIXED>(rb, img, hi, row_bytes, qs, hhi);
}

// Generic (any RoI size) single-bin evaluation on a channels-last map; the rare path for
// RoIs whose rows span more than kRowSlots feature rows or whose x taps do not fit.
__device__ __forceinline__ float4 bin_fwd_generic(const RoiGeom& g, int ph, int pw, int H, int W,
                                                  const char* __restrict__ img, int row_bytes,
                                                  int px_bytes, float inv) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
    if (!ty.valid) continue;
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
      if (!tx.valid) continue;
      const char* r0 = img + (size_t)ty.low * row_bytes;
      const char* r1 = img + (size_t)ty.high * row_bytes;
      fma4(a, ty.h * tx.h, __ldg(reinterpret_cast<const float4*>(r0 + tx.low * px_bytes)));
      fma4(a, ty.h * tx.l, __ldg(reinterpret_cast<const float4*>(r0 + tx.high * px_bytes)));
      fma4(a, ty.l * tx.h, __ldg(reinterpret_cast<const float4*>(r1 + tx.low * px_bytes)));
      fma4(a, ty.l * tx.l, __ldg(reinterpret_cast<const float4*>(r1 + tx.high * px_bytes)));
    }
  }
  return make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

template <bool FIXED = false>
__device__ __forceinline__ void bin_bwd_generic(const RoiGeom& g, int ph, int pw, int H, int W,
                                                char* __restrict__ img, int row_bytes,
                                                int px_bytes, float inv, float4 gv) {
  constexpr int kW = FIXED ? 2 : 1;
  gv.x *= inv; gv.y *= inv; gv.z *= inv; gv.w *= inv;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
    if (!ty.valid) continue;
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
      if (!tx.valid) continue;
      char* r0 = img + (size_t)kW * ty.low * row_bytes;
      char* r1 = img + (size_t)kW * ty.high * row_bytes;
      const float w[4] = {ty.h * tx.h, ty.h * tx.l, ty.l * tx.h, ty.l * tx.l};
      char* dst[4] = {r0 + (size_t)kW * tx.low * px_bytes, r0 + (size_t)kW * tx.high * px_bytes,
                      r1 + (size_t)kW * tx.low * px_bytes, r1 + (size_t)kW * tx.high * px_bytes};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 v = make_float4(gv.x * w[t], gv.y * w[t], gv.z * w[t], gv.w * w[t]);
        if (FIXED) red_fixed_f4(dst[t], v);
        else red_add_f4(reinterpret_cast<float4*>(dst[t]), v);
      }
    }
  }
}

// ------------------------------------------------------- channels-last in and out --
// What the model runs (ResNetRoIHead): one CTA per (RoI, produced output row), lanes over
// channel quads, two quads per thread (c and c + blockDim.x) so that the taps are looked up
// once per 8 channels and twice the loads are in flight.
struct RowKernelSmem {
  int ok;
  RowBlend rb;
  float4 xt[kColTaps];
};

__global__ void __launch_bounds__(128, 6)   // (128, 8) = 64 registers was measured: slower
roi_align_nhwc_fwd_kernel(const float4* __restrict__ src, const float* __restrict__ rois,
                          float4* __restrict__ dst, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale, int sampling_ratio,
                          int round_out, int n_img) {
  __shared__ RowKernelSmem sm;
  const int r = blockIdx.x / oh_s;
  const int row = blockIdx.x - r * oh_s;
  const int ph = row * bin_stride;
  const int px_bytes = C4 * 16, row_bytes = W * px_bytes;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
  const bool fits = outw * g.grid_w <= kColTaps;
  if (threadIdx.x == 0) sm.ok = fits ? 1 : 0;
  __syncthreads();
  build_row_blends(&sm.rb, &sm.ok, g, ph, 1, 1, H, row_bytes);
  if (fits) build_col_taps(sm.xt, g, outw, W, px_bytes);
  __syncthreads();
  const bool fast = sm.ok != 0;
  const char* img = reinterpret_cast<const char*>(src + (size_t)g.batch * H * W * C4);
  float4* out = dst + ((size_t)r * oh_s + row) * ow_s * C4;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  const int T = blockDim.x;
  for (int c = threadIdx.x; c < C4; c += 2 * T) {
    float4* o = out + c;
    auto emit = [&](int q, int j, float4 v) {
      if (round_out) {
        v.x = round_tf32_rn(v.x); v.y = round_tf32_rn(v.y);
        v.z = round_tf32_rn(v.z); v.w = round_tf32_rn(v.w);
      }
      o[(size_t)q * C4 + j * T] = v;
    };
    const char* base = img + (size_t)c * 16;
    if (!fast) {
      for (int j = 0; j < 2 && c + j * T < C4; ++j)
        for (int q = 0; q < ow_s; ++q)
          emit(q, j, bin_fwd_generic(g, ph, q * bin_stride, H, W, base + (size_t)j * T * 16,
                                     row_bytes, px_bytes, inv));
    } else if (c + T < C4) {
      walk_row_fwd<2>(sm.rb, sm.xt, g.grid_w, 0, bin_stride, ow_s, base, row_bytes, T * 16, inv,
                      emit);
    } else {
      walk_row_fwd<1>(sm.rb, sm.xt, g.grid_w, 0, bin_stride, ow_s, base, row_bytes, 0, inv, emit);
    }
  }
}

template <bool FIXED>
__global__ void __launch_bounds__(128, 6)
roi_align_nhwc_bwd_kernel(const float4* __restrict__ gy, const float* __restrict__ rois,
                          float4* __restrict__ gx, int H, int W, int C4, int outh, int outw,
                          int bin_stride, int oh_s, int ow_s, float scale, int sampling_ratio,
                          int n_img) {
  __shared__ RowKernelSmem sm;
  const int r = blockIdx.x / oh_s;
  const int row = blockIdx.x - r * oh_s;
  const int ph = row * bin_stride;
  const int px_bytes = C4 * 16, row_bytes = W * px_bytes;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
  const bool fits = outw * g.grid_w <= kColTaps;
  if (threadIdx.x == 0) sm.ok = fits ? 1 : 0;
  __syncthreads();
  build_row_blends(&sm.rb, &sm.ok, g, ph, 1, 1, H, row_bytes);
  if (fits) build_col_taps(sm.xt, g, outw, W, px_bytes);
  __syncthreads();
  const bool fast = sm.ok != 0;
  // (FIXED: gx is the int64 fixed-point image, two float4 slots per quad)
  char* img = reinterpret_cast<char*>(gx + (size_t)(FIXED ? 2 : 1) * g.batch * H * W * C4);
  const float4* gyrow = gy + ((size_t)r * oh_s + row) * ow_s * C4;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  const int T = blockDim.x;
  for (int c = threadIdx.x; c < C4; c += 2 * T) {
    const float4* gp = gyrow + c;
    auto fetch = [&](int q, int j) { return __ldg(gp + (size_t)q * C4 + j * T); };
    char* base = img + (size_t)(FIXED ? 2 : 1) * c * 16;
    if (!fast) {
      for (int j = 0; j < 2 && c + j * T < C4; ++j)
        for (int q = 0; q < ow_s; ++q)
          bin_bwd_generic<FIXED>(g, ph, q * bin_stride, H, W,
                                 base + (size_t)(FIXED ? 2 : 1) * j * T * 16, row_bytes,
                                 px_bytes, inv, fetch(q, j));
    } else if (c + T < C4) {
      walk_row_bwd<2, FIXED>(sm.rb, sm.xt, g.grid_w, 0, bin_stride, ow_s, base, row_bytes, T * 16, inv,
                      fetch);
    } else {
      walk_row_bwd<1, FIXED>(sm.rb, sm.xt, g.grid_w, 0, bin_stride, ow_s, base, row_bytes, 0, inv, fetch);
    }
  }
}

// ------------------------------------------- reference layout over a channels-last map --
// The drop-in operator's fast path: the feature map is read channels-last (the model's own
// layout; a plain NCHW input is re-laid once -- it is 2 % of the pooled tensor), the pooled
// tensor is written / read in the reference's (R, C, outh, outw) layout.
//
// One CTA = one RoI x CH channels.  Thread (ph, quad) walks output row ph of one channel
// quad with the separable two-column window: every load is a float4 of 4 channels and a
// warp's loads cover whole 128-byte lines of the map.  The CTA's (CH, outh*outw) output
// block is contiguous in the NCHW pooled tensor: it is staged in shared memory (transposed
// on the way in, plane stride outh*outw + 1 to spread the banks) and written out as full
// 128-byte lines (forward) / read in as full lines and walked from shared memory (backward).
struct RoiTileSmem {
  int ok;
  int pad_[3];
  float4 xt[kColTaps];
};

// Element (channel c, bin p) of the CTA's staging tile.  kVec (outw even, outh*outw % 4 == 0,
// the 14 x 14 case): rows of PS = round_up(P, 32) floats whose 16-byte groups are XOR-ed with
// the channel quad's index -- the walkers access it a channel quad x two bins at a time
// (8-byte accesses, banks spread by the XOR), the copy to / from global memory runs over whole
// float4s of one channel (conflict-free, 512 B per warp instruction).  Otherwise: plain rows
// of P + 1 floats, scalar accesses.
template <bool kVec>
__device__ __forceinline__ int tile_index(int c, int p, int PS) {
  return kVec ? c * PS + (p ^ (((c >> 2) & 7) << 2)) : c * PS + p;
}

// One CTA serves `cpc` consecutive channel chunks of its RoI, so the tap tables are built
// once per cpc * CH channels (they are 15 % of the instructions when built per chunk).
// Backward walker: staging tile of 32 channels (128 threads, 64 registers, 29 KB: six to seven
// CTAs = 24-28 warps per SM) and a grid of at least 96 CTAs per SM.  Same-box A/B at
// R = 300 / 1000 / 2000 / 6000, GB/s: 64 channels, 3 CTAs of 256 threads, grid >= 8 per SM
// 1382 / 1490 / 1468 / 1684; 2 CTAs 1107 / 1201 / 1188 / 1309 (it is latency-bound: warps
// count); 32 channels with the old grid rule 1189 / 1373 / 1593 / 1824; grid >= 48 per SM
// 1565 / 1716 / 1718 / 1823; >= 96: 1571 / 1736 / 1774 / 1860.  (Macros: A/B builds.)
#ifndef CMR_CL_BWD_CTAS
#define CMR_CL_BWD_CTAS 6
#endif
constexpr int kClBwdCtas = CMR_CL_BWD_CTAS;
#ifndef CMR_CL_BWD_CH
#define CMR_CL_BWD_CH 32
#endif
constexpr int kClBwdChannels = CMR_CL_BWD_CH;   // channels per staging tile of the backward walker
#ifndef CMR_CL_BWD_GRID
#define CMR_CL_BWD_GRID 96
#endif
constexpr int kClBwdGridCtas = CMR_CL_BWD_GRID;  // CTAs per SM the grid should at least have

template <int CH, bool kVec>
__global__ void __launch_bounds__(256, 3)      // the staging tile allows three CTAs per SM
roi_align_cl_fwd_kernel(const float4* __restrict__ src, const float* __restrict__ rois,
                        float* __restrict__ dst, int H, int W, int C, int outh, int outw,
                        float scale, int sampling_ratio, int groups, int cpc, int n_img) {
  extern __shared__ __align__(16) unsigned char roi_smem[];
  RoiTileSmem* sm = reinterpret_cast<RoiTileSmem*>(roi_smem);
  float* tile = reinterpret_cast<float*>(roi_smem + sizeof(RoiTileSmem));
  constexpr int Q = CH / 4;
  const int r = blockIdx.x / groups;
  const int c_begin = (blockIdx.x - r * groups) * cpc * CH;
  const int c_end = min(C, c_begin + cpc * CH);
  const int P = outh * outw;
  const int PS = kVec ? (P + 31) / 32 * 32 : P + 1;
  RowBlend* rb = reinterpret_cast<RowBlend*>(tile + (size_t)CH * PS);
  const int C4 = C >> 2;
  const int px_bytes = C4 * 16, row_bytes = W * px_bytes;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
  if (threadIdx.x == 0) sm->ok = outw * g.grid_w <= kColTaps ? 1 : 0;
  __syncthreads();
  build_row_blends(rb, &sm->ok, g, 0, 1, outh, H, row_bytes);
  if (outw * g.grid_w <= kColTaps) build_col_taps(sm->xt, g, outw, W, px_bytes);
  __syncthreads();
  const bool fast = sm->ok != 0;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  const int quad = threadIdx.x % Q;
  const int row0 = threadIdx.x / Q, rows_per_pass = blockDim.x / Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int cq = 4 * quad;
  // this thread's four channel rows of the tile and the XOR of its 16-byte groups
  float* const t0 = tile + (size_t)cq * PS;
  const int swz = kVec ? (quad & 7) << 2 : 0;
  for (int c0 = c_begin; c0 < c_end; c0 += CH) {
    const int nch = min(CH, C - c0);
    if (cq < nch) {
      const char* img = reinterpret_cast<const char*>(src + (size_t)g.batch * H * W * C4 +
                                                      (c0 >> 2) + quad);
      for (int ph = row0; ph < outh; ph += rows_per_pass) {
        const int p0 = ph * outw;
        float4 even = make_float4(0.f, 0.f, 0.f, 0.f);
        auto emit = [&](int pw, int, const float4& v) {
          if (kVec) {
            // two bins of a channel are 8 contiguous bytes of the tile (p0 + pw is even)
            if ((pw & 1) == 0) {
              even = v;
            } else {
              float* t = t0 + ((p0 + pw - 1) ^ swz);
              *reinterpret_cast<float2*>(t) = make_float2(even.x, v.x);
              *reinterpret_cast<float2*>(t + PS) = make_float2(even.y, v.y);
              *reinterpret_cast<float2*>(t + 2 * PS) = make_float2(even.z, v.z);
              *reinterpret_cast<float2*>(t + 3 * PS) = make_float2(even.w, v.w);
            }
          } else {
            float* t = t0 + p0 + pw;
            t[0] = v.x; t[PS] = v.y; t[2 * PS] = v.z; t[3 * PS] = v.w;
          }
        };
        if (fast) {
          walk_row_fwd<1>(rb[ph], sm->xt, g.grid_w, 0, 1, outw, img, row_bytes, 0, inv, emit);
        } else {
          for (int pw = 0; pw < outw; ++pw)
            emit(pw, 0, bin_fwd_generic(g, ph, pw, H, W, img, row_bytes, px_bytes, inv));
        }
      }
    }
    __syncthreads();
    // the chunk's (nch, P) block is contiguous in the pooled tensor: a warp per channel
    float* out = dst + ((size_t)r * C + c0) * P;
    for (int c = warp; c < nch; c += nwarp) {
      const float* tp = tile + (size_t)c * PS;
      float* op = out + (size_t)c * P;
      if (kVec) {
        const int sw = ((c >> 2) & 7) << 2;
        for (int q4 = lane; q4 < (P >> 2); q4 += 32)
          reinterpret_cast<float4*>(op)[q4] = *reinterpret_cast<const float4*>(tp + ((4 * q4) ^ sw));
      } else {
        for (int p = lane; p < P; p += 32) op[p] = tp[p];
      }
    }
    if (c0 + CH < c_end) __syncthreads();      // the tile is rewritten by the next chunk
  }
}

template <int CH, bool kVec>
__global__ void __launch_bounds__(CH * 4, kClBwdCtas)
roi_align_cl_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ rois,
                        float4* __restrict__ gx, int H, int W, int C, int outh, int outw,
                        float scale, int sampling_ratio, int groups, int cpc, int n_img) {
  extern __shared__ __align__(16) unsigned char roi_smem[];
  RoiTileSmem* sm = reinterpret_cast<RoiTileSmem*>(roi_smem);
  float* tile = reinterpret_cast<float*>(roi_smem + sizeof(RoiTileSmem));
  constexpr int Q = CH / 4;
  const int r = blockIdx.x / groups;
  const int c_begin = (blockIdx.x - r * groups) * cpc * CH;
  const int c_end = min(C, c_begin + cpc * CH);
  const int P = outh * outw;
  const int PS = kVec ? (P + 31) / 32 * 32 : P + 1;
  RowBlend* rb = reinterpret_cast<RowBlend*>(tile + (size_t)CH * PS);
  const int C4 = C >> 2;
  const int px_bytes = C4 * 16, row_bytes = W * px_bytes;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
  if (threadIdx.x == 0) sm->ok = outw * g.grid_w <= kColTaps ? 1 : 0;
  __syncthreads();
  build_row_blends(rb, &sm->ok, g, 0, 1, outh, H, row_bytes);
  if (outw * g.grid_w <= kColTaps) build_col_taps(sm->xt, g, outw, W, px_bytes);
  __syncthreads();
  const bool fast = sm->ok != 0;
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  const int quad = threadIdx.x % Q;
  const int row0 = threadIdx.x / Q, rows_per_pass = blockDim.x / Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int cq = 4 * quad;
  const float* const t0 = tile + (size_t)cq * PS;
  const int swz = kVec ? (quad & 7) << 2 : 0;
  for (int c0 = c_begin; c0 < c_end; c0 += CH) {
    const int nch = min(CH, C - c0);
    // the chunk's block of the pooled gradient: (nch, P) contiguous floats
    const float* in = gy + ((size_t)r * C + c0) * P;
    for (int c = warp; c < nch; c += nwarp) {
      float* tp = tile + (size_t)c * PS;
      const float* ip = in + (size_t)c * P;
      if (kVec) {
        const int sw = ((c >> 2) & 7) << 2;
        for (int q4 = lane; q4 < (P >> 2); q4 += 32)
          *reinterpret_cast<float4*>(tp + ((4 * q4) ^ sw)) =
              __ldg(reinterpret_cast<const float4*>(ip) + q4);
      } else {
        for (int p = lane; p < P; p += 32) tp[p] = __ldg(ip + p);
      }
    }
    __syncthreads();
    if (cq < nch) {
      char* img = reinterpret_cast<char*>(gx + (size_t)g.batch * H * W * C4 + (c0 >> 2) + quad);
      for (int ph = row0; ph < outh; ph += rows_per_pass) {
        const int p0 = ph * outw;
        float4 odd = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fetch = [&](int pw, int) {
          if (kVec) {
            if (pw & 1) return odd;
            const float* t = t0 + ((p0 + pw) ^ swz);
            const float2 a = *reinterpret_cast<const float2*>(t);
            const float2 b = *reinterpret_cast<const float2*>(t + PS);
            const float2 c = *reinterpret_cast<const float2*>(t + 2 * PS);
            const float2 d = *reinterpret_cast<const float2*>(t + 3 * PS);
            odd = make_float4(a.y, b.y, c.y, d.y);
            return make_float4(a.x, b.x, c.x, d.x);
          }
          const float* t = t0 + p0 + pw;
          return make_float4(t[0], t[PS], t[2 * PS], t[3 * PS]);
        };
        if (fast) {
          walk_row_bwd<1>(rb[ph], sm->xt, g.grid_w, 0, 1, outw, img, row_bytes, 0, inv, fetch);
        } else {
          for (int pw = 0; pw < outw; ++pw)
            bin_bwd_generic<>(g, ph, pw, H, W, img, row_bytes, px_bytes, inv, fetch(pw, 0));
        }
      }
    }
    if (c0 + CH < c_end) __syncthreads();      // the tile is refilled by the next chunk
  }
}

// ------------------------------------------------ two-pass form of the same operator --
// The walkers above spend ~100 instructions per output float4 on window bookkeeping (ncu:
// issue-bound, 0.6 IPC).  For the usual pooled sizes the bilinear sum is cheaper as two small
// dense products through shared memory, both branch-free:
//   pass A  G[c][ph][x] = sum_k wy[ph][k] * F[y0(ph) + k][xf + x][c]     (vertical blend of
//           every footprint column, float4 loads over channel quads, merged row weights)
//   pass B  out[c][ph][pw] = sum_j wx[pw][j] * G[c][ph][x0(pw) - xf + j]  (merged column
//           weights, 1 / count folded in; lanes run over bins, so the (c, ph, pw) results of
//           a chunk leave as full, contiguous float4 lines of the reference-layout tensor)
// G is channel-major, so pass B needs no transposing tile; its size (channels x outh x
// footprint columns) sets how many channels one round covers (64 for RoIs up to ~13 columns,
// fewer for wide ones).  Same products as the reference, merged weights (a few ulp).
constexpr int kBlendSlots = 8;
constexpr int kTwoPassThreads = 256;
// Two CTAs per SM (127 registers, no spills, none of the re-derived addresses of the 80-register
// build), 60 KB of G, four rows per load batch: same-box A/B at R = 1000, forward GB/s: 3 CTAs /
// 44 KB / 3 rows 1843, 2 CTAs 1956, + 60 KB 2069 (76 KB: 2070), + 4 rows 2104 (2 rows: 1997).
// A grid of at least 16 (instead of 8) CTAs per SM: 2138 (2197 at R = 2000, was 2095).
// The macros exist for such A/B builds (build.py: CMR_EXTRA_NVCC_FLAGS, tools/roi_ab.sh).
#ifndef CMR_TWO_PASS_CTAS
#define CMR_TWO_PASS_CTAS 2
#endif
constexpr int kTwoPassCtas = CMR_TWO_PASS_CTAS;
#ifndef CMR_TWO_PASS_GRID
#define CMR_TWO_PASS_GRID 16
#endif
constexpr int kTwoPassGridCtas = CMR_TWO_PASS_GRID;   // CTAs per SM the grid should at least have
#ifndef CMR_TWO_PASS_G_KB
#define CMR_TWO_PASS_G_KB 60
#endif
constexpr int kTwoPassGFloats = CMR_TWO_PASS_G_KB * 256;      // G per CTA; the rest of the SM's 256 KB stays L1 (pass A re-reads rows)
constexpr int kTwoPassGSlack = 4;                             // floats readable past the last plane
#ifndef CMR_TWO_PASS_ROW_BATCH
#define CMR_TWO_PASS_ROW_BATCH 4
#endif
constexpr int kRowBatch = CMR_TWO_PASS_ROW_BATCH;      // output rows whose loads are in flight together

struct AxisBlend {
  int first;                // index of the first feature row / column
  int n;                    // rows / columns spanned (0: every sample skipped)
  float w[kBlendSlots];
};

// ab[i], i < count: merged taps of output index i along one axis.  *ok = 0 when an output
// spans more than kBlendSlots rows / columns.
__device__ __forceinline__ void build_axis_blends(AxisBlend* ab, int* ok, float start, float bin,
                                                  int grid, int count, int limit, float wscale) {
  for (int e = threadIdx.x; e < count * kBlendSlots; e += blockDim.x) {
    const int i = e / kBlendSlots, k = e - i * kBlendSlots;
    int lo = 1 << 30, hi = -1;
    for (int s = 0; s < grid; ++s) {
      const AxisTap t = axis_tap(start, bin, i, s, grid, limit);
      if (!t.valid) continue;
      lo = min(lo, t.low);
      hi = max(hi, t.high);
    }
    const int idx = lo + k;
    float w = 0.f;
    if (hi >= 0 && idx <= hi) {
      for (int s = 0; s < grid; ++s) {
        const AxisTap t = axis_tap(start, bin, i, s, grid, limit);
        if (!t.valid) continue;
        if (t.low == idx) w += t.h;
        if (t.high == idx) w += t.l;
      }
    }
    ab[i].w[k] = w * wscale;
    if (k == 0) {
      ab[i].first = hi >= 0 ? lo : 0;
      ab[i].n = hi >= 0 ? hi - lo + 1 : 0;
      if (hi - lo + 1 > kBlendSlots) atomicExch(ok, 0);
    }
  }
}

// Pass B's view of one float4 group of bins (p = 4 * p4 + e), laid out so that consecutive
// threads read consecutive 16-byte words: where each bin's taps start in a G plane, and the
// j-th merged column weight (x 1 / count, zero-padded to four taps) of each bin: bin_goff[P/4]
// (int4) and bin_w[4][P/4] (float4) behind G in the dynamic shared memory.

struct TwoPassSmem {
  int ok, xf, ncol, nx_max;
  AxisBlend rows[64];       // outh <= 64; `first` is turned into a byte offset for pass A
  AxisBlend cols[64];       // outw <= 64
};

// Pass B for one thread: its float4 group (four bins) of every channel cl, cl + nct, ...
// NT = taps per bin (the CTA's widest bin; narrower bins carry zero weights and read the
// next floats of G, which exist: G has four floats of slack).
template <int NT>
__device__ __forceinline__ void two_pass_b(const float* __restrict__ g0, float4* __restrict__ op,
                                           int nch, int cl, int nct, int GS, int P4,
                                           const int (&goff)[4], const float (&wx)[4][4]) {
  const float* ga = g0 + goff[0];
  const float* gb = g0 + goff[1];
  const float* gc = g0 + goff[2];
  const float* gd = g0 + goff[3];
  const int gstep = nct * GS, ostep = nct * P4;
  for (int c = cl; c < nch; c += nct) {
    float4 o;
    o.x = wx[0][0] * ga[0]; o.y = wx[1][0] * gb[0]; o.z = wx[2][0] * gc[0]; o.w = wx[3][0] * gd[0];
#pragma unroll
    for (int j = 1; j < NT; ++j) {
      o.x = fmaf(wx[0][j], ga[j], o.x);
      o.y = fmaf(wx[1][j], gb[j], o.y);
      o.z = fmaf(wx[2][j], gc[j], o.z);
      o.w = fmaf(wx[3][j], gd[j], o.w);
    }
    *op = o;
    ga += gstep; gb += gstep; gc += gstep; gd += gstep;
    op += ostep;
  }
}

template <int CH>
__global__ void __launch_bounds__(kTwoPassThreads, kTwoPassCtas)
roi_align_cl2_fwd_kernel(const float4* __restrict__ src, const float* __restrict__ rois,
                         float* __restrict__ dst, int H, int W, int C, int outh, int outw,
                         float scale, int sampling_ratio, int groups, int cpc, int n_img) {
  extern __shared__ __align__(16) unsigned char roi_smem[];
  TwoPassSmem* sm = reinterpret_cast<TwoPassSmem*>(roi_smem);
  float* G = reinterpret_cast<float*>(roi_smem + sizeof(TwoPassSmem));
  int4* bin_goff = reinterpret_cast<int4*>(G + kTwoPassGFloats + kTwoPassGSlack);     // [P / 4]
  float4* bin_w = reinterpret_cast<float4*>(bin_goff + ((outh * outw) >> 2));          // [4][P / 4]
  const int T = blockDim.x, tid = threadIdx.x;
  const int r = blockIdx.x / groups;
  const int c_begin = (blockIdx.x - r * groups) * cpc * CH;
  const int c_end = min(C, c_begin + cpc * CH);
  const int P = outh * outw, P4 = P >> 2;
  const int C4 = C >> 2;
  const int px_bytes = C4 * 16, row_bytes = W * px_bytes;
  const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, outh, outw, sampling_ratio, n_img);
  if (tid == 0) sm->ok = 1;
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, g.inv_count_den);
  build_axis_blends(sm->rows, &sm->ok, g.start_h, g.bin_h, g.grid_h, outh, H, 1.0f);
  build_axis_blends(sm->cols, &sm->ok, g.start_w, g.bin_w, g.grid_w, outw, W, inv);
  __syncthreads();
  if (tid == 0) {
    int xf = 1 << 30, xl = -1, nxm = 1;
    for (int pw = 0; pw < outw; ++pw) {
      if (sm->cols[pw].n <= 0) continue;
      xf = min(xf, sm->cols[pw].first);
      xl = max(xl, sm->cols[pw].first + sm->cols[pw].n - 1);
      nxm = max(nxm, sm->cols[pw].n);
    }
    sm->xf = xl >= 0 ? xf : 0;
    sm->ncol = xl >= 0 ? xl - xf + 1 : 1;
    sm->nx_max = nxm;
  }
  if (tid < outh) sm->rows[tid].first *= row_bytes;       // byte offset of the first row
  __syncthreads();
  const int xf = sm->xf, ncol = sm->ncol, nxm = sm->nx_max;
  // One round = `che` channels x `rs` output rows of G.  Narrow RoIs take all rows and 64
  // channels in one round; for wide ones the rows are cut first (8, then 4 rows: whole float4
  // groups and whole 32-byte sectors of every plane), the channels only after that.
  int rs = outh, che = CH;
  {
    const int cand[3] = {outh, 8, 4};
    bool found = false;
    for (int i = 0; i < 3 && !found; ++i) {
      const int c = cand[i];
      if (c > outh || (i > 0 && (c >= outh || ((c * outw) & 3) || (((outh % c) * outw) & 3)))) continue;
      if (CH * ((c * ncol) | 1) <= kTwoPassGFloats) { rs = c; found = true; }
    }
    if (!found) {
      rs = (outh > 4 && ((4 * outw) & 3) == 0 && (((outh % 4) * outw) & 3) == 0) ? 4 : outh;
      while (che > 4 && che * ((rs * ncol) | 1) > kTwoPassGFloats) che >>= 1;
    }
  }
  const int GS = (rs * ncol) | 1;              // odd plane stride: channels spread over banks
  const bool fits = sm->ok != 0 && che * GS <= kTwoPassGFloats;
  float* out_roi = dst + (size_t)r * C * P;
  const char* img_b = reinterpret_cast<const char*>(src + (size_t)g.batch * H * W * C4);
  if (!fits) {
    // very wide / tall RoI: per-bin evaluation straight from the map (rare, correct, slow)
    for (int c4 = (c_begin >> 2) + tid; c4 < (c_end >> 2); c4 += T) {
      const char* img = img_b + (size_t)c4 * 16;
      for (int p = 0; p < P; ++p) {
        const float4 v = bin_fwd_generic(g, p / outw, p % outw, H, W, img, row_bytes, px_bytes, inv);
        float* o = out_roi + (size_t)(4 * c4) * P + p;
        o[0] = v.x; o[P] = v.y; o[2 * P] = v.z; o[3 * P] = v.w;
      }
    }
    return;
  }
  // pass B reads up to three floats past a bin's last tap (zero weights): every float of G
  // must be finite, so the planes start as zeros (once per CTA)
  for (int i = tid; i < (kTwoPassGFloats + kTwoPassGSlack) / 4; i += T)
    reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = tid; p < P; p += T) {           // per-bin tap tables of pass B
    const int ph = p / outw, pw = p - ph * outw;
    const AxisBlend& cb = sm->cols[pw];
    reinterpret_cast<int*>(bin_goff)[p] = cb.n > 0 ? ph * ncol + cb.first - xf : ph * ncol;
    for (int j = 0; j < 4; ++j)
      reinterpret_cast<float*>(bin_w + j * P4)[p] = j < cb.n ? cb.w[j] : 0.f;
  }
  int lq = 0;
  while ((4 << lq) < che) ++lq;               // che = 4 << lq channels = 1 << lq quads
  const int Qe = 1 << lq;
  const int M = ncol * Qe;                     // (column, quad) pairs of one round
  // pass A: a thread owns one (column, quad) pair and a stride of the round's output rows
  const int nsplit = M <= T ? min(rs, T / M) : 1;
  const int a_k0 = M <= T ? tid % M : tid, a_ph0 = M <= T ? tid / M : 0;
  __syncthreads();                             // G is zero before pass A writes it
  for (int c0 = c_begin; c0 < c_end; c0 += che) {
    const int nch = min(che, c_end - c0);
    for (int pb = 0; pb < outh; pb += rs) {    // the round's rows [pb, pe)
      const int pe = min(outh, pb + rs);
      // ---- pass A: kRowBatch output rows per batch, every load of the batch issued before
      // the first use (the row loop is otherwise one L2 round trip per row: latency-bound)
      if (a_ph0 < nsplit) {
        for (int k = a_k0; k < M; k += T) {
          const int x = k >> lq, q = k & (Qe - 1);
          if (4 * q >= nch) continue;
          const char* col = img_b + ((size_t)(xf + x) * C4 + (c0 >> 2) + q) * 16;
          float* gcol = G + (4 * q) * GS + x - pb * ncol;
          for (int ph = pb + a_ph0; ph < pe; ph += kRowBatch * nsplit) {
            float4 v[kRowBatch][3];
            int nn[kRowBatch];
#pragma unroll
            for (int u = 0; u < kRowBatch; ++u) {
              const int phu = ph + u * nsplit;
              nn[u] = phu < pe ? sm->rows[phu].n : -1;
              const char* p = col + sm->rows[phu < pe ? phu : ph].first;
#pragma unroll
              for (int j = 0; j < 3; ++j)
                if (j < nn[u]) v[u][j] = __ldg(reinterpret_cast<const float4*>(p + j * row_bytes));
            }
#pragma unroll
            for (int u = 0; u < kRowBatch; ++u) {
              if (nn[u] < 0) continue;
              const int phu = ph + u * nsplit;
              const AxisBlend& rbl = sm->rows[phu];
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 3; ++j)
                if (j < nn[u]) fma4(acc, rbl.w[j], v[u][j]);
              for (int j = 3; j < nn[u]; ++j)
                fma4(acc, rbl.w[j],
                     __ldg(reinterpret_cast<const float4*>(col + rbl.first + j * row_bytes)));
              float* gp = gcol + phu * ncol;
              gp[0] = acc.x; gp[GS] = acc.y; gp[2 * GS] = acc.z; gp[3 * GS] = acc.w;
            }
          }
        }
      }
      __syncthreads();
      // ---- pass B: thread = (float4 group of the round's bins, channel lane)
      const int P4r = ((pe - pb) * outw) >> 2;
      const int nct = T / P4r;                 // channel lanes (P4r <= P4 <= T)
      const int p4 = tid % P4r, cl = tid / P4r;
      if (cl < nct) {
        // the four bins' merged column taps, zero-padded to four (re-read per round so that
        // they do not occupy registers during pass A)
        int goff[4];
        float wx[4][4];
        {
          const int g4 = ((pb * outw) >> 2) + p4;
          const int4 go = bin_goff[g4];
          goff[0] = go.x - pb * ncol; goff[1] = go.y - pb * ncol;
          goff[2] = go.z - pb * ncol; goff[3] = go.w - pb * ncol;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w4 = bin_w[j * P4 + g4];
            wx[0][j] = w4.x; wx[1][j] = w4.y; wx[2][j] = w4.z; wx[3][j] = w4.w;
          }
        }
        float4* op = reinterpret_cast<float4*>(out_roi + (size_t)c0 * P + (size_t)pb * outw) +
                     (size_t)cl * P4 + p4;
        const float* g0 = G + cl * GS;
        if (nxm <= 2) {
          two_pass_b<2>(g0, op, nch, cl, nct, GS, P4, goff, wx);
        } else if (nxm == 3) {
          two_pass_b<3>(g0, op, nch, cl, nct, GS, P4, goff, wx);
        } else if (nxm == 4) {
          two_pass_b<4>(g0, op, nch, cl, nct, GS, P4, goff, wx);
        } else {     // bins wider than four columns (explicit sampling_ratio on a small RoI)
          for (int c = cl; c < nch; c += nct, op += nct * P4) {
            const float* Gc = G + (size_t)c * GS;
            float o[4];
            for (int e = 0; e < 4; ++e) {
              const int pl = 4 * p4 + e, phl = pl / outw, pw = pl - phl * outw;
              const AxisBlend& cb = sm->cols[pw];
              const float* gp = Gc + phl * ncol + cb.first - xf;
              float a = 0.f;
              for (int j = 0; j < cb.n; ++j) a = fmaf(cb.w[j], gp[j], a);
              o[e] = a;
            }
            *op = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kClChannels = 64;   // channels per CTA of the two kernels above

bool cl_vec(int outh, int outw) { return (outw & 1) == 0 && ((outh * outw) & 3) == 0; }

size_t cl_smem_bytes(int outh, int outw, int ch = kClChannels) {
  const int P = outh * outw;
  const int PS = cl_vec(outh, outw) ? (P + 31) / 32 * 32 : P + 1;
  return sizeof(RoiTileSmem) + sizeof(float) * (size_t)ch * PS + sizeof(RowBlend) * (size_t)outh;
}

int cl_threads(int outh, int ch = kClChannels) {
  int t = outh * (ch / 4);
  t = (t + 31) / 32 * 32;
  return t > ch * 4 ? ch * 4 : (t < 64 ? 64 : t);
}

// Chunks per CTA: as many as keep >= 8 CTAs per SM in the grid (a power of two <= 16).
int cl_chunks_per_cta(int R, int chunks, int ctas_per_sm = 8) {
  int cpc = 1;
  while (cpc < 16 && cpc * 2 <= chunks &&
         (long long)R * ceil_div(chunks, cpc * 2) >= (long long)ctas_per_sm * sm_count())
    cpc *= 2;
  return cpc;
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
          x, rois, y, C, H, W, outh, outw, spatial_scale, sampling_ratio, chunks, N);
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
          gy, rois, gx, C, H, W, outh, outw, spatial_scale, sampling_ratio, chunks, N);
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
      outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio, round_tf32, N);
  prof_end(as_stream(stream));
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

namespace {
int roi_align_nhwc_bwd_impl(const float* gy, const float* rois, int R, int N, int H, int W, int C,
                            int outh, int outw, int bin_stride, float spatial_scale,
                            int sampling_ratio, float* gx, bool zero_fill, void* stream,
                            bool fixed = false) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && bin_stride >= 1 && C % 4 == 0 && gx);
  const int oh_s = ceil_div(outh, bin_stride), ow_s = ceil_div(outw, bin_stride);
  CMR_REQUIRE((long long)R * oh_s < (1ll << 31));
  CMR_REQUIRE((long long)N * H * W * (C / 4) < (1ll << 31));
  CMR_REQUIRE((long long)H * W * C * 4 < (1ll << 31));   // byte offsets inside an image are ints
  CMR_REQUIRE(R == 0 || (gy && rois));
  // the zero fill of gx is part of the operator: it is inside the timed bracket
  prof_begin(kProfRoiAlignBwd, roi_align_bytes(R, C, oh_s, ow_s, N, H, W), as_stream(stream));
  cudaError_t me = zero_fill ? cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)N * C * H * W,
                                               as_stream(stream))
                             : cudaSuccess;
  if (me != cudaSuccess || R == 0) {
    prof_end(as_stream(stream));
    CMR_CUDA_TRY(me);
    return CMR_OK;
  }
  if (fixed)
    roi_align_nhwc_bwd_kernel<true><<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(gy), rois, reinterpret_cast<float4*>(gx), H, W, C / 4,
        outh, outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio, N);
  else
    roi_align_nhwc_bwd_kernel<false><<<R * oh_s, nhwc_threads(C / 4), 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(gy), rois, reinterpret_cast<float4*>(gx), H, W, C / 4,
        outh, outw, bin_stride, oh_s, ow_s, spatial_scale, sampling_ratio, N);
  prof_end(as_stream(stream));
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
}  // namespace

extern "C" int cmr_roi_align_nhwc_bwd_fixed(const float* gy, const float* rois, int R, int N,
                                            int H, int W, int C, int outh, int outw,
                                            int bin_stride, float spatial_scale,
                                            int sampling_ratio, long long* gx_fixed,
                                            void* stream) {
  return roi_align_nhwc_bwd_impl(gy, rois, R, N, H, W, C, outh, outw, bin_stride, spatial_scale,
                                 sampling_ratio, reinterpret_cast<float*>(gx_fixed), false,
                                 stream, true);
}

extern "C" int cmr_roi_align_nhwc_bwd(const float* gy, const float* rois, int R, int N, int H,
                                      int W, int C, int outh, int outw, int bin_stride,
                                      float spatial_scale, int sampling_ratio, float* gx,
                                      void* stream) {
  return roi_align_nhwc_bwd_impl(gy, rois, R, N, H, W, C, outh, outw, bin_stride, spatial_scale,
                                 sampling_ratio, gx, true, stream);
}

extern "C" int cmr_roi_align_nhwc_bwd_accum(const float* gy, const float* rois, int R, int N,
                                            int H, int W, int C, int outh, int outw,
                                            int bin_stride, float spatial_scale,
                                            int sampling_ratio, float* gx, void* stream) {
  return roi_align_nhwc_bwd_impl(gy, rois, R, N, H, W, C, outh, outw, bin_stride, spatial_scale,
                                 sampling_ratio, gx, false, stream);
}

// ---- reference-layout operator through the channels-last kernels -----------------------
namespace {
bool cl_supported(int N, int C, int H, int W, int R, int outh, int outw) {
  return C % 4 == 0 && (long long)N * H * W * (C / 4) < (1ll << 31) &&
         (long long)H * W * C * 4 < (1ll << 31) &&
         (long long)R * ceil_div(C, kClChannels) < (1ll << 31) &&
         cl_smem_bytes(outh, outw) <= 200 * 1024;
}

template <typename K>
int cl_configure(K kernel, size_t bytes) {
  // (a process-wide maximum: the attribute only ever grows)
  static size_t configured = 0;
  if (bytes > configured) {
    CMR_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)bytes));
    configured = bytes;
  }
  return CMR_OK;
}
}  // namespace

extern "C" int cmr_roi_align_cl_supported(int N, int C, int H, int W, int R, int outh, int outw) {
  return cl_supported(N, C, H, W, R, outh, outw) ? 1 : 0;
}

namespace {
// The two-pass kernel needs whole float4 groups per plane that one CTA's threads can cover.
// CMR_ROI_TWO_PASS=0 keeps the walker kernels (A/B measurements).
bool two_pass_ok(int outh, int outw, const void* y) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CMR_ROI_TWO_PASS");
    on = e ? atoi(e) != 0 : 1;
  }
  const int P = outh * outw;
  return on && (P & 3) == 0 && (P >> 2) <= kTwoPassThreads && outh <= 64 && outw <= 64 &&
         (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

template <bool kVec>
int launch_cl_fwd(const float* x_nhwc, int N, int H, int W, int C, const float* rois, int R,
                  int outh, int outw, float spatial_scale, int sampling_ratio, float* y,
                  cudaStream_t st) {
  const size_t smem = cl_smem_bytes(outh, outw);
  int rc = cl_configure(roi_align_cl_fwd_kernel<kClChannels, kVec>, smem);
  if (rc != CMR_OK) return rc;
  const int chunks = ceil_div(C, kClChannels);
  const int cpc = cl_chunks_per_cta(R, chunks), groups = ceil_div(chunks, cpc);
  prof_begin(kProfRoiAlignApi, roi_align_bytes(R, C, outh, outw, N, H, W), st);
  roi_align_cl_fwd_kernel<kClChannels, kVec><<<R * groups, cl_threads(outh), smem, st>>>(
      reinterpret_cast<const float4*>(x_nhwc), rois, y, H, W, C, outh, outw, spatial_scale,
      sampling_ratio, groups, cpc, N);
  prof_end(st);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

template <bool kVec>
int launch_cl_bwd(const float* gy, const float* rois, int R, int N, int H, int W, int C, int outh,
                  int outw, float spatial_scale, int sampling_ratio, float* gx_nhwc,
                  cudaStream_t st) {
  const size_t smem = cl_smem_bytes(outh, outw, kClBwdChannels);
  int rc = cl_configure(roi_align_cl_bwd_kernel<kClBwdChannels, kVec>, smem);
  if (rc != CMR_OK) return rc;
  prof_begin(kProfRoiAlignApiBwd, roi_align_bytes(R, C, outh, outw, N, H, W), st);
  cudaError_t me = cudaMemsetAsync(gx_nhwc, 0, sizeof(float) * (size_t)N * C * H * W, st);
  if (me != cudaSuccess || R == 0) {
    prof_end(st);
    CMR_CUDA_TRY(me);
    return CMR_OK;
  }
  const int chunks = ceil_div(C, kClBwdChannels);
  const int cpc = cl_chunks_per_cta(R, chunks, kClBwdGridCtas), groups = ceil_div(chunks, cpc);
  roi_align_cl_bwd_kernel<kClBwdChannels, kVec>
      <<<R * groups, cl_threads(outh, kClBwdChannels), smem, st>>>(
      gy, rois, reinterpret_cast<float4*>(gx_nhwc), H, W, C, outh, outw, spatial_scale,
      sampling_ratio, groups, cpc, N);
  prof_end(st);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
}  // namespace

extern "C" int cmr_roi_align_fwd_cl(const float* x_nhwc, int N, int H, int W, int C,
                                    const float* rois, int R, int outh, int outw,
                                    float spatial_scale, int sampling_ratio, float* y,
                                    void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0);
  if (!cl_supported(N, C, H, W, R, outh, outw)) return CMR_ERR_UNSUPPORTED;
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x_nhwc && rois && y);
  cudaStream_t st = as_stream(stream);
  if (two_pass_ok(outh, outw, y)) {
    const size_t smem = sizeof(TwoPassSmem) + sizeof(float) * (kTwoPassGFloats + kTwoPassGSlack) +
                        20 * (size_t)(outh * outw);      // bin_goff + 4 weights per bin
    int rc = cl_configure(roi_align_cl2_fwd_kernel<kClChannels>, smem);
    if (rc != CMR_OK) return rc;
    const int chunks = ceil_div(C, kClChannels);
    const int cpc = cl_chunks_per_cta(R, chunks, kTwoPassGridCtas), groups = ceil_div(chunks, cpc);
    prof_begin(kProfRoiAlignApi, roi_align_bytes(R, C, outh, outw, N, H, W), st);
    roi_align_cl2_fwd_kernel<kClChannels><<<R * groups, kTwoPassThreads, smem, st>>>(
        reinterpret_cast<const float4*>(x_nhwc), rois, y, H, W, C, outh, outw, spatial_scale,
        sampling_ratio, groups, cpc, N);
    prof_end(st);
    CMR_LAUNCH_CHECK();
    return CMR_OK;
  }
  if (cl_vec(outh, outw))
    return launch_cl_fwd<true>(x_nhwc, N, H, W, C, rois, R, outh, outw, spatial_scale,
                               sampling_ratio, y, st);
  return launch_cl_fwd<false>(x_nhwc, N, H, W, C, rois, R, outh, outw, spatial_scale,
                              sampling_ratio, y, st);
}

extern "C" int cmr_roi_align_bwd_cl(const float* gy, const float* rois, int R, int N, int H,
                                    int W, int C, int outh, int outw, float spatial_scale,
                                    int sampling_ratio, float* gx_nhwc, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  CMR_REQUIRE(sampling_ratio >= 0 && gx_nhwc);
  if (!cl_supported(N, C, H, W, R, outh, outw)) return CMR_ERR_UNSUPPORTED;
  CMR_REQUIRE(R == 0 || (gy && rois));
  cudaStream_t st = as_stream(stream);
  if (cl_vec(outh, outw))
    return launch_cl_bwd<true>(gy, rois, R, N, H, W, C, outh, outw, spatial_scale,
                               sampling_ratio, gx_nhwc, st);
  return launch_cl_bwd<false>(gy, rois, R, N, H, W, C, outh, outw, spatial_scale,
                              sampling_ratio, gx_nhwc, st);
}

extern "C" size_t cmr_roi_align_workspace_bytes(int N, int C, int H, int W) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return sizeof(float) * (size_t)N * C * H * W;
}

extern "C" int cmr_roi_align_fwd_ws(const float* x, int N, int C, int H, int W, const float* rois,
                                    int R, int outh, int outw, float spatial_scale,
                                    int sampling_ratio, float* y, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0);
  if (!workspace || workspace_bytes < cmr_roi_align_workspace_bytes(N, C, H, W) ||
      !cl_supported(N, C, H, W, R, outh, outw))
    return cmr_roi_align_fwd(x, N, C, H, W, rois, R, outh, outw, spatial_scale, sampling_ratio,
                             y, stream);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x && rois && y);
  float* xt = static_cast<float*>(workspace);
  int rc = cmr_transpose_batched(x, N, C, H * W, xt, stream);     // (N,C,HW) -> (N,HW,C)
  if (rc != CMR_OK) return rc;
  return cmr_roi_align_fwd_cl(xt, N, H, W, C, rois, R, outh, outw, spatial_scale, sampling_ratio,
                              y, stream);
}

extern "C" int cmr_roi_align_bwd_ws(const float* gy, const float* rois, int R, int N, int C,
                                    int H, int W, int outh, int outw, float spatial_scale,
                                    int sampling_ratio, float* gx, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && outh > 0 && outw > 0 && gx);
  if (!workspace || workspace_bytes < cmr_roi_align_workspace_bytes(N, C, H, W) ||
      !cl_supported(N, C, H, W, R, outh, outw))
    return cmr_roi_align_bwd(gy, rois, R, N, C, H, W, outh, outw, spatial_scale, sampling_ratio,
                             gx, stream);
  float* gt = static_cast<float*>(workspace);
  int rc = cmr_roi_align_bwd_cl(gy, rois, R, N, H, W, C, outh, outw, spatial_scale,
                                sampling_ratio, gt, stream);
  if (rc != CMR_OK) return rc;
  return cmr_transpose_batched(gt, N, H * W, C, gx, stream);       // (N,HW,C) -> (N,C,HW)
}
