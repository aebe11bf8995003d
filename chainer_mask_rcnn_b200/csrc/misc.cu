// HBM-bound helpers around the tensor-core kernels (sm_100a): image packing for
// the stem, pooling, bias-gradient column sums, filter re-layout for the data
// gradient, and the fused MomentumSGD + WeightDecay update.  All NHWC, all
// vectorised 16 B accesses with the channel axis innermost (coalesced).
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cmr {
namespace {

// ------------------------------------------------------------ pack image ---
// (B,3,H,W) fp32 planes -> (B,Hp,Wp,4) pixels {c0,c1,c2,0}, image placed at
// (pad_top, pad_left), zeros elsewhere, values rounded to tf32.  The stem conv
// (models/resnet_extractor.py:63-66: conv1 7x7 stride 2 pad 3) then reads each
// filter row as 32 contiguous floats (8 pixels x 4 channels, the 8th weight
// column zero) -- an implicit GEMM with K = 7 x 32.
__global__ void __launch_bounds__(256)
pack_image_kernel(const float* __restrict__ img, int H, int W, int Hp, int Wp, int pad_top,
                  int pad_left, float4* __restrict__ out) {
  const int b = blockIdx.z;
  const int yp = blockIdx.y;
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  if (xp >= Wp) return;
  const int y = yp - pad_top, x = xp - pad_left;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
    const size_t plane = (size_t)H * W;
    const float* p = img + (size_t)b * 3 * plane + (size_t)y * W + x;
    v.x = tc::round_tf32(__ldg(p));
    v.y = tc::round_tf32(__ldg(p + plane));
    v.z = tc::round_tf32(__ldg(p + 2 * plane));
  }
  out[((size_t)b * Hp + yp) * Wp + xp] = v;
}

// -------------------------------------------------------------- max pool ---
// chainer.functions.max_pooling_2d(ksize, stride, pad) with cover_all
// (models/resnet_extractor.py:67-69): padding behaves as -inf.
__global__ void __launch_bounds__(256)
max_pool_nhwc_kernel(const float4* __restrict__ x, int H, int W, int C4, int k, int stride,
                     int pad, int oh, int ow, float4* __restrict__ y, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  size_t pix = i / C4;
  const int ox = (int)(pix % ow);
  pix /= ow;
  const int oy = (int)(pix % oh);
  const int b = (int)(pix / oh);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const int y0 = oy * stride - pad, x0 = ox * stride - pad;
  for (int fy = 0; fy < k; ++fy) {
    const int iy = y0 + fy;
    if ((unsigned)iy >= (unsigned)H) continue;
    for (int fx = 0; fx < k; ++fx) {
      const int ix = x0 + fx;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float4 v = __ldg(x + ((size_t)(b * H + iy) * W + ix) * C4 + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y);
      m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  y[i] = m;
}

// -------------------------------------------------------------- avg pool ---
// average_pooling_2d(7, stride=7) on a 7x7 map = mean over the HW positions
// (models/mask_rcnn_resnet.py:187).  x (R, HW, C) -> y (R, C).
__global__ void __launch_bounds__(256)
avg_pool_fwd_kernel(const float4* __restrict__ x, int HW, int C4, float4* __restrict__ y,
                    size_t total, int round_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  const size_t r = i / C4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* p = x + r * HW * C4 + c;
  for (int q = 0; q < HW; ++q) {
    const float4 v = __ldg(p + (size_t)q * C4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float inv = 1.0f / (float)HW;
  s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
  if (round_out) {
    s.x = tc::round_tf32(s.x); s.y = tc::round_tf32(s.y);
    s.z = tc::round_tf32(s.z); s.w = tc::round_tf32(s.w);
  }
  y[i] = s;
}

// out[r,q,c] = relu_mask(out[r,q,c] + g[r,c] / HW)   (backward of the mean, added to
// the gradient already in `out`, then masked by the forward activation > 0).
__global__ void __launch_bounds__(256)
avg_pool_bwd_accum_kernel(const float4* __restrict__ g, int HW, int C4, float4* __restrict__ out,
                          const float4* __restrict__ mask, size_t total, int round_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  const size_t r = i / ((size_t)C4 * HW);
  const float inv = 1.0f / (float)HW;
  const float4 gv = __ldg(g + r * C4 + c);
  float4 o = out[i];
  o.x += gv.x * inv; o.y += gv.y * inv; o.z += gv.z * inv; o.w += gv.w * inv;
  if (mask) {
    const float4 m = __ldg(mask + i);
    o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f;
    o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
  }
  if (round_out) {
    o.x = tc::round_tf32(o.x); o.y = tc::round_tf32(o.y);
    o.z = tc::round_tf32(o.z); o.w = tc::round_tf32(o.w);
  }
  out[i] = o;
}

// --------------------------------------------------------------- col sum ---
// out[j] += sum_m g[m*ld + c0 + j]  (bias gradients: gb = gy.sum over pixels).
// Block = 32 columns x 8 row lanes; grid.y splits the rows.
__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ g, long long M, int ld, int c0, int n,
               float* __restrict__ out, int rows_per_block, int fixed) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const long long m0 = (long long)blockIdx.y * rows_per_block;
  const long long m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  if (j < n)
    for (long long m = m0 + ty; m < m1; m += 8) s += __ldg(g + m * ld + c0 + j);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][tx];
    if (fixed) red_fixed(reinterpret_cast<long long*>(out) + j, t);   // deterministic mode
    else atomicAdd(out + j, t);
  }
}

// ------------------------------------------------------ filter re-layout ---
// out[(i*T + (flip ? T-1-t : t))*ld_out + col0 + o] = round_tf32(w[o*so + t*st + i] * scale[o])
// (the filter bank of the data-gradient GEMM: transposed, spatially flipped for a
// correlation, with the frozen AffineChannel2D slope folded in, since
// gx = conv^T(W_affine * gy); functions/affine_channel_2d.py:48-52).
__global__ void __launch_bounds__(256)
prep_dgrad_weight_kernel(const float* __restrict__ w, int O, int T, int I, long long so,
                         long long st, const float* __restrict__ scale, int flip,
                         float* __restrict__ out, int ld_out, int col0) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, i = i0 + tx;
    float v = 0.f;
    if (o < O && i < I) {
      v = __ldg(w + o * so + t * st + i);
      if (scale) v *= __ldg(scale + o);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  const int to = flip ? T - 1 - t : t;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, o = o0 + tx;
    if (i < I && o < O)
      out[((size_t)i * T + to) * ld_out + col0 + o] = tc::round_tf32(tile[tx][r]);
  }
}

// All layers' re-layouts in one launch: CTA -> (descriptor, 32x32 tile, tap).
__global__ void __launch_bounds__(256)
prep_dgrad_weight_batch_kernel(const cmr_prep_desc* __restrict__ descs, int n_desc) {
  __shared__ float tile[32][33];
  __shared__ int which;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_desc - 1;             // last descriptor with tile_begin <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].tile_begin <= (int)blockIdx.x) lo = mid;
      else hi = mid - 1;
    }
    which = lo;
  }
  __syncthreads();
  const cmr_prep_desc d = descs[which];
  int local = blockIdx.x - d.tile_begin;
  const int ti = (d.I + 31) / 32, to_ = (d.O + 31) / 32;
  const int t = local / (ti * to_);
  local -= t * ti * to_;
  const int i0 = (local % ti) * 32, o0 = (local / ti) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, i = i0 + tx;
    float v = 0.f;
    if (o < d.O && i < d.I) {
      v = __ldg(d.w + o * d.stride_o + t * d.stride_t + i);
      if (d.scale) v *= __ldg(d.scale + o);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  const int tt = d.flip ? d.T - 1 - t : t;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, o = o0 + tx;
    if (i < d.I && o < d.O)
      d.out[((size_t)i * d.T + tt) * d.ld_out + d.col0 + o] = tc::round_tf32(tile[tx][r]);
  }
}

// ------------------------------------------------------------------- SGD ---
// chainer.optimizers.MomentumSGD + optimizer_hooks.WeightDecay
// (examples/train_common.py:176-180):  g' = grad_scale*g + wd*p;
// v = momentum*v - lr*g';  p += v.
__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ v,
                    float4* __restrict__ rounded, size_t n4, float lr, float momentum, float wd,
                    float grad_scale) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p[i], vv = v[i];
    const float4 gg = __ldg(g + i);
    vv.x = momentum * vv.x - lr * (grad_scale * gg.x + wd * pp.x);
    vv.y = momentum * vv.y - lr * (grad_scale * gg.y + wd * pp.y);
    vv.z = momentum * vv.z - lr * (grad_scale * gg.z + wd * pp.z);
    vv.w = momentum * vv.w - lr * (grad_scale * gg.w + wd * pp.w);
    pp.x += vv.x; pp.y += vv.y; pp.z += vv.z; pp.w += vv.w;
    p[i] = pp;
    v[i] = vv;
    if (rounded)   // the tf32 copy the next forward pass reads (saves a pass over the weights)
      rounded[i] = make_float4(tc::round_tf32(pp.x), tc::round_tf32(pp.y), tc::round_tf32(pp.z),
                               tc::round_tf32(pp.w));
  }
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_pack_image_nhwc4(const float* img, int B, int H, int W, int Hp, int Wp,
                                    int pad_top, int pad_left, float* out, void* stream) {
  CMR_REQUIRE(img && out && B > 0 && H > 0 && W > 0 && Hp >= H + pad_top && Wp >= W + pad_left);
  CMR_REQUIRE(pad_top >= 0 && pad_left >= 0 && B < 65536 && Hp < 65536);
  dim3 grid(ceil_div(Wp, 256), Hp, B);
  pack_image_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, H, W, Hp, Wp, pad_top, pad_left,
                                                         reinterpret_cast<float4*>(out));
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_max_pool_nhwc(const float* x, int B, int H, int W, int C, int ksize,
                                 int stride, int pad, int out_h, int out_w, float* y,
                                 void* stream) {
  CMR_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
  CMR_REQUIRE(ksize > 0 && stride > 0 && pad >= 0 && out_h > 0 && out_w > 0);
  // every window must contain at least one input pixel
  CMR_REQUIRE((out_h - 1) * stride - pad < H && (out_w - 1) * stride - pad < W && pad < ksize);
  const size_t total = (size_t)B * out_h * out_w * (C / 4);
  max_pool_nhwc_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), H, W, C / 4, ksize, stride, pad, out_h, out_w,
      reinterpret_cast<float4*>(y), total);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_avg_pool_nhwc_fwd(const float* x, int R, int HW, int C, float* y,
                                     int round_tf32, void* stream) {
  CMR_REQUIRE(R >= 0 && HW > 0 && C > 0 && C % 4 == 0);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(x && y);
  const size_t total = (size_t)R * (C / 4);
  avg_pool_fwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), HW, C / 4, reinterpret_cast<float4*>(y), total,
      round_tf32);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_avg_pool_nhwc_bwd_accum(const float* g, int R, int HW, int C, float* out,
                                           const float* mask, int round_tf32, void* stream) {
  CMR_REQUIRE(R >= 0 && HW > 0 && C > 0 && C % 4 == 0);
  if (R == 0) return CMR_OK;
  CMR_REQUIRE(g && out);
  const size_t total = (size_t)R * HW * (C / 4);
  avg_pool_bwd_accum_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(g), HW, C / 4, reinterpret_cast<float4*>(out),
      reinterpret_cast<const float4*>(mask), total, round_tf32);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

namespace {
int col_sum_impl(const float* g, long long M, int ld, int c0, int n, float* out, int fixed,
                 void* stream);
}

extern "C" int cmr_col_sum(const float* g, long long M, int ld, int c0, int n, float* out,
                           void* stream) {
  return col_sum_impl(g, M, ld, c0, n, out, 0, stream);
}

extern "C" int cmr_col_sum_fixed(const float* g, long long M, int ld, int c0, int n,
                                 long long* out_fixed, void* stream) {
  return col_sum_impl(g, M, ld, c0, n, reinterpret_cast<float*>(out_fixed), 1, stream);
}

namespace {
int col_sum_impl(const float* g, long long M, int ld, int c0, int n, float* out, int fixed,
                 void* stream) {
  CMR_REQUIRE(M >= 0 && n > 0 && ld >= c0 + n && c0 >= 0 && out);
  cudaStream_t st = as_stream(stream);
  CMR_CUDA_TRY(cudaMemsetAsync(out, 0, (fixed ? sizeof(long long) : sizeof(float)) * n, st));
  if (M == 0) return CMR_OK;
  CMR_REQUIRE(g);
  const int col_blocks = ceil_div(n, 32);
  long long splits = (4LL * sm_count() + col_blocks - 1) / col_blocks;
  const long long max_splits = (M + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int rows_per_block = (int)((M + splits - 1) / splits);
  dim3 grid(col_blocks, (unsigned)((M + rows_per_block - 1) / rows_per_block));
  col_sum_kernel<<<grid, 256, 0, st>>>(g, M, ld, c0, n, out, rows_per_block, fixed);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
}  // namespace

extern "C" int cmr_prep_dgrad_weight(const float* w, int O, int T, int I, long long stride_o,
                                     long long stride_t, const float* scale, int flip,
                                     float* out, int ld_out, int col0, void* stream) {
  CMR_REQUIRE(w && out && O > 0 && T > 0 && I > 0 && T < 65536);
  CMR_REQUIRE(col0 >= 0 && ld_out >= col0 + O);
  dim3 grid(ceil_div(I, 32), ceil_div(O, 32), T);
  CMR_REQUIRE(grid.y < 65536);
  prep_dgrad_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, O, T, I, stride_o, stride_t,
                                                                scale, flip, out, ld_out, col0);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_prep_dgrad_weight_batch(const cmr_prep_desc* descs_dev, int n_desc,
                                           int total_tiles, void* stream) {
  CMR_REQUIRE(descs_dev && n_desc > 0 && total_tiles > 0);
  prep_dgrad_weight_batch_kernel<<<total_tiles, 256, 0, as_stream(stream)>>>(descs_dev, n_desc);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_sgd_momentum(float* param, const float* grad, float* velocity, size_t n,
                                float lr, float momentum, float weight_decay, float grad_scale,
                                float* param_tf32, void* stream) {
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(param && grad && velocity && n % 4 == 0);
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                reinterpret_cast<uintptr_t>(velocity) | reinterpret_cast<uintptr_t>(param_tf32)) &
               15) == 0);
  const size_t n4 = n / 4;
  const int blocks = (int)min((size_t)sm_count() * 8, (n4 + 255) / 256);
  sgd_momentum_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
      reinterpret_cast<float4*>(velocity), reinterpret_cast<float4*>(param_tf32), n4, lr, momentum,
      weight_decay, grad_scale);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
