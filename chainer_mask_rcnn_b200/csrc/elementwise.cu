// Small element-wise helpers (HBM-bound, vectorised).
#include "common.cuh"
#include "tc_common.cuh"

namespace cmr {
namespace {
__global__ void __launch_bounds__(256)
round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (size_t k = i; k < n4; k += stride) {
      float4 v = reinterpret_cast<const float4*>(in)[k];
      v.x = tc::round_tf32(v.x); v.y = tc::round_tf32(v.y);
      v.z = tc::round_tf32(v.z); v.w = tc::round_tf32(v.w);
      reinterpret_cast<float4*>(out)[k] = v;
    }
    for (size_t k = n4 * 4 + i; k < n; k += stride) out[k] = tc::round_tf32(in[k]);
  } else {
    for (size_t k = i; k < n; k += stride) out[k] = tc::round_tf32(in[k]);
  }
}

// x = hi + lo + O(2^-22 |x|) with hi = tf32(x), lo = tf32(x - hi): the operand split of the
// 3 x TF32 parity mode.  One thread per channel quad of a row.
__global__ void __launch_bounds__(256)
split_tf32x3_kernel(const float* __restrict__ x, float* __restrict__ out, size_t rows, int c_in,
                    int c_pad, int order) {
  const int qpr = c_pad >> 2;
  const size_t total = rows * (size_t)qpr;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool vec = (c_in & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t row = i / qpr;
    const int c = (int)(i - row * qpr) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = x + row * (size_t)c_in + c;
    if (vec && c + 3 < c_in) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < c_in) v[e] = __ldg(src + e);
    }
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = tc::round_tf32(v[e]);
      lo[e] = tc::round_tf32(__fsub_rn(v[e], hi[e]));
    }
    const float4 h4 = make_float4(hi[0], hi[1], hi[2], hi[3]);
    const float4 l4 = make_float4(lo[0], lo[1], lo[2], lo[3]);
    float4* dst = reinterpret_cast<float4*>(out + row * (size_t)(3 * c_pad) + c);
    dst[0] = h4;
    dst[qpr] = order == 0 ? l4 : h4;
    dst[2 * qpr] = order == 0 ? h4 : l4;
  }
}

// out = mask > 0 ? g : 0, optionally rounded to tf32 (ReLU backward as its own pass).
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float4* __restrict__ g, const float4* __restrict__ mask,
                 float4* __restrict__ out, size_t n4, int round_out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = g[i];
    const float4 m = __ldg(mask + i);
    v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
    v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    if (round_out) {
      v.x = tc::round_tf32(v.x); v.y = tc::round_tf32(v.y);
      v.z = tc::round_tf32(v.z); v.w = tc::round_tf32(v.w);
    }
    out[i] = v;
  }
}

// Fixed-point accumulators (common.cuh) back to fp32: out = (accumulate ? out : 0) +
// float(in * 2^-40); in is zeroed for the next step when zero_src is set.
__global__ void __launch_bounds__(256)
fixed_to_float_kernel(long long* __restrict__ in, float* __restrict__ out, size_t n,
                      int accumulate, int zero_src) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = (float)((double)in[i] * (1.0 / 1099511627776.0));
    out[i] = accumulate ? out[i] + v : v;
    if (zero_src) in[i] = 0;
  }
}
}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_fixed_to_float(long long* in, float* out, size_t n, int accumulate,
                                  int zero_src, void* stream) {
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(in && out);
  const int blocks = (int)min((size_t)sm_count() * 8, (n + 255) / 256);
  fixed_to_float_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, n, accumulate, zero_src);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_relu_mask(const float* g, const float* mask, float* out, size_t n,
                             int round_tf32, void* stream) {
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(g && mask && out && n % 4 == 0);
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mask) |
                reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  const int blocks = (int)min((size_t)sm_count() * 8, (n / 4 + 255) / 256);
  relu_mask_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(mask),
      reinterpret_cast<float4*>(out), n / 4, round_tf32);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_split_tf32x3(const float* x, size_t rows, int c_in, int c_pad, int order,
                                float* out, void* stream) {
  if (rows == 0) return CMR_OK;
  CMR_REQUIRE(x && out && c_in > 0 && c_pad >= c_in && (c_pad & 3) == 0 && (order == 0 || order == 1));
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const size_t quads = rows * (size_t)(c_pad >> 2);
  const int blocks = (int)min((size_t)sm_count() * 8, (quads + 255) / 256);
  split_tf32x3_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, out, rows, c_in, c_pad, order);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_round_tf32(const float* in, float* out, size_t n, void* stream) {
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(in && out);
  const int blocks = (int)min((size_t)sm_count() * 8, (n / 4 + 255) / 256 + 1);
  round_tf32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, n);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
