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
}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_round_tf32(const float* in, float* out, size_t n, void* stream) {
  if (n == 0) return CMR_OK;
  CMR_REQUIRE(in && out);
  const int blocks = (int)min((size_t)sm_count() * 8, (n / 4 + 255) / 256 + 1);
  round_tf32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, n);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
